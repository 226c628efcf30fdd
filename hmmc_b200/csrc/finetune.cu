// Fine-tune (hierarchical matching) head: loose_similarity, CrossEn and the fused
// symmetric-CE loss over the text x video and the F text x frame similarity matrices.
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace hmmc {

// dx = (g - x_hat (x_hat . g)) / ||x||  for x_hat = x/||x|| (no eps, loose_similarity).
// One warp per row; optional row remap of the *destination/source* x rows:
//   src_row(r) = (r % remap_B) * remap_F + (r / remap_B - remap_skip)   when remap_B > 0
// lets the gallery-ordered gradient rows (f*B + j) land in the [B,F,D] frame layout.
__global__ void unnormalize_grad_kernel(const float* __restrict__ x, const float* __restrict__ ghat,
                                        float* __restrict__ dx, int64_t R, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* xr = x + row * D;
  const float* gr = ghat + row * D;
  float ss = 0.f, xg = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d];
    ss = fmaf(v, v, ss);
    xg = fmaf(v, gr[d], xg);
  }
  ss = warp_sum(ss);
  xg = warp_sum(xg);
  const float n = sqrtf(ss);
  const float proj = xg / (n * n);            // (x_hat . g) / ||x||  folded: x_hat*(x_hat.g) = x * (x.g)/n^2
  for (int d = lane; d < D; d += 32) dx[row * D + d] = (gr[d] - xr[d] * proj) / n;
}

// ------------------------------------------------------------------ CrossEn
// one block per row: lse_i = logsumexp(S[i,:]); row_loss[i] = (lse_i - S[i,i]) / B;
// dS[i,j] = (exp(S[i,j]-lse_i) - [i==j]) / B
__global__ void cross_en_row_kernel(const float* __restrict__ S, int64_t lds, int B, float* __restrict__ row_loss,
                                    float* __restrict__ dS, int64_t ldds) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  const float* row = S + int64_t(i) * lds;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) m = fmaxf(m, row[j]);
  m = block_max(m, red);
  float s = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) s += expf(row[j] - m);
  s = block_sum(s, red);
  const float lse = m + logf(s);
  if (threadIdx.x == 0) row_loss[i] = (lse - row[i]) / float(B);
  if (dS != nullptr) {
    float* drow = dS + int64_t(i) * ldds;
    for (int j = threadIdx.x; j < B; j += blockDim.x)
      drow[j] = (expf(row[j] - lse) - (j == i ? 1.f : 0.f)) / float(B);
  }
}

__global__ void sum_to_scalar_kernel(const float* __restrict__ v, int n, float* __restrict__ out, int accumulate) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + acc : acc;
}

// ------------------------------------------------------------------ fused fine-tune head
// Gallery rows are ordered g = fp*B + j with fp = 0 the video embedding and fp = 1..F frame fp-1,
// so the similarity matrix S_all [B, (1+F)*B] holds the 1+F square blocks side by side.
__global__ void gallery_norm_kernel(const float* __restrict__ video, const float* __restrict__ frames, int B, int F,
                                    int voff, int D, float* __restrict__ ghat) {
  const int lane = threadIdx.x & 31;
  const int64_t g = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= int64_t(voff + F) * B) return;
  const int fp = int(g / B), j = int(g % B);
  const float* x = (fp < voff) ? video + int64_t(j) * D : frames + (int64_t(j) * F + (fp - voff)) * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = x[d]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float n = sqrtf(ss);
  for (int d = lane; d < D; d += 32) ghat[g * D + d] = x[d] / n;
}

// row statistics: one block per text row i, loops over the 1+F blocks
__global__ void symce_row_lse_kernel(const float* __restrict__ S, int B, int NB, float* __restrict__ lse_row) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  for (int fp = 0; fp < NB; ++fp) {
    const float* row = S + int64_t(i) * NB * B + int64_t(fp) * B;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < B; j += blockDim.x) m = fmaxf(m, row[j]);
    m = block_max(m, red);
    float s = 0.f;
    for (int j = threadIdx.x; j < B; j += blockDim.x) s += expf(row[j] - m);
    s = block_sum(s, red);
    if (threadIdx.x == 0) lse_row[fp * B + i] = m + logf(s);
  }
}

// column statistics: block = 32 consecutive columns of S_all, 8 warps stride the rows
__global__ void __launch_bounds__(256)
symce_col_lse_kernel(const float* __restrict__ S, int B, int NB, float* __restrict__ lse_col) {
  __shared__ float sm[8][32], ss[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ncol = int64_t(NB) * B;
  const int64_t col = int64_t(blockIdx.x) * 32 + lane;
  float m = -INFINITY, s = 0.f;
  if (col < ncol) {
    for (int i = warp; i < B; i += 8) {
      const float v = S[int64_t(i) * ncol + col];
      if (v > m) { s = s * expf(m - v) + 1.f; m = v; } else { s += expf(v - m); }
    }
  }
  sm[warp][lane] = m;
  ss[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && col < ncol) {
    float M = sm[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) M = fmaxf(M, sm[w][lane]);
    float Ssum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) Ssum += (sm[w][lane] == -INFINITY) ? 0.f : ss[w][lane] * expf(sm[w][lane] - M);
    lse_col[col] = M + logf(Ssum);      // index fp*B + j == col
  }
}

// loss terms and G = dL/dS in place:  G_ij = w_fp/B * (exp(S_ij - lse_row) + exp(S_ij - lse_col) - 2 [i==j])
// block per row i; also the row's loss share  w_fp/B * (lse_row + lse_col - 2 S_ii)
__global__ void symce_grad_kernel(float* __restrict__ S, int B, int NB, const float* __restrict__ lse_row,
                                  const float* __restrict__ lse_col, int voff, float w0, float wf,
                                  float* __restrict__ row_loss, int write_grad) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  const int64_t ncol = int64_t(NB) * B;
  float* row = S + int64_t(i) * ncol;
  float loss = 0.f;
  for (int64_t c = threadIdx.x; c < ncol; c += blockDim.x) {
    const int fp = int(c / B), j = int(c - int64_t(fp) * B);
    const float w = ((fp < voff) ? w0 : wf) / float(B);
    const float v = row[c];
    const float lr = lse_row[fp * B + i], lc = lse_col[c];
    if (j == i) loss += w * (lr + lc - 2.f * v);
    if (write_grad) row[c] = w * (expf(v - lr) + expf(v - lc) - (j == i ? 2.f : 0.f));
  }
  loss = block_sum(loss, red);
  if (threadIdx.x == 0) row_loss[i] = loss;
}

// scatter gallery-ordered gradient rows back: dvideo[j] = dG[0*B + j], dframes[j,f] = dG[(f+1)*B + j]
__global__ void gallery_scatter_kernel(const float* __restrict__ dG, int B, int F, int voff, int D,
                                       float* __restrict__ dvideo, float* __restrict__ dframes) {
  const int64_t g = blockIdx.x;
  const int fp = int(g / B), j = int(g % B);
  float* dst = (fp < voff) ? dvideo + int64_t(j) * D : dframes + (int64_t(j) * F + (fp - voff)) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = dG[g * D + d];
}

// ------------------------------------------------------------------ fused kernels of the tensor-core path
// Row r of the stacked operand matrix: r < B -> text row r; else gallery row g = r - B = fp*B + j.
// The three operands with their row strides (elements): contiguous tensors have ldt = ldv = D,
// ldf = F*D; the packed layout [text | video | frames] of the all-gather has all three = (2+F)*D.
struct SymLossAcc {
  unsigned long long sum;       // units of 2^-36
  unsigned arrived;
  unsigned pad;
};
struct SymOperands {
  const float* text; int64_t ldt;
  const float* video; int64_t ldv;
  const float* frames; int64_t ldf;
};
struct SymGrads {
  float* text; int64_t ldt;
  float* video; int64_t ldv;
  float* frames; int64_t ldf;
};
__device__ __forceinline__ const float* symce_src_row(const SymOperands& o, int B, int voff, int D, int r) {
  if (r < B) return o.text + int64_t(r) * o.ldt;
  const int g = r - B, fp = g / B, j = g - fp * B;
  return (fp < voff) ? o.video + int64_t(j) * o.ldv : o.frames + int64_t(j) * o.ldf + int64_t(fp - voff) * D;
}

// Normalise (x / ||x||, no eps: loose_similarity) the B text rows and the NG gallery rows and write the
// bf16 plane packs the three GEMMs read: straight [rows, planes*D] and transposed [D, planes*rows].
// Block = 32 consecutive rows (B % 32 == 0, so a block never mixes text and gallery rows), staged whole
// in shared memory (32 x (D+1) floats): every input element is read once, one block-wide sync.
__global__ void __launch_bounds__(1024)
symce_prep_kernel(SymOperands src, int B, int F, int voff, int D, int planes, int Bk, int NGk,
                  __nv_bfloat16* __restrict__ Tp,
                  __nv_bfloat16* __restrict__ TTp, __nv_bfloat16* __restrict__ Gp, __nv_bfloat16* __restrict__ GTp,
                  SymLossAcc* __restrict__ loss_acc) {
  extern __shared__ float tile[];                 // [32][D + 1]
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) *loss_acc = SymLossAcc{0ull, 0u, 0u};
  const int ld = D + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 32;
  const bool is_text = r0 < B;
  const int R = is_text ? Bk : NGk;               // K extent (padded to 64) of the transposed pack
  const int q0 = is_text ? r0 : r0 - B;           // first row inside its operand
  __nv_bfloat16* straight = is_text ? Tp : Gp;
  __nv_bfloat16* transposed = is_text ? TTp : GTp;
  {
    const int rr = warp;                          // 32 warps: one row each
    const float* x = symce_src_row(src, B, voff, D, r0 + rr);
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = x[d]; tile[rr * ld + d] = v; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float n = sqrtf(ss);
    for (int d = lane; d < D; d += 32) {
      const float v = tile[rr * ld + d] / n;
      tile[rr * ld + d] = v;
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      const int64_t o = int64_t(q0 + rr) * planes * D + d;
      straight[o] = hi;
      if (planes == 2) straight[o + D] = lo;
    }
  }
  if (transposed == nullptr) return;
  __syncthreads();
  for (int d = warp; d < D; d += 32) {
    __nv_bfloat16 hi, lo;
    split_bf16(tile[lane * ld + d], hi, lo);
    const int64_t o = int64_t(d) * planes * R + q0 + lane;
    transposed[o] = hi;
    if (planes == 2) transposed[o + R] = lo;
  }
}

// Row and column log-sum-exp of S_all in one launch, and the loss
//   loss = sum_fp w_fp/B sum_i (lse_row + lse_col - 2 S_ii).
// Blocks [0, row_blocks) take one warp per (text row, similarity block) pair, 8 pairs per block; the remaining
// blocks take 32 columns each.  Every block adds its share of the loss to a 64-bit fixed-point accumulator with a
// fire-and-forget reduction (integer sums: the same from run to run) and signals with a release reduction; nobody
// waits for an answer.  The highest block waits for the arrival count and converts (see finish_losses in
// pretrain.cu: the publish / fence / returning-atomic hand-over kept every block resident for two round trips).
constexpr float SYM_FIX_SCALE = 68719476736.0f;          // 2^36: |loss| < 2^27

__global__ void __launch_bounds__(256)
symce_lse_kernel(const float* __restrict__ S, int B, int NB, int row_blocks, float* __restrict__ lse_row,
                 float* __restrict__ lse_col, int voff, float w0, float wf, float* __restrict__ loss_out,
                 SymLossAcc* __restrict__ acc) {
  __shared__ float sm[8][32], ss[8][32], sd[32];
  __shared__ float wsum[8];
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ncol = int64_t(NB) * B;
  float share = 0.f;                                 // this thread's part of the loss (lane 0 / warp 0 only)
  if (int(blockIdx.x) < row_blocks) {
    const int pair = blockIdx.x * 8 + warp;          // = i * NB + fp: consecutive warps read consecutive memory
    if (pair < B * NB) {
      const int i = pair / NB, fp = pair - i * NB;
      const float* row = S + int64_t(i) * ncol + int64_t(fp) * B;
      float m = -INFINITY;
      for (int j = lane; j < B; j += 32) m = fmaxf(m, row[j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < B; j += 32) s += expf(row[j] - m);
      s = warp_sum(s);
      if (lane == 0) {
        const float l = m + logf(s);
        lse_row[fp * B + i] = l;
        share = (((fp < voff) ? w0 : wf) / float(B)) * (l - row[i]);
      }
    }
  } else {
    const int64_t col = int64_t(blockIdx.x - row_blocks) * 32 + lane;
    const int fp = int(col / B), j = int(col - int64_t(fp) * B);       // the column's own text row
    // four independent running (max, sum) pairs per thread: rows warp, warp+8, .. dealt round-robin to them
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, s[4] = {0.f, 0.f, 0.f, 0.f};
    float diag = 0.f;
    if (col < ncol) {
      for (int i0 = warp; i0 < B; i0 += 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 8 * u;
          if (i < B) {
            const float v = S[int64_t(i) * ncol + col];
            if (i == j) diag = v;
            if (v > m[u]) { s[u] = s[u] * expf(m[u] - v) + 1.f; m[u] = v; } else { s[u] += expf(v - m[u]); }
          }
        }
      }
    }
    float M = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), Sx = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) Sx += (m[u] == -INFINITY) ? 0.f : s[u] * expf(m[u] - M);
    sm[warp][lane] = M;
    ss[warp][lane] = Sx;
    if ((j & 7) == warp && col < ncol) sd[lane] = diag;                // the warp that read row j holds S_jj
    __syncthreads();
    if (warp == 0 && col < ncol) {
      float Mx = sm[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) Mx = fmaxf(Mx, sm[w][lane]);
      float Ssum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) Ssum += (sm[w][lane] == -INFINITY) ? 0.f : ss[w][lane] * expf(sm[w][lane] - Mx);
      const float l = Mx + logf(Ssum);
      lse_col[col] = l;
      share = warp_sum((((fp < voff) ? w0 : wf) / float(B)) * (l - sd[lane]));
    } else if (warp == 0) {
      share = warp_sum(0.f);
    }
  }
  // ---- the block's share of the loss: warp order, then one fixed-point reduction
  if (lane == 0) wsum[warp] = share;
  __syncthreads();
  if (threadIdx.x != 0) return;
  float v = 0.f;
  if (int(blockIdx.x) < row_blocks) {
#pragma unroll
    for (int w = 0; w < 8; ++w) v += wsum[w];
  } else {
    v = wsum[0];
  }
  atomicAdd(&acc->sum, static_cast<unsigned long long>(__float2ll_rn(v * SYM_FIX_SCALE)));
  ptx::red_release_add(&acc->arrived, 1u);
  if (blockIdx.x != gridDim.x - 1) return;
  const long long t0 = clock64();
  while (ptx::ld_acquire(&acc->arrived) != gridDim.x)
    if (clock64() - t0 > (1ll << 33)) __trap();
  loss_out[0] = float(double(static_cast<long long>(__ldcg(&acc->sum))) * (1.0 / double(SYM_FIX_SCALE)));
  acc->sum = 0ull;
  acc->arrived = 0u;
}

// G = dL/dS computed on the fly from S and the two LSE vectors and written straight as the bf16
// plane packs of the backward GEMMs: Sp [B, planes*NG] and STp [NG, planes*B].  32x32 tiles.
__global__ void __launch_bounds__(256)
symce_gradpack_kernel(const float* __restrict__ S, int B, int NB, const float* __restrict__ lse_row,
                      const float* __restrict__ lse_col, int voff, float w0, float wf, int planes, int Bk, int NGk,
                      __nv_bfloat16* __restrict__ Sp, __nv_bfloat16* __restrict__ STp) {
  __shared__ float tile[32][33];
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NG = NB * B;
  const int c0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int c = c0 + lane;
  const int fp = c0 / B;                       // B % 32 == 0: a tile lies inside one similarity block
  const int j = c - fp * B;
  const float w = ((fp < voff) ? w0 : wf) / float(B);
  const float lc = lse_col[c];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ii = warp * 4 + k, i = i0 + ii;
    const float v = S[int64_t(i) * NG + c];
    const float g = w * (expf(v - lse_row[fp * B + i]) + expf(v - lc) - (j == i ? 2.f : 0.f));
    __nv_bfloat16 hi, lo;
    split_bf16(g, hi, lo);
    const int64_t o = int64_t(i) * planes * NGk + c;
    Sp[o] = hi;
    if (planes == 2) Sp[o + NGk] = lo;
    tile[ii][lane] = g;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cc = warp * 4 + k;
    __nv_bfloat16 hi, lo;
    split_bf16(tile[lane][cc], hi, lo);
    const int64_t o = int64_t(c0 + cc) * planes * Bk + i0 + lane;
    STp[o] = hi;
    if (planes == 2) STp[o + Bk] = lo;
  }
}

// Chain the gradients w.r.t. the normalised rows back to the raw inputs, all operands in one launch.
// Text row i: g = fixed-order sum of the split-K partials; gallery row g: read in gallery order, written
// to dvideo / dframes in their own layouts.  One warp per row, the row kept in registers (D <= 32*NV).
template <int NV>
__global__ void symce_unnorm_kernel(SymOperands src, int B, int F, int voff, int D,
                                    const float* __restrict__ gt_parts, int n_splits, int64_t split_stride,
                                    const float* __restrict__ gg, SymGrads out) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int NG = (voff + F) * B;
  if (r >= B + NG) return;
  const float* x = symce_src_row(src, B, voff, D, r);
  float* dst;
  const float* grow;
  int ns = 1;
  if (r < B) {
    dst = out.text ? out.text + int64_t(r) * out.ldt : nullptr;
    grow = gt_parts + int64_t(r) * D;
    ns = n_splits;
  } else {
    const int g = r - B, fp = g / B, j = g - fp * B;
    dst = (fp < voff) ? (out.video ? out.video + int64_t(j) * out.ldv : nullptr)
                      : (out.frames ? out.frames + int64_t(j) * out.ldf + int64_t(fp - voff) * D : nullptr);
    grow = gg + int64_t(g) * D;
  }
  if (dst == nullptr) return;
  float gv[NV], xv[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int d = lane + 32 * k;
    gv[k] = 0.f;
    xv[k] = (d < D) ? x[d] : 0.f;
  }
#pragma unroll 4
  for (int s0 = 0; s0 < ns; ++s0) {
    const float* gs = grow + int64_t(s0) * split_stride;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int d = lane + 32 * k;
      if (d < D) gv[k] += gs[d];
    }
  }
  float ss = 0.f, xg = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) { ss = fmaf(xv[k], xv[k], ss); xg = fmaf(xv[k], gv[k], xg); }
  ss = warp_sum(ss);
  xg = warp_sum(xg);
  const float n = sqrtf(ss);
  const float proj = xg / (n * n);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int d = lane + 32 * k;
    if (d < D) dst[d] = (gv[k] - xv[k] * proj) / n;
  }
}

// The same with 16-byte accesses: D == 128 * V, every row is V float4 per lane (rows and leading dimensions
// 16-byte aligned, checked by the host).
template <int V>
__global__ void __launch_bounds__(256)
symce_unnorm_vec_kernel(SymOperands src, int B, int F, int voff, const float* __restrict__ gt_parts, int n_splits,
                        int64_t split_stride, const float* __restrict__ gg, SymGrads out) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  constexpr int D = 128 * V;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int NG = (voff + F) * B;
  if (r >= B + NG) return;
  const float4* x = reinterpret_cast<const float4*>(symce_src_row(src, B, voff, D, r)) + lane;
  float* dst;
  const float* grow;
  int ns = 1;
  if (r < B) {
    dst = out.text ? out.text + int64_t(r) * out.ldt : nullptr;
    grow = gt_parts + int64_t(r) * D;
    ns = n_splits;
  } else {
    const int g = r - B, fp = g / B, j = g - fp * B;
    dst = (fp < voff) ? (out.video ? out.video + int64_t(j) * out.ldv : nullptr)
                      : (out.frames ? out.frames + int64_t(j) * out.ldf + int64_t(fp - voff) * D : nullptr);
    grow = gg + int64_t(g) * D;
  }
  if (dst == nullptr) return;
  float4 xv[V], gv[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    xv[k] = __ldg(x + 32 * k);
    gv[k] = __ldg(reinterpret_cast<const float4*>(grow) + lane + 32 * k);
  }
  for (int s0 = 1; s0 < ns; ++s0) {                     // split-K partials in split order
    const float4* gs = reinterpret_cast<const float4*>(grow + int64_t(s0) * split_stride) + lane;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float4 t = __ldg(gs + 32 * k);
      gv[k].x += t.x; gv[k].y += t.y; gv[k].z += t.z; gv[k].w += t.w;
    }
  }
  float ss = 0.f, xg = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    ss = fmaf(xv[k].x, xv[k].x, ss); ss = fmaf(xv[k].y, xv[k].y, ss);
    ss = fmaf(xv[k].z, xv[k].z, ss); ss = fmaf(xv[k].w, xv[k].w, ss);
    xg = fmaf(xv[k].x, gv[k].x, xg); xg = fmaf(xv[k].y, gv[k].y, xg);
    xg = fmaf(xv[k].z, gv[k].z, xg); xg = fmaf(xv[k].w, gv[k].w, xg);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    xg += __shfl_xor_sync(0xffffffffu, xg, o);
  }
  const float n = sqrtf(ss);
  const float proj = xg / (n * n);
  float4* o4 = reinterpret_cast<float4*>(dst) + lane;
#pragma unroll
  for (int k = 0; k < V; ++k)
    o4[32 * k] = make_float4((gv[k].x - xv[k].x * proj) / n, (gv[k].y - xv[k].y * proj) / n,
                             (gv[k].z - xv[k].z * proj) / n, (gv[k].w - xv[k].w * proj) / n);
}

}  // namespace hmmc

using namespace hmmc;

extern "C" {

size_t hmmc_similarity_workspace_bytes(int64_t Bt, int64_t Bv, int Fv, int D, int prec) {
  Workspace ws(nullptr, 0);
  const int planes = planes_of(prec);
  const int64_t N = Bv * Fv;
  if (prec == HMMC_PREC_FP32) {
    ws.take<float>(size_t(Bt) * D);
    ws.take<float>(size_t(N) * D);
  } else {
    ws.take<__nv_bfloat16>(size_t(Bt) * planes * D);
    ws.take<__nv_bfloat16>(size_t(N) * planes * D);
  }
  // backward: normalised copies + two gradient buffers
  Workspace wb(nullptr, 0);
  wb.take<float>(size_t(Bt) * D);
  wb.take<float>(size_t(N) * D);
  wb.take<float>(size_t(Bt) * D);
  wb.take<float>(size_t(N) * D);
  return (ws.used > wb.used ? ws.used : wb.used) + 256;
}

int hmmc_loose_similarity_fwd(const float* seq, int64_t Bt, const float* vis, int64_t Bv, int Fv, int D, float scale,
                              int prec, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_REQUIRE(seq && vis && out && Bt > 0 && Bv > 0 && Fv > 0 && D > 0, "loose_similarity: bad arguments");
  HMMC_REQUIRE(prec >= 0 && prec <= 2, "loose_similarity: unknown precision %d", prec);
  const int64_t N = Bv * Fv;
  HMMC_REQUIRE(N < (int64_t(1) << 31) && Bt < (int64_t(1) << 31), "loose_similarity: tile too large");
  Workspace ws(workspace, workspace_bytes);
  int rc;
  if (prec == HMMC_PREC_FP32) {
    float* sh = ws.take<float>(size_t(Bt) * D);
    float* vh = ws.take<float>(size_t(N) * D);
    if (!ws.ok()) { set_error("loose_similarity: workspace too small (%zu > %zu)", ws.used, workspace_bytes); return HMMC_ERR_WORKSPACE; }
    if ((rc = rownorm_pack(seq, Bt, D, D, 0.f, 1, sh, nullptr, nullptr, 0, st))) return rc;
    if ((rc = rownorm_pack(vis, N, D, D, 0.f, 1, vh, nullptr, nullptr, 0, st))) return rc;
    return gemm_f32(sh, D, 1, vh, D, 1, out, N, int(Bt), int(N), D, scale, st);
  }
  HMMC_REQUIRE(D % 64 == 0, "loose_similarity: tensor-core path needs D %% 64 == 0 (D=%d)", D);
  const int planes = planes_of(prec);
  __nv_bfloat16* sp = ws.take<__nv_bfloat16>(size_t(Bt) * planes * D);
  __nv_bfloat16* vp = ws.take<__nv_bfloat16>(size_t(N) * planes * D);
  if (!ws.ok()) { set_error("loose_similarity: workspace too small (%zu > %zu)", ws.used, workspace_bytes); return HMMC_ERR_WORKSPACE; }
  if ((rc = rownorm_pack(seq, Bt, D, D, 0.f, planes, nullptr, nullptr, sp, int64_t(planes) * D, st))) return rc;
  if ((rc = rownorm_pack(vis, N, D, D, 0.f, planes, nullptr, nullptr, vp, int64_t(planes) * D, st))) return rc;
  return umma_gemm_store(sp, int64_t(planes) * D, vp, int64_t(planes) * D, out, N, 0, int(Bt), int(N), D, planes, 1,
                         scale, st);
}

int hmmc_loose_similarity_bwd(const float* seq, int64_t Bt, const float* vis, int64_t Bv, int Fv, int D, float scale,
                              const float* dout, float* dseq, float* dvis, void* workspace, size_t workspace_bytes,
                              void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_REQUIRE(seq && vis && dout && Bt > 0 && Bv > 0 && Fv > 0 && D > 0, "loose_similarity_bwd: bad arguments");
  const int64_t N = Bv * Fv;
  Workspace ws(workspace, workspace_bytes);
  float* sh = ws.take<float>(size_t(Bt) * D);
  float* vh = ws.take<float>(size_t(N) * D);
  float* gs = ws.take<float>(size_t(Bt) * D);
  float* gv = ws.take<float>(size_t(N) * D);
  if (!ws.ok()) { set_error("loose_similarity_bwd: workspace too small (%zu > %zu)", ws.used, workspace_bytes); return HMMC_ERR_WORKSPACE; }
  int rc;
  if ((rc = rownorm_pack(seq, Bt, D, D, 0.f, 1, sh, nullptr, nullptr, 0, st))) return rc;
  if ((rc = rownorm_pack(vis, N, D, D, 0.f, 1, vh, nullptr, nullptr, 0, st))) return rc;
  if (dseq != nullptr) {
    // g_shat[i,d] = scale * sum_j dout[i,j] vhat[j,d]
    if ((rc = gemm_f32(dout, N, 1, vh, 1, D, gs, D, int(Bt), D, int(N), scale, st))) return rc;
    unnormalize_grad_kernel<<<unsigned((Bt + 7) / 8), 256, 0, st>>>(seq, gs, dseq, Bt, D);
    HMMC_CHECK_LAUNCH();
  }
  if (dvis != nullptr) {
    // g_vhat[j,d] = scale * sum_i dout[i,j] shat[i,d]
    if ((rc = gemm_f32(dout, 1, N, sh, 1, D, gv, D, int(N), D, int(Bt), scale, st))) return rc;
    unnormalize_grad_kernel<<<unsigned((N + 7) / 8), 256, 0, st>>>(vis, gv, dvis, N, D);
    HMMC_CHECK_LAUNCH();
  }
  return HMMC_OK;
}

int hmmc_cross_en_fwd_bwd(const float* S, int64_t lds, int B, float* loss_out, float* dS, int64_t ldds,
                          float* row_scratch, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_REQUIRE(S && loss_out && row_scratch && B > 0 && lds >= B, "cross_en: bad arguments");
  cross_en_row_kernel<<<B, 256, 0, st>>>(S, lds, B, row_scratch, dS, ldds);
  HMMC_CHECK_LAUNCH();
  sum_to_scalar_kernel<<<1, 256, 0, st>>>(row_scratch, B, loss_out, 0);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

struct SymCeWs {
  float *that, *ghat, *S, *lse_row, *lse_col, *row_loss, *gt, *gg;
  SymLossAcc* counter;
  // tensor-core path: plane-packed operands (straight and transposed) and split-K partials
  __nv_bfloat16 *Tp, *TTp, *Gp, *GTp, *Sp, *STp;
  float* parts;
  int splits;
};
constexpr int SYMCE_MAX_SPLITS = 32;
// split-K cap of the text-gradient GEMM: more slices shorten the GEMM but lengthen the un-normalise pass that adds
// them (B=256 raw call: 83.4 us at 16, 78.6 at 8, 80.7 at 4)
constexpr int SYMCE_SPLIT_CAP = 8;
static bool symce_tensor_ok(int B, int D, int prec) {
  // bf16x3 (fp32-parity) takes any multiple of 32 rows; the single-plane bf16 mode keeps its former
  // 64-row granularity and leaves smaller batches on the exact CUDA-core path
  if (prec == HMMC_PREC_FP32 || D % 64 != 0 || D > 1024) return false;
  return prec == HMMC_PREC_BF16X3 ? (B % 32 == 0) : (B % 64 == 0);
}
static void symce_carve(Workspace& ws, SymCeWs& w, int B, int F, int D, int prec) {
  const size_t NB = size_t(1 + F);
  w.that = ws.take<float>(size_t(B) * D);
  w.ghat = ws.take<float>(NB * B * D);
  w.S = ws.take<float>(size_t(B) * NB * B);
  w.lse_row = ws.take<float>(NB * B);
  w.lse_col = ws.take<float>(NB * B);
  w.row_loss = ws.take<float>(size_t(B));
  w.counter = ws.take<SymLossAcc>(1);
  w.gt = ws.take<float>(size_t(B) * D);
  w.gg = ws.take<float>(NB * B * D);
  w.Tp = w.TTp = w.Gp = w.GTp = w.Sp = w.STp = nullptr;
  w.parts = nullptr;
  w.splits = 1;
  if (symce_tensor_ok(B, D, prec)) {
    // contraction extents of the backward GEMMs, padded with zero columns to the 64-wide k-block
    const size_t P = size_t(planes_of(prec));
    const size_t Bk = align_up(size_t(B), 64), NGk = align_up(NB * B, 64);
    w.Tp = ws.take<__nv_bfloat16>(size_t(B) * P * D);
    w.TTp = ws.take<__nv_bfloat16>(size_t(D) * P * Bk);
    w.Gp = ws.take<__nv_bfloat16>(NB * B * P * D);
    w.GTp = ws.take<__nv_bfloat16>(size_t(D) * P * NGk);
    w.Sp = ws.take<__nv_bfloat16>(size_t(B) * P * NGk);
    w.STp = ws.take<__nv_bfloat16>(NB * B * P * Bk);
    w.parts = ws.take<float>(size_t(SYMCE_MAX_SPLITS) * B * D);
  }
}

size_t hmmc_sym_ce_workspace_bytes(int B, int F, int D, int prec) {
  Workspace ws(nullptr, 0);
  SymCeWs w;
  symce_carve(ws, w, B, F, D, prec);
  return ws.used + 256;
}

static int sym_ce_impl(const SymOperands& src, int B, int F, int D, float scale, float w_vtm, float w_ftm, int prec,
                       float* loss_out, const SymGrads& grads, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
  const float *text = src.text, *video = src.video, *frames = src.frames;
  float *dtext = grads.text, *dvideo = grads.video, *dframes = grads.frames;
  HMMC_REQUIRE(text && loss_out && B > 0 && D > 0 && F >= 0, "sym_ce: bad arguments");
  const int voff = (video != nullptr) ? 1 : 0;   // video == NULL: frame_loss alone (modules/modeling.py:665-673)
  HMMC_REQUIRE(voff + F > 0, "sym_ce: neither video nor frames given");
  HMMC_REQUIRE(voff == 1 || dvideo == nullptr, "sym_ce: dvideo requested without video");
  HMMC_REQUIRE(F == 0 || frames != nullptr, "sym_ce: frames is null but F=%d", F);
  HMMC_REQUIRE(prec >= 0 && prec <= 2, "sym_ce: unknown precision %d", prec);
  const int NB = voff + F;
  Workspace ws(workspace, workspace_bytes);
  SymCeWs w;
  symce_carve(ws, w, B, F, D, prec);
  if (!ws.ok()) { set_error("sym_ce: workspace too small (%zu > %zu)", ws.used, workspace_bytes); return HMMC_ERR_WORKSPACE; }
  const bool need_grad = dtext || dvideo || dframes;
  int rc;
  if (symce_tensor_ok(B, D, prec)) {
    // tensor-core path: 6 launches (3 without gradients)
    HMMC_REQUIRE(D <= 1024, "sym_ce: D=%d above the tensor-core path's limit of 1024", D);
    const int NG = NB * B, P = planes_of(prec);
    const int Bk = int(align_up(size_t(B), 64)), NGk = int(align_up(size_t(NG), 64));
    const float wf = (F > 0) ? w_ftm / float(F) : 0.f;
    if (need_grad && Bk != B) {
      HMMC_CHECK_CUDA(cudaMemsetAsync(w.TTp, 0, size_t(D) * P * Bk * sizeof(__nv_bfloat16), st));
      HMMC_CHECK_CUDA(cudaMemsetAsync(w.STp, 0, size_t(NG) * P * Bk * sizeof(__nv_bfloat16), st));
    }
    if (need_grad && NGk != NG) {
      HMMC_CHECK_CUDA(cudaMemsetAsync(w.GTp, 0, size_t(D) * P * NGk * sizeof(__nv_bfloat16), st));
      HMMC_CHECK_CUDA(cudaMemsetAsync(w.Sp, 0, size_t(B) * P * NGk * sizeof(__nv_bfloat16), st));
    }
    const size_t prep_smem = size_t(32) * (D + 1) * sizeof(float);
    static int prep_smem_set = 0;
    if (prep_smem > 48 * 1024 && prep_smem_set < int(prep_smem)) {
      HMMC_CHECK_CUDA(cudaFuncSetAttribute(symce_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(prep_smem)));
      prep_smem_set = int(prep_smem);
    }
    // every launch of the chain is programmatically serialised: a kernel's blocks become resident (and run their
    // prologue) while the previous kernel drains, and wait for its results with griddepcontrol.wait
    count_launch();
    HMMC_CHECK_CUDA(launch_pdl(symce_prep_kernel, dim3(unsigned((B + NG) / 32)), dim3(1024), prep_smem, st, src, B, F, voff,
                               D, P, Bk, NGk, w.Tp, need_grad ? w.TTp : nullptr, w.Gp, need_grad ? w.GTp : nullptr,
                               w.counter));
    // 128-wide tiles: at these sizes the GEMMs are a fraction of a wave, narrower tiles put twice the SMs to work
    if ((rc = umma_gemm_store(w.Tp, int64_t(P) * D, w.Gp, int64_t(P) * D, w.S, NG, 0, B, NG, D, P, 1, scale, st,
                              (int64_t(B) * NG <= int64_t(148) * 128 * 256) ? 128 : 0))) return rc;
    const int row_blocks = (B * NB + 7) / 8;
    count_launch();
    HMMC_CHECK_CUDA(launch_pdl(symce_lse_kernel, dim3(unsigned(row_blocks + (NG + 31) / 32)), dim3(256), 0, st, w.S, B, NB,
                               row_blocks, w.lse_row, w.lse_col, voff, w_vtm, wf, loss_out, w.counter));
    if (!need_grad) return HMMC_OK;
    count_launch();
    HMMC_CHECK_CUDA(launch_pdl(symce_gradpack_kernel, dim3(NG / 32, B / 32), dim3(256), 0, st, w.S, B, NB, w.lse_row,
                               w.lse_col, voff, w_vtm, wf, P, Bk, NGk, w.Sp, w.STp));
    // backward contractions in one grouped launch:
    //   g_that[i,d] = scale * sum_g G[i,g] ghat[g,d]   long K (= NG), few output tiles: split-K partials
    //   g_ghat[g,d] = scale * sum_i G[i,g] that[i,d]
    // the split count fills the SMs the g_ghat tiles leave idle
    const int tiles = ((B + 127) / 128) * ((D + 255) / 256);
    const int other = (dvideo != nullptr || dframes != nullptr) ? ((NG + 127) / 128) * ((D + 255) / 256) : 0;
    int sp = (sm_count() - other) / (tiles > 0 ? tiles : 1);
    if (sp < 1) sp = 1;
    if (sp > SYMCE_SPLIT_CAP) sp = SYMCE_SPLIT_CAP;
    const int eff = umma_effective_splits(NGk, P, sp);
    StoreGemm gm[2];
    int ng = 0;
    if (dtext != nullptr)
      gm[ng++] = StoreGemm{w.Sp, int64_t(P) * NGk, w.GTp, int64_t(P) * NGk, w.parts, D, int64_t(B) * D, B, D, NGk, sp};
    if (dvideo != nullptr || dframes != nullptr)
      gm[ng++] = StoreGemm{w.STp, int64_t(P) * Bk, w.TTp, int64_t(P) * Bk, w.gg, D, 0, NG, D, Bk, 1};
    if ((rc = umma_gemm_store_grouped(gm, ng, P, scale, st))) return rc;
    count_launch();
    auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec = a16(src.text) && a16(src.video) && a16(src.frames) && a16(grads.text) && a16(grads.video) &&
                     a16(grads.frames) && src.ldt % 4 == 0 && src.ldv % 4 == 0 && src.ldf % 4 == 0 &&
                     grads.ldt % 4 == 0 && grads.ldv % 4 == 0 && grads.ldf % 4 == 0;
    const dim3 ugrid(unsigned((B + NG + 7) / 8));
    if (vec && D == 512)
      HMMC_CHECK_CUDA(launch_pdl(symce_unnorm_vec_kernel<4>, ugrid, dim3(256), 0, st, src, B, F, voff, w.parts, eff,
                                 int64_t(B) * D, w.gg, grads));
    else if (vec && D == 256)
      HMMC_CHECK_CUDA(launch_pdl(symce_unnorm_vec_kernel<2>, ugrid, dim3(256), 0, st, src, B, F, voff, w.parts, eff,
                                 int64_t(B) * D, w.gg, grads));
    else if (D <= 512)
      HMMC_CHECK_CUDA(launch_pdl(symce_unnorm_kernel<16>, dim3(unsigned((B + NG + 7) / 8)), dim3(256), 0, st, src, B, F, voff,
                                 D, w.parts, eff, int64_t(B) * D, w.gg, grads));
    else
      HMMC_CHECK_CUDA(launch_pdl(symce_unnorm_kernel<32>, dim3(unsigned((B + NG + 7) / 8)), dim3(256), 0, st, src, B, F, voff,
                                 D, w.parts, eff, int64_t(B) * D, w.gg, grads));
    return HMMC_OK;
  }
  HMMC_REQUIRE(src.ldt == D && (video == nullptr || src.ldv == D) && (F == 0 || src.ldf == int64_t(F) * D) &&
                   (dtext == nullptr || grads.ldt == D) && (dvideo == nullptr || grads.ldv == D) &&
                   (dframes == nullptr || grads.ldf == int64_t(F) * D),
               "sym_ce: the CUDA-core path needs contiguous operands");
  if ((rc = rownorm_pack(text, B, D, D, 0.f, 1, w.that, nullptr, nullptr, 0, st))) return rc;
  gallery_norm_kernel<<<unsigned((int64_t(NB) * B + 7) / 8), 256, 0, st>>>(video, frames, B, F, voff, D, w.ghat);
  HMMC_CHECK_LAUNCH();
  // CUDA-core path (fp32 mode, or shapes the tensor-core tiling does not take)
  const int NG = NB * B;
  // S_all = scale * that . ghat^T   [B, NG]
  if ((rc = gemm_f32(w.that, D, 1, w.ghat, D, 1, w.S, NG, B, NG, D, scale, st))) return rc;
  symce_row_lse_kernel<<<B, 256, 0, st>>>(w.S, B, NB, w.lse_row);
  HMMC_CHECK_LAUNCH();
  symce_col_lse_kernel<<<unsigned((NG + 31) / 32), 256, 0, st>>>(w.S, B, NB, w.lse_col);
  HMMC_CHECK_LAUNCH();
  const float wf = (F > 0) ? w_ftm / float(F) : 0.f;
  symce_grad_kernel<<<B, 256, 0, st>>>(w.S, B, NB, w.lse_row, w.lse_col, voff, w_vtm, wf, w.row_loss, need_grad ? 1 : 0);
  HMMC_CHECK_LAUNCH();
  sum_to_scalar_kernel<<<1, 256, 0, st>>>(w.row_loss, B, loss_out, 0);
  HMMC_CHECK_LAUNCH();
  if (!need_grad) return HMMC_OK;
  if (dtext != nullptr) {
    // g_that[i,d] = scale * sum_g G[i,g] ghat[g,d]
    if ((rc = gemm_f32(w.S, NG, 1, w.ghat, 1, D, w.gt, D, B, D, NG, scale, st))) return rc;
    unnormalize_grad_kernel<<<unsigned((B + 7) / 8), 256, 0, st>>>(text, w.gt, dtext, B, D);
    HMMC_CHECK_LAUNCH();
  }
  if (dvideo != nullptr || dframes != nullptr) {
    // g_ghat[g,d] = scale * sum_i G[i,g] that[i,d]
    if ((rc = gemm_f32(w.S, 1, NG, w.that, 1, D, w.gg, D, NG, D, B, scale, st))) return rc;
    if (dvideo != nullptr) {
      unnormalize_grad_kernel<<<unsigned((B + 7) / 8), 256, 0, st>>>(video, w.gg, dvideo, B, D);
      HMMC_CHECK_LAUNCH();
    }
    if (dframes != nullptr && F > 0) {
      // gather gallery-ordered g_hat rows into [B,F,D] order (into ghat, no longer needed), then un-normalise
      gallery_scatter_kernel<<<unsigned(int64_t(NB) * B), 128, 0, st>>>(w.gg, B, F, voff, D, /*dvideo=*/w.ghat, /*dframes=*/w.ghat + int64_t(B) * D);
      HMMC_CHECK_LAUNCH();
      unnormalize_grad_kernel<<<unsigned((int64_t(B) * F + 7) / 8), 256, 0, st>>>(frames, w.ghat + int64_t(B) * D, dframes, int64_t(B) * F, D);
      HMMC_CHECK_LAUNCH();
    }
  }
  return HMMC_OK;
}


int hmmc_sym_ce_fwd_bwd(const float* text, const float* video, const float* frames, int B, int F, int D, float scale,
                        float w_vtm, float w_ftm, int prec, float* loss_out, float* dtext, float* dvideo,
                        float* dframes, void* workspace, size_t workspace_bytes, void* stream) {
  const SymOperands src{text, D, video, D, frames, int64_t(F) * D};
  const SymGrads grads{dtext, D, dvideo, D, dframes, int64_t(F) * D};
  return sym_ce_impl(src, B, F, D, scale, w_vtm, w_ftm, prec, loss_out, grads, workspace, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
}

size_t hmmc_sym_ce_packed_workspace_bytes(int B, int F, int D, int prec) {
  // + contiguous copies of the operands and of their gradients for the CUDA-core path
  return hmmc_sym_ce_workspace_bytes(B, F, D, prec) + 2 * (size_t(B) * (2 + F) * D * sizeof(float) + 512);
}

int hmmc_sym_ce_packed_fwd_bwd(const float* packed, int B, int F, int D, float scale, float w_vtm, float w_ftm,
                               int prec, float* loss_out, float* dpacked, void* workspace, size_t workspace_bytes,
                               void* stream) {
  HMMC_REQUIRE(packed && loss_out && B > 0 && D > 0 && F >= 0, "sym_ce_packed: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t ld = int64_t(2 + F) * D;
  if (symce_tensor_ok(B, D, prec)) {
    const SymOperands src{packed, ld, packed + D, ld, packed + 2 * D, ld};
    const SymGrads grads{dpacked, ld, dpacked ? dpacked + D : nullptr, ld, dpacked ? dpacked + 2 * D : nullptr, ld};
    return sym_ce_impl(src, B, F, D, scale, w_vtm, w_ftm, prec, loss_out, grads, workspace, workspace_bytes, st);
  }
  // CUDA-core path: un-pack into contiguous operands, run, re-pack the gradients
  Workspace ws(workspace, workspace_bytes);
  float* ops3 = ws.take<float>(size_t(B) * ld);
  float* grd3 = ws.take<float>(size_t(B) * ld);
  HMMC_REQUIRE(ws.ok(), "sym_ce_packed: workspace too small (%zu needed, %zu given)", ws.used, workspace_bytes);
  float* t = ops3;
  float* v = t + int64_t(B) * D;
  float* f = v + int64_t(B) * D;
  const int32_t widths[3] = {D, D, F * D};
  const int n = (F > 0) ? 3 : 2;
  const uint64_t src_ptrs[3] = {reinterpret_cast<uint64_t>(t), reinterpret_cast<uint64_t>(v), reinterpret_cast<uint64_t>(f)};
  int rc;
  if ((rc = hmmc_unpack_rows(packed, src_ptrs, widths, n, B, stream))) return rc;
  float* dt = dpacked ? grd3 : nullptr;
  float* dv = dpacked ? dt + int64_t(B) * D : nullptr;
  float* df = (dpacked && F > 0) ? dv + int64_t(B) * D : nullptr;
  char* rest = static_cast<char*>(workspace) + align_up(ws.used, 256);
  const SymOperands src{t, D, v, D, F > 0 ? f : nullptr, int64_t(F) * D};
  const SymGrads grads{dt, D, dv, D, df, int64_t(F) * D};
  if ((rc = sym_ce_impl(src, B, F, D, scale, w_vtm, w_ftm, prec, loss_out, grads, rest,
                        workspace_bytes - align_up(ws.used, 256), st))) return rc;
  if (dpacked != nullptr) {
    const uint64_t g_ptrs[3] = {reinterpret_cast<uint64_t>(dt), reinterpret_cast<uint64_t>(dv), reinterpret_cast<uint64_t>(df)};
    if ((rc = hmmc_pack_rows(g_ptrs, widths, n, B, dpacked, nullptr, 0, stream))) return rc;
  }
  return HMMC_OK;
}

}  // extern "C"

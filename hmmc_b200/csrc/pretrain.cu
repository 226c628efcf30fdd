// Pre-train (MoCo) head: InfoNCE against the negative queues, queue packing, enqueue,
// momentum EMA and the pack/unpack helpers of the key all-gather.
#include "common.cuh"
#include "umma_gemm.cuh"
#include <map>
#include <mutex>
#include <tuple>

namespace hmmc {

// ------------------------------------------------------------------ queue packing
// dk [D,Kq] fp32  ->  pack_kd [Kq, planes*D] and pack_dk [D, planes*Kq] (bf16 hi / lo planes).
// 32x32 tile transpose through shared memory so both global sides stay coalesced.
__global__ void queue_pack_kernel(const float* __restrict__ dk, __nv_bfloat16* __restrict__ pack_kd,
                                  __nv_bfloat16* __restrict__ pack_dk, int D, int Kq, int planes) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int d = d0 + r, j = j0 + tx;
    float v = (d < D && j < Kq) ? dk[int64_t(d) * Kq + j] : 0.f;
    tile[r][tx] = v;
    if (pack_dk != nullptr && d < D && j < Kq) {
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      pack_dk[int64_t(d) * planes * Kq + j] = hi;
      if (planes == 2) pack_dk[int64_t(d) * planes * Kq + Kq + j] = lo;
    }
  }
  __syncthreads();
  if (pack_kd != nullptr) {
    for (int r = ty; r < 32; r += 8) {
      const int j = j0 + r, d = d0 + tx;
      if (j < Kq && d < D) {
        __nv_bfloat16 hi, lo;
        split_bf16(tile[tx][r], hi, lo);
        pack_kd[int64_t(j) * planes * D + d] = hi;
        if (planes == 2) pack_kd[int64_t(j) * planes * D + D + d] = lo;
      }
    }
  }
}

int pack_dual(const float* src, int rows, int cols, int planes, void* straight, void* transposed, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return HMMC_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  // queue_pack_kernel with dk = src (D = rows, Kq = cols): pack_dk = straight, pack_kd = transposed
  queue_pack_kernel<<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(transposed),
                                          static_cast<__nv_bfloat16*>(straight), rows, cols, planes);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

// ------------------------------------------------------------------ FP32 path helpers
// E = exp(S/T - c) in place, row sums -> rowsum[r]
__global__ void exp_rowsum_kernel(float* __restrict__ S, int64_t lds, int Kq, float invT, float c,
                                  float* __restrict__ rowsum) {
  __shared__ float red[32];
  float* row = S + int64_t(blockIdx.x) * lds;
  float acc = 0.f;
  for (int j = threadIdx.x; j < Kq; j += blockDim.x) {
    const float e = expf(row[j] * invT - c);
    row[j] = e;
    acc += e;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) rowsum[blockIdx.x] = acc;
}

// ------------------------------------------------------------------ prep: normalise + pack several q tensors
constexpr int MAX_GROUPS = 4;   // distinct query tensors of one fused call
constexpr int MAX_BLOCKS = 6;   // (query tensor, queue) pairs = GEMM problems

struct LossAcc {
  unsigned long long sum[3];      // loss slots (FAM, VTM, FTM) in units of 2^-40
  unsigned arrived;
  unsigned pad;
};
struct PrepArgs {
  const float* x[MAX_GROUPS];
  float* xhat[MAX_GROUPS];              // fp32 normalised copy (FP32 path) or nullptr
  __nv_bfloat16* packed[MAX_GROUPS];    // bf16 planes (tensor-core path) or nullptr
  int row_begin[MAX_GROUPS + 1];
  int n;
};

// One warp per row, all query tensors of the call in one launch (F.normalize, eps 1e-12).
// The bf16 operand copy carries the factor pack_scale = log2(e)/T, so that the S-GEMM's accumulators are
// logits in base-2 units (see EpiInfoNCE); the fp32 copy stays the plain unit vector.
// Block 0 also clears the finish kernel's loss accumulators.
template <int V>   // float4 per lane (D == 128 * V), 0 = any D (scalar accesses)
__global__ void __launch_bounds__(256)
prep_rows_kernel(const __grid_constant__ PrepArgs a, int D, int planes, float pack_scale, LossAcc* loss_acc) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_acc != nullptr) *loss_acc = LossAcc{{0ull, 0ull, 0ull}, 0u, 0u};
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.row_begin[a.n]) return;
  int gi = 0;
#pragma unroll
  for (int i = 1; i < MAX_GROUPS; ++i)
    if (i < a.n && row >= a.row_begin[i]) gi = i;
  const int r = row - a.row_begin[gi];
  const float* xr = a.x[gi] + int64_t(r) * D;
  float* xh = a.xhat[gi];
  __nv_bfloat16* pk = a.packed[gi];
  if (V > 0) {
    float4 v[V > 0 ? V : 1];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = __ldg(reinterpret_cast<const float4*>(xr) + lane + 32 * i);
      ss = fmaf(v[i].x, v[i].x, ss); ss = fmaf(v[i].y, v[i].y, ss);
      ss = fmaf(v[i].z, v[i].z, ss); ss = fmaf(v[i].w, v[i].w, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);     // x * (1/n): within 1 ulp of F.normalize's x / n
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 h = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
      const int d = 4 * (lane + 32 * i);
      if (xh != nullptr) *reinterpret_cast<float4*>(xh + int64_t(r) * D + d) = h;
      if (pk != nullptr) {
        const float s[4] = {h.x * pack_scale, h.y * pack_scale, h.z * pack_scale, h.w * pack_scale};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(s[j], hi[j], lo[j]);
        __nv_bfloat16* o = pk + int64_t(r) * planes * D + d;
        *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
        if (planes == 2) *reinterpret_cast<uint2*>(o + D) = *reinterpret_cast<const uint2*>(lo);
      }
    }
  } else {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = xr[d]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int d = lane; d < D; d += 32) {
      const float v = xr[d] * inv;
      if (xh != nullptr) xh[int64_t(r) * D + d] = v;
      if (pk != nullptr) {
        __nv_bfloat16 hi, lo;
        split_bf16(v * pack_scale, hi, lo);
        pk[int64_t(r) * planes * D + d] = hi;
        if (planes == 2) pk[int64_t(r) * planes * D + D + d] = lo;
      }
    }
  }
}

// ------------------------------------------------------------------ finish kernel
// One WARP per query row r = n*Fq + f of a query tensor.  For every (queue) contribution of
// that tensor it consumes the negatives' row sum S_r and U_r = sum_j e_rj Q_j, evaluates the
// positive terms selected by pos_mode, and writes the row's loss shares and dL/dq_r
// (SURVEY.md 8a').  A tensor that is the query of two losses (title_fea: VTM and FTM) gets
// both contributions here, so its gradient is written once.
// The last block to finish adds the per-row loss shares in a fixed order and writes the scalars
// (no separate reduction launch; the order does not depend on which block is last).
constexpr int FIN_MAXD = 2048;            // D <= 32 lanes * FIN_MAXE
constexpr int FIN_MAXE = FIN_MAXD / 32;

struct Contribution {
  const float* keys;
  const float* rowsum_part;   // [n_parts, rows]
  const float* U_part;        // [n_splits][rows, D]
  int64_t split_stride;
  int pos_mode, Fk, n_parts, n_splits;
  float coef;                 // weight / b
  int kind;                   // loss slot (0 FAM, 1 VTM, 2 FTM)
};
struct RowGroup {
  const float* q;
  float* dq;                  // nullptr: forward only
  int rows, Fq, ncontrib;
  Contribution c[2];
};
struct FinishArgs {
  RowGroup g[MAX_GROUPS];
  int row_begin[MAX_GROUPS + 1];
  int n;
  int heavy_rows;             // vector kernel: the first heavy_rows rows get a whole block each (finish_row_block)
};

// Per-kind sums in a fixed order (deterministic) and the final scalars:
//   mode 0: out[0] += fam + vtm + ftm                     (hmmc_infonce_queue_fwd_bwd)
//   mode 1: out[0..3] = total, FAM, VTM, FTM; the slots arrive weighted (w * loss) and are
//           reported unweighted as well               (hmmc_pretrain_head_fwd_bwd)
struct LossFinal {
  float* out;
  int mode;
  float w_fam, w_vtm, w_ftm;
  int use_frame_fea;
};

// Loss sums without a hand-over on the critical path.  Every block adds its three shares to 64-bit fixed-point
// accumulators with fire-and-forget reductions (integer addition commutes: the sums are identical from run to run
// whatever the order) and signals its arrival with a release reduction; nobody waits for an answer, so the block
// retires at once.  (Measured: the former "publish a partial, fence, returning atomic, last block adds" sequence
// kept every block resident for two more round trips, 7 us of a 26 us kernel at b = 256, profiles/r2_finish_kernel.md.)
// The block with the highest index then waits for the arrival count (all other blocks are resident or done by the
// time it runs, so this cannot deadlock), converts the sums and re-arms the accumulators for the next call.
constexpr float LOSS_FIX_SCALE = 1099511627776.0f;            // 2^40

__device__ __forceinline__ void finish_losses(const float (&loss_kind)[3], LossAcc* acc, const LossFinal& f) {
  __shared__ float wl[3][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;       // nw <= 16
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) wl[k][warp] = loss_kind[k];
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = 0.f;
    for (int w = 0; w < nw; ++w) v += wl[k][w];
    atomicAdd(&acc->sum[k], static_cast<unsigned long long>(__float2ll_rn(v * LOSS_FIX_SCALE)));
  }
  ptx::red_release_add(&acc->arrived, 1u);
  if (blockIdx.x != gridDim.x - 1) return;
  const long long t0 = clock64();
  while (ptx::ld_acquire(&acc->arrived) != gridDim.x)
    if (clock64() - t0 > (1ll << 33)) __trap();            // seconds: a lost arrival is a bug, not a wait
  float kinds[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    kinds[k] = float(double(static_cast<long long>(__ldcg(&acc->sum[k]))) * (1.0 / double(LOSS_FIX_SCALE)));
    acc->sum[k] = 0ull;                     // ready for the next call on this workspace
  }
  acc->arrived = 0u;
  if (f.mode == 0) {
    f.out[0] += kinds[0] + kinds[1] + kinds[2];
  } else {
    const float fam = kinds[0], vtm = kinds[1], ftm = f.use_frame_fea ? kinds[2] : 0.f;
    f.out[0] = fam + vtm + ftm;
    f.out[1] = (f.w_fam != 0.f) ? fam / f.w_fam : 0.f;
    f.out[2] = (f.w_vtm != 0.f) ? vtm / f.w_vtm : 0.f;
    f.out[3] = (f.w_ftm != 0.f) ? ftm / f.w_ftm : 0.f;
  }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}
__device__ __forceinline__ void axpy4(float w, const float4& x, float4& y) {
  y.x = fmaf(w, x.x, y.x); y.y = fmaf(w, x.y, y.y); y.z = fmaf(w, x.z, y.z); y.w = fmaf(w, x.w, y.w);
}

#ifndef HMMC_FIN_OCC
#define HMMC_FIN_OCC 2      // resident blocks per SM the vector finish kernel is compiled for (measured: see profiles/)
#endif
#ifndef HMMC_FIN_WARPS
#define HMMC_FIN_WARPS 4    // warps (= rows) per block of the vector finish kernel (measured: profiles/r2_finish_kernel.md)
#endif
constexpr int FIN_WARPS = HMMC_FIN_WARPS;
#ifndef HMMC_FIN_HEAVY
#define HMMC_FIN_HEAVY 1    // rows with a ONE_TO_FRAMES contribution get a whole block each: 0 never, 1 small grids, 2 always
#endif
// number of positive terms of a contribution and the key row of term t (false: the term does not exist)
__device__ __forceinline__ int contrib_terms(const Contribution& C) {
  return (C.pos_mode == HMMC_POS_FRAME_NEIGHBOUR) ? 2 : (C.pos_mode == HMMC_POS_ONE_TO_FRAMES ? C.Fk : 1);
}
__device__ __forceinline__ bool contrib_key_row(const Contribution& C, int r, int n, int f, int t, int& kr) {
  if (C.pos_mode == HMMC_POS_PAIR) { kr = r; return t < 1; }
  if (C.pos_mode == HMMC_POS_FRAME_NEIGHBOUR) {
    const int fk = (t == 0) ? f + 1 : f - 1;            // pairs (i, i+1) and (i+1, i) of frame_self_loss
    kr = n * C.Fk + fk;
    return t < 2 && fk >= 0 && fk < C.Fk;
  }
  if (C.pos_mode == HMMC_POS_ONE_TO_FRAMES) { kr = n * C.Fk + t; return t < C.Fk; }
  kr = n;
  return t < 1;
}

// One row handled by a whole block.  A row with a ONE_TO_FRAMES contribution walks Fk key rows and several split-K
// partials: with one warp that is 8-10 dependent round trips and those rows set the kernel's duration (SMs 42 %
// idle, profiles/r2_ncu_kernels.md).  Here the positive terms and then the U partials are dealt to the eight warps
// (one or two each, all loads of a phase in flight together); the partial gradients meet in shared memory and are
// added in warp order, so the result does not depend on timing.
template <int V>
__device__ __forceinline__ void finish_row_block(const RowGroup& G, int r, float invT, float cmax, float kexp,
                                                 float (&loss_kind)[3]) {
  constexpr int D = 128 * V;
  __shared__ float4 gs[FIN_WARPS][32 * V];            // the warps' partial gradients
  __shared__ float invz[2][FIN_WARPS];
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = r / G.Fq, f = r - n * G.Fq;
  // terms of all contributions in one list; warp w takes the terms w, w + FIN_WARPS, ... (two key rows in flight)
  const int nt0 = contrib_terms(G.c[0]);
  const int T = nt0 + (G.ncontrib > 1 ? contrib_terms(G.c[1]) : 0);
  float4 qv[V], g[V];
  {
    const float4* qp = reinterpret_cast<const float4*>(G.q + int64_t(r) * D) + lane;
#pragma unroll
    for (int i = 0; i < V; ++i) qv[i] = __ldg(qp + 32 * i);
  }
  float4 kv[2][V];
  bool on[2];
  int cix[2];
  auto load_terms = [&](int t0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int tt = t0 + FIN_WARPS * k;
      cix[k] = (tt >= nt0) ? 1 : 0;
      int kr = 0;
      on[k] = tt < T && contrib_key_row(G.c[cix[k]], r, n, f, tt - (cix[k] ? nt0 : 0), kr);
      const float4* kp = reinterpret_cast<const float4*>(G.c[cix[k]].keys + int64_t(on[k] ? kr : 0) * D) + lane;
#pragma unroll
      for (int i = 0; i < V; ++i) kv[k][i] = on[k] ? __ldg(kp + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_terms(warp);
  // the negatives' sums of both contributions (every warp: a few hundred bytes from L2)
  float S0 = 0.f, S1 = 0.f;
  {
    float acc = 0.f;
    for (int p = lane; p < G.c[0].n_parts; p += 32) acc += __ldg(G.c[0].rowsum_part + int64_t(p) * G.rows + r);
    S0 = warp_sum(acc) * kexp;
    if (G.ncontrib > 1) {
      acc = 0.f;
      for (int p = lane; p < G.c[1].n_parts; p += 32) acc += __ldg(G.c[1].rowsum_part + int64_t(p) * G.rows + r);
      S1 = warp_sum(acc) * kexp;
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    ss = dot4(qv[i], qv[i], ss);
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  ss = warp_sum(ss);
  const float nq_raw = sqrtf(ss);
  const float inq = 1.0f / fmaxf(nq_raw, 1e-12f);
#pragma unroll
  for (int i = 0; i < V; ++i) { qv[i].x *= inq; qv[i].y *= inq; qv[i].z *= inq; qv[i].w *= inq; }   // q_hat
  float sum_invZ0 = 0.f, sum_invZ1 = 0.f;
  for (int t0 = warp; t0 < T; t0 += 2 * FIN_WARPS) {
    if (t0 != warp) load_terms(t0);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (!on[k]) continue;
      const Contribution& C = G.c[cix[k]];
      float kk0 = 0.f, kk1 = 0.f, qk0 = 0.f, qk1 = 0.f;
#pragma unroll
      for (int i = 0; i < V; i += 2) {
        kk0 = dot4(kv[k][i], kv[k][i], kk0);
        qk0 = dot4(kv[k][i], qv[i], qk0);
        if (i + 1 < V) {
          kk1 = dot4(kv[k][i + 1], kv[k][i + 1], kk1);
          qk1 = dot4(kv[k][i + 1], qv[i + 1], qk1);
        }
      }
      float kk = kk0 + kk1, qk = qk0 + qk1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        kk += __shfl_xor_sync(0xffffffffu, kk, o);
        qk += __shfl_xor_sync(0xffffffffu, qk, o);
      }
      const float ink = (kk > 1e-24f) ? rsqrtf(kk) : 1e12f;
      const float lpos = qk * ink * invT;
      const float epos = __expf(lpos - cmax);
      const float Z = epos + (cix[k] ? S1 : S0);
      const float iZ = __fdividef(1.0f, Z);
      loss_kind[C.kind] += C.coef * (__logf(Z) + cmax - lpos);
      if (cix[k]) sum_invZ1 += iZ; else sum_invZ0 += iZ;
      const float w = C.coef * invT * (epos * iZ - 1.0f) * ink;
#pragma unroll
      for (int i = 0; i < V; ++i) axpy4(w, kv[k][i], g[i]);
    }
  }
  if (G.dq == nullptr) return;
  if (lane == 0) { invz[0][warp] = sum_invZ0; invz[1][warp] = sum_invZ1; }
  __syncthreads();
  // U partials of all contributions in one list, dealt from the last warp down (the first warps had two terms)
  {
    float tot0 = 0.f, tot1 = 0.f;
#pragma unroll
    for (int w = 0; w < FIN_WARPS; ++w) { tot0 += invz[0][w]; tot1 += invz[1][w]; }
    const int ns0 = G.c[0].n_splits;
    const int NU = ns0 + (G.ncontrib > 1 ? G.c[1].n_splits : 0);
    for (int u = FIN_WARPS - 1 - warp; u < NU; u += FIN_WARPS) {
      const int ci = (u >= ns0) ? 1 : 0;
      const Contribution& C = G.c[ci];
      const float wu = C.coef * invT * (ci ? tot1 : tot0) * kexp;
      const float4* up = reinterpret_cast<const float4*>(C.U_part + int64_t(u - (ci ? ns0 : 0)) * C.split_stride +
                                                         int64_t(r) * D) + lane;
      float4 t[V];
#pragma unroll
      for (int i = 0; i < V; ++i) t[i] = __ldg(up + 32 * i);
#pragma unroll
      for (int i = 0; i < V; ++i) axpy4(wu, t[i], g[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) gs[warp][32 * i + lane] = g[i];
  __syncthreads();
  // every thread owns D / blockDim columns: add the partials in warp order, project through the normalisation
  constexpr int FIN_COLS = (D + 32 * FIN_WARPS - 1) / (32 * FIN_WARPS);
  const float* gflat = reinterpret_cast<const float*>(&gs[0][0]);
  float gsum[FIN_COLS], qh[FIN_COLS];
  float part = 0.f;
#pragma unroll
  for (int j = 0; j < FIN_COLS; ++j) {
    const int c = threadIdx.x + 32 * FIN_WARPS * j;
    gsum[j] = 0.f;
    qh[j] = 0.f;
    if (c < D) {
      // column c of the row lives in float4 number c / 4 = 32 * i + lane  ->  same slot in every warp's partial
#pragma unroll
      for (int w = 0; w < FIN_WARPS; ++w) gsum[j] += gflat[w * D + c];
      qh[j] = __ldg(G.q + int64_t(r) * D + c) * inq;
      part = fmaf(gsum[j], qh[j], part);
    }
  }
  const float qg = block_sum(part, red);
  const bool clamped = nq_raw < 1e-12f;
#pragma unroll
  for (int j = 0; j < FIN_COLS; ++j) {
    const int c = threadIdx.x + 32 * FIN_WARPS * j;
    if (c < D) G.dq[int64_t(r) * D + c] = clamped ? gsum[j] * inq : (gsum[j] - qh[j] * qg) * inq;
  }
}

// kexp: factor that brings the negatives' sums (row sums and U) to the e^{l - cmax} scale the positives use
// (tensor-core path: the GEMM epilogue stores 2^{l log2 e} without the constant max, kexp = e^{-cmax}; fp32 path: 1).
// Vector version: D == 128 * V, every row is V float4 per lane; all of a row's independent loads (query, two
// key rows, two split-K partials of U) are issued together, 16 warps per SM.
template <int V>
__global__ void __launch_bounds__(32 * FIN_WARPS, (V <= 4 ? HMMC_FIN_OCC * 8 / FIN_WARPS : 1))
infonce_finish_vec_kernel(const __grid_constant__ FinishArgs a, float invT, float cmax, float kexp,
                          const LossFinal fin, LossAcc* acc) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  constexpr int D = 128 * V;
  const int lane = threadIdx.x & 31;
  float loss_kind[3] = {0.f, 0.f, 0.f};
  // Blocks [0, heavy_rows): one row each (finish_row_block).  The others: FIN_WARPS consecutive rows of ONE query
  // tensor (the host checks the row counts), so the tensor, its contributions and every loop bound below depend on
  // blockIdx only: the compiler keeps the control flow on the uniform path, without convergence barriers around
  // the warp shuffles.  What differs between the warps of a block (a frame without a left or right neighbour) is
  // handled with a 0/1 factor instead of a branch.
  const bool whole_block = int(blockIdx.x) < a.heavy_rows;
  const int row0 = whole_block ? int(blockIdx.x) : a.heavy_rows + (int(blockIdx.x) - a.heavy_rows) * FIN_WARPS;
  int gi = 0;
#pragma unroll
  for (int i = 1; i < MAX_GROUPS; ++i)
    if (i < a.n && row0 >= a.row_begin[i]) gi = i;
  const RowGroup& G = a.g[gi];
  if (whole_block) {
    finish_row_block<V>(G, row0 - a.row_begin[gi], invT, cmax, kexp, loss_kind);
  } else {
    const int r = row0 - a.row_begin[gi] + int(threadIdx.x >> 5);
    const int n = r / G.Fq, f = r - n * G.Fq;
    float4 qv[V], g[V];
    {
      const float4* qp = reinterpret_cast<const float4*>(G.q + int64_t(r) * D) + lane;
#pragma unroll
      for (int i = 0; i < V; ++i) qv[i] = __ldg(qp + 32 * i);
    }
    // Everything a contribution needs first is independent of the query: its row-sum partials and its first two
    // key rows are requested together with the query row, one round trip instead of three.
    float sp[3];
    float4 kv[2][V];
    float on[2];                       // 1: the term exists for this row, 0: it does not (its key row is a dummy)
    auto load_keys = [&](const Contribution& C, int t0, int nterm) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        on[k] = 0.f;
        if (t0 + k < nterm) {          // uniform
          int kr;
          on[k] = 1.f;
          if (C.pos_mode == HMMC_POS_PAIR) kr = r;
          else if (C.pos_mode == HMMC_POS_FRAME_NEIGHBOUR) {
            const int fk = (t0 + k == 0) ? f + 1 : f - 1;       // pairs (i, i+1) and (i+1, i) of frame_self_loss
            on[k] = (fk >= 0 && fk < C.Fk) ? 1.f : 0.f;
            kr = n * C.Fk + min(max(fk, 0), C.Fk - 1);
          }
          else if (C.pos_mode == HMMC_POS_ONE_TO_FRAMES) kr = n * C.Fk + t0 + k;
          else kr = n;
          const float4* kp = reinterpret_cast<const float4*>(C.keys + int64_t(kr) * D) + lane;
#pragma unroll
          for (int i = 0; i < V; ++i) kv[k][i] = __ldg(kp + 32 * i);
        }
      }
    };
    auto prefetch = [&](const Contribution& C) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int p = lane + 32 * j;
        sp[j] = (p < C.n_parts) ? __ldg(C.rowsum_part + int64_t(p) * G.rows + r) : 0.f;
      }
      load_keys(C, 0, contrib_terms(C));
    };
    prefetch(G.c[0]);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      ss = dot4(qv[i], qv[i], ss);
      g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ss = warp_sum(ss);
    const float nq_raw = sqrtf(ss);
    const float nq = fmaxf(nq_raw, 1e-12f);
    const float inq = 1.0f / nq;              // x * (1/n): within 1 ulp of F.normalize's x / n
#pragma unroll
    for (int i = 0; i < V; ++i) { qv[i].x *= inq; qv[i].y *= inq; qv[i].z *= inq; qv[i].w *= inq; }   // q_hat

    for (int ci = 0; ci < G.ncontrib; ++ci) {
      const Contribution& C = G.c[ci];
      if (ci > 0) prefetch(C);
      // ---- S_r: the negatives' sum, from the row-sum partials
      float S = (sp[0] + sp[1]) + sp[2];
      for (int p = lane + 96; p < C.n_parts; p += 32) S += __ldg(C.rowsum_part + int64_t(p) * G.rows + r);
      S = warp_sum(S) * kexp;
      const int nterm = contrib_terms(C);
      float loss = 0.f, sum_invZ = 0.f;
      const float scale = C.coef * invT;
      // positive terms, two key rows in flight per iteration (the first pair is already on its way)
      for (int t0 = 0; t0 < nterm; t0 += 2) {
        if (t0 > 0) load_keys(C, t0, nterm);
        const bool two = t0 + 1 < nterm;        // uniform
        // ||k||^2 and q_hat.k of both rows, reduced in one butterfly
        float kk[2] = {0.f, 0.f}, qk[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k == 1 && !two) break;
          float kk1 = 0.f, qk1 = 0.f;
#pragma unroll
          for (int i = 0; i < V; i += 2) {
            kk[k] = dot4(kv[k][i], kv[k][i], kk[k]);
            qk[k] = dot4(kv[k][i], qv[i], qk[k]);
            if (i + 1 < V) {
              kk1 = dot4(kv[k][i + 1], kv[k][i + 1], kk1);
              qk1 = dot4(kv[k][i + 1], qv[i + 1], qk1);
            }
          }
          kk[k] += kk1;
          qk[k] += qk1;
        }
        if (two) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            kk[0] += __shfl_xor_sync(0xffffffffu, kk[0], o);
            qk[0] += __shfl_xor_sync(0xffffffffu, qk[0], o);
            kk[1] += __shfl_xor_sync(0xffffffffu, kk[1], o);
            qk[1] += __shfl_xor_sync(0xffffffffu, qk[1], o);
          }
        } else {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            kk[0] += __shfl_xor_sync(0xffffffffu, kk[0], o);
            qk[0] += __shfl_xor_sync(0xffffffffu, qk[0], o);
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k == 1 && !two) break;
          // fast-math intrinsics (2^-21 relative): the row's loss share and weights are O(1) scalars whose error
          // averages over b*F rows; measured against the float64 oracle in tests/test_gpu_pretrain.py
          const float ink = (kk[k] > 1e-24f) ? rsqrtf(kk[k]) : 1e12f;          // 1 / max(||k||, 1e-12)
          const float lpos = qk[k] * ink * invT;
          const float epos = __expf(lpos - cmax);
          const float Z = epos + S;
          const float iZ = __fdividef(1.0f, Z);
          loss = fmaf(on[k], __logf(Z) + cmax - lpos, loss);
          sum_invZ = fmaf(on[k], iZ, sum_invZ);
          // g_hat += coef/T * (p+ - 1) k_hat_t
          const float w = on[k] * scale * (epos * iZ - 1.0f) * ink;
#pragma unroll
          for (int i = 0; i < V; ++i) axpy4(w, kv[k][i], g[i]);
        }
      }
      loss_kind[C.kind] += C.coef * loss;
      if (G.dq != nullptr) {
        // g_hat += coef/T * (sum_t 1/Z_t) U_r ;  U_r = the split-K partials, streamed in split order, two in flight
        const float wu = scale * sum_invZ * kexp;
        for (int s0 = 0; s0 < C.n_splits; s0 += 2) {
          const float4* up = reinterpret_cast<const float4*>(C.U_part + int64_t(s0) * C.split_stride + int64_t(r) * D) + lane;
          float4 t[2][V];
#pragma unroll
          for (int i = 0; i < V; ++i) t[0][i] = __ldg(up + 32 * i);
          if (s0 + 1 < C.n_splits) {       // uniform
            const float4* up1 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(up) + C.split_stride);
#pragma unroll
            for (int i = 0; i < V; ++i) t[1][i] = __ldg(up1 + 32 * i);
#pragma unroll
            for (int i = 0; i < V; ++i) axpy4(wu, t[0][i], g[i]);
#pragma unroll
            for (int i = 0; i < V; ++i) axpy4(wu, t[1][i], g[i]);
          } else {
#pragma unroll
            for (int i = 0; i < V; ++i) axpy4(wu, t[0][i], g[i]);
          }
        }
      }
    }
    if (G.dq != nullptr) {
      // dq = (g_hat - q_hat (q_hat . g_hat)) / ||q||
      float qg = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) qg = dot4(qv[i], g[i], qg);
      qg = warp_sum(qg);
      // F.normalize clamps: q_hat = q/eps is then linear in q and the projection term drops out
      const float proj = (nq_raw < 1e-12f) ? 0.f : qg;
      float4* op = reinterpret_cast<float4*>(G.dq + int64_t(r) * D) + lane;
#pragma unroll
      for (int i = 0; i < V; ++i)
        op[32 * i] = make_float4((g[i].x - qv[i].x * proj) * inq, (g[i].y - qv[i].y * proj) * inq,
                                 (g[i].z - qv[i].z * proj) * inq, (g[i].w - qv[i].w * proj) * inq);
    }
  }
  finish_losses(loss_kind, acc, fin);
}

// Generic version (any D <= FIN_MAXD, scalar accesses).
template <int NE>   // NE = elements per lane = D / 32 rounded up
__global__ void __launch_bounds__(256, (NE <= 16 ? 2 : 1))
infonce_finish_kernel(const __grid_constant__ FinishArgs a, int D, float invT, float cmax, float kexp,
                      const LossFinal fin, LossAcc* acc) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int total_rows = a.row_begin[a.n];
  float loss_kind[3] = {0.f, 0.f, 0.f};
  if (row < total_rows) {
    int gi = 0;
#pragma unroll
    for (int i = 1; i < MAX_GROUPS; ++i)
      if (i < a.n && row >= a.row_begin[i]) gi = i;
    const RowGroup& G = a.g[gi];
    const int r = row - a.row_begin[gi];
    const int n = r / G.Fq, f = r - n * G.Fq;

    float qv[NE], g[NE];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int d = lane + i * 32;
      qv[i] = (d < D) ? G.q[int64_t(r) * D + d] : 0.f;
      g[i] = 0.f;
      ss = fmaf(qv[i], qv[i], ss);
    }
    ss = warp_sum(ss);
    const float nq_raw = sqrtf(ss);
    const float nq = fmaxf(nq_raw, 1e-12f);
#pragma unroll
    for (int i = 0; i < NE; ++i) qv[i] = qv[i] / nq;   // q_hat

    for (int ci = 0; ci < G.ncontrib; ++ci) {
      const Contribution& C = G.c[ci];
      // S_r: negatives' sum of exp(l - cmax)
      float S = 0.f, S1 = 0.f;
      for (int p = lane; p < C.n_parts; p += 64) {
        const float a0 = C.rowsum_part[int64_t(p) * G.rows + r];
        const float a1 = (p + 32 < C.n_parts) ? C.rowsum_part[int64_t(p + 32) * G.rows + r] : 0.f;
        S += a0;
        S1 += a1;
      }
      S = warp_sum(S + S1) * kexp;
      int nterm, kbase, kstep;
      if (C.pos_mode == HMMC_POS_PAIR) { nterm = 1; kbase = r; kstep = 0; }
      else if (C.pos_mode == HMMC_POS_FRAME_NEIGHBOUR) { nterm = 2; kbase = 0; kstep = 0; }
      else if (C.pos_mode == HMMC_POS_ONE_TO_FRAMES) { nterm = C.Fk; kbase = n * C.Fk; kstep = 1; }
      else { nterm = 1; kbase = n; kstep = 0; }
      float loss = 0.f, sum_invZ = 0.f;
      const float scale = C.coef * invT;
      for (int t = 0; t < nterm; ++t) {
        int kr = kbase + t * kstep;
        if (C.pos_mode == HMMC_POS_FRAME_NEIGHBOUR) {
          const int fk = (t == 0) ? f + 1 : f - 1;      // pairs (i, i+1) and (i+1, i) of frame_self_loss
          if (fk < 0 || fk >= C.Fk) continue;           // warp-uniform
          kr = n * C.Fk + fk;
        }
        float kv[NE];
        float kk = 0.f, qk = 0.f;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          const int d = lane + i * 32;
          kv[i] = (d < D) ? __ldg(C.keys + int64_t(kr) * D + d) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          kk = fmaf(kv[i], kv[i], kk);
          qk = fmaf(kv[i], qv[i], qk);
        }
        kk = warp_sum(kk);
        qk = warp_sum(qk);
        const float nk = fmaxf(sqrtf(kk), 1e-12f);
        const float lpos = (qk / nk) * invT;
        const float epos = expf(lpos - cmax);
        const float Z = epos + S;
        loss += logf(Z) + cmax - lpos;
        sum_invZ += 1.0f / Z;
        // g_hat += coef/T * (p+ - 1) k_hat_t
        const float w = scale * (epos / Z - 1.0f) / nk;
#pragma unroll
        for (int i = 0; i < NE; ++i) g[i] = fmaf(w, kv[i], g[i]);
      }
      loss_kind[C.kind] += C.coef * loss;
      if (G.dq != nullptr) {
        // g_hat += coef/T * (sum_t 1/Z_t) U_r ;  U_r = sum over the split-K partials, in split order
        const float wu = scale * sum_invZ * kexp;
        float us[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) us[i] = 0.f;
        for (int sidx = 0; sidx < C.n_splits; ++sidx) {
          const float* up = C.U_part + int64_t(sidx) * C.split_stride + int64_t(r) * D;
#pragma unroll
          for (int i = 0; i < NE; ++i) {
            const int d = lane + i * 32;
            us[i] += (d < D) ? __ldg(up + d) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < NE; ++i) g[i] = fmaf(wu, us[i], g[i]);
      }
    }
    if (G.dq != nullptr) {
      // dq = (g_hat - q_hat (q_hat . g_hat)) / ||q||
      float qg = 0.f;
#pragma unroll
      for (int i = 0; i < NE; ++i) qg = fmaf(qv[i], g[i], qg);
      qg = warp_sum(qg);
      const bool clamped = nq_raw < 1e-12f;   // F.normalize clamps: q_hat = q/eps is then linear in q
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const int d = lane + i * 32;
        if (d < D) G.dq[int64_t(r) * D + d] = clamped ? g[i] / nq : (g[i] - qv[i] * qg) / nq;
      }
    }
  }
  finish_losses(loss_kind, acc, fin);
}

// ------------------------------------------------------------------ EMA
constexpr int EMA_THREADS = 256;
constexpr int EMA_BLOCK_ELEMS = 8192;

template <typename T>
__device__ __forceinline__ T ema_one(T pk, T p, float m, float omm);
template <>
__device__ __forceinline__ float ema_one<float>(float pk, float p, float m, float omm) {
  return __fadd_rn(__fmul_rn(pk, m), __fmul_rn(p, omm));   // three roundings, no FMA contraction
}
template <>
__device__ __forceinline__ __half ema_one<__half>(__half pk, __half p, float m, float omm) {
  const __half a = __float2half_rn(__fmul_rn(__half2float(pk), m));
  const __half b = __float2half_rn(__fmul_rn(__half2float(p), omm));
  return __float2half_rn(__fadd_rn(__half2float(a), __half2float(b)));
}
template <>
__device__ __forceinline__ __nv_bfloat16 ema_one<__nv_bfloat16>(__nv_bfloat16 pk, __nv_bfloat16 p, float m, float omm) {
  const __nv_bfloat16 a = __float2bfloat16_rn(__fmul_rn(__bfloat162float(pk), m));
  const __nv_bfloat16 b = __float2bfloat16_rn(__fmul_rn(__bfloat162float(p), omm));
  return __float2bfloat16_rn(__fadd_rn(__bfloat162float(a), __bfloat162float(b)));
}

template <typename T>
__device__ __forceinline__ void ema_span(T* __restrict__ pk, const T* __restrict__ p, int64_t begin, int64_t end,
                                         float m, float omm) {
  constexpr int V = 16 / sizeof(T);   // elements per 128-bit access
  const bool aligned = ((reinterpret_cast<uintptr_t>(pk) | reinterpret_cast<uintptr_t>(p)) & 15) == 0 && (begin % V) == 0;
  if (aligned) {
    const int64_t nvec = (end - begin) / V;
    uint4* pk4 = reinterpret_cast<uint4*>(pk + begin);
    const uint4* p4 = reinterpret_cast<const uint4*>(p + begin);
    for (int64_t i = threadIdx.x; i < nvec; i += EMA_THREADS) {
      uint4 a = pk4[i];
      const uint4 c = __ldg(p4 + i);
      T* av = reinterpret_cast<T*>(&a);
      const T* cv = reinterpret_cast<const T*>(&c);
#pragma unroll
      for (int j = 0; j < V; ++j) av[j] = ema_one<T>(av[j], cv[j], m, omm);
      pk4[i] = a;
    }
    for (int64_t i = begin + nvec * V + threadIdx.x; i < end; i += EMA_THREADS) pk[i] = ema_one<T>(pk[i], p[i], m, omm);
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += EMA_THREADS) pk[i] = ema_one<T>(pk[i], p[i], m, omm);
  }
}

__global__ void __launch_bounds__(EMA_THREADS)
ema_multi_kernel(const uint64_t* __restrict__ pk_ptrs, const uint64_t* __restrict__ p_ptrs,
                 const int64_t* __restrict__ numels, const int32_t* __restrict__ dtypes,
                 const int64_t* __restrict__ block_offsets, int n, float m, float omm) {
  // tensor owning this block: last t with block_offsets[t] <= blockIdx.x
  const int64_t blk = blockIdx.x;
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (block_offsets[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  const int t = lo;
  const int64_t begin = (blk - block_offsets[t]) * EMA_BLOCK_ELEMS;
  const int64_t end = min(begin + int64_t(EMA_BLOCK_ELEMS), numels[t]);
  if (begin >= end) return;
  const int dt = dtypes[t];
  if (dt == 0) ema_span<float>(reinterpret_cast<float*>(pk_ptrs[t]), reinterpret_cast<const float*>(p_ptrs[t]), begin, end, m, omm);
  else if (dt == 1) ema_span<__half>(reinterpret_cast<__half*>(pk_ptrs[t]), reinterpret_cast<const __half*>(p_ptrs[t]), begin, end, m, omm);
  else ema_span<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(pk_ptrs[t]), reinterpret_cast<const __nv_bfloat16*>(p_ptrs[t]), begin, end, m, omm);
}

// ------------------------------------------------------------------ enqueue
struct EnqueueArgs {
  float* dk[5];
  __nv_bfloat16* pack_kd[5];
  __nv_bfloat16* pack_dk[5];
  const float* src[5];      // first sample of this queue's keys
  int64_t src_stride[5];    // elements between consecutive samples
  int Kq[5];                // columns of each queue
  int mult[5];              // columns per sample: 1 or F
  int norm_off[5];          // offset of each queue's norms in the scratch array
  int planes;
  // peer exchange (hmmc_peer_push_rows): the keys sit in one of two slots of the receive buffer, chosen on the
  // device by the exchange counter (NULL: no slots)
  const int32_t* slot_epoch;
  int64_t slot_stride;      // elements between the two slots
  int prenormalised;        // the keys arrive as unit vectors (hmmc_pack_rows normalised them): written as they are
};
__device__ __forceinline__ int64_t enqueue_slot_offset(const EnqueueArgs& a) {
  return a.slot_epoch != nullptr ? int64_t((*a.slot_epoch - 1) & 1) * a.slot_stride : 0;
}

constexpr int ENQ_DCHUNK = 128;

// Scalar scatter (odd D, unaligned sources, no norm scratch): block = CB = 32 consecutive queue columns x
// dchunk embedding dims of one queue.  Phase 1: the block's key norms (read from the pre-pass, or computed
// here).  Phase 2: 32(d) x 32(col) tiles go through shared memory so the [D,Kq] layouts are written 32
// columns (128 contiguous bytes) at a time.  The normal case runs enqueue_vec_kernel below.
// ptr >= 0: the host-tracked pointer (block 0 stores new_ptr);  ptr < 0: read the pointer from
// queue_ptr[0] on the device (CUDA-graph replay: no host value can be baked in) and leave the
// advance to advance_ptr_kernel.
// max(||x||, 1e-12) of every key vector of the five queues, one warp per vector;
// norms[qi][c] with c = sample*mult + f, queue qi starting at norm_off[qi]
__global__ void key_norms_kernel(int nsamples, int D, EnqueueArgs a, float* __restrict__ norms,
                                 const int32_t* __restrict__ staged) {
  if (staged != nullptr && *staged == 0) return;      // nothing staged: these keys were enqueued already
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.y;
  const int mult = a.mult[qi];
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= nsamples * mult) return;
  const int smp = c / mult, f = c - smp * mult;
  const float* x = a.src[qi] + enqueue_slot_offset(a) + int64_t(smp) * a.src_stride[qi] + int64_t(f) * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = x[d]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if (lane == 0) norms[a.norm_off[qi] + c] = fmaxf(sqrtf(ss), 1e-12f);
}

template <int CB>
__global__ void __launch_bounds__(256)
enqueue_kernel(int nsamples, int D, int dchunk, EnqueueArgs a, const float* __restrict__ norms,
               int64_t* __restrict__ queue_ptr, int ptr, int new_ptr, const int32_t* __restrict__ staged) {
  if (staged != nullptr && *staged == 0) return;
  __shared__ float inv[CB];            // 1 / max(||x||, 1e-12) of the block's key vectors
  __shared__ int64_t soff[CB];         // element offset of each key vector in its source tensor
  __shared__ float tile[CB][33];
  const int qi = blockIdx.z;
  const int mult = a.mult[qi];
  const int ncols = nsamples * mult;
  const int c0 = blockIdx.x * CB;
  if (ptr >= 0) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) queue_ptr[0] = new_ptr;
  } else {
    ptr = int(queue_ptr[0]);
  }
  if (c0 >= ncols) return;
  const int Kq = a.Kq[qi];
  const int col_base = ptr * mult;       // first destination column
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src = a.src[qi] + enqueue_slot_offset(a);
  for (int cc = threadIdx.x; cc < CB; cc += blockDim.x) {
    const int c = min(c0 + cc, ncols - 1);         // c = local column = sample*mult + f
    const int smp = c / mult, f = c - smp * mult;
    soff[cc] = int64_t(smp) * a.src_stride[qi] + int64_t(f) * D;
  }
  __syncthreads();
  if (a.prenormalised) {
    for (int cc = threadIdx.x; cc < CB; cc += blockDim.x) inv[cc] = (c0 + cc < ncols) ? 1.0f : 0.f;
  } else if (norms != nullptr) {
    for (int cc = threadIdx.x; cc < CB; cc += blockDim.x)
      inv[cc] = (c0 + cc < ncols) ? 1.0f / norms[a.norm_off[qi] + c0 + cc] : 0.f;
  } else {
    for (int cc = warp; cc < CB; cc += 8) {
      float ss = 0.f;
      if (c0 + cc < ncols) {
        const float* x = src + soff[cc];
        for (int d = lane; d < D; d += 32) { const float v = x[d]; ss = fmaf(v, v, ss); }
      }
      ss = warp_sum(ss);
      if (lane == 0) inv[cc] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
  }
  __syncthreads();
  const int planes = a.planes;
  float* dk = a.dk[qi];
  __nv_bfloat16* pkd = a.pack_kd[qi];
  __nv_bfloat16* pdk = a.pack_dk[qi];
  const int dbeg = blockIdx.y * dchunk;
  const int dend = min(dbeg + dchunk, D);
  for (int d0 = dbeg; d0 < dend; d0 += 32) {
    // read: warp w handles columns w, w+8, ..; lane = d.  x * (1/n): within 1 ulp of F.normalize's x / n
#pragma unroll 4
    for (int cc = warp; cc < CB; cc += 8) {
      const int c = c0 + cc, d = d0 + lane;
      float v = 0.f;
      if (c < ncols && d < dend) {
        v = src[soff[cc] + d] * inv[cc];
        if (pkd != nullptr) {
          __nv_bfloat16 hi, lo;
          split_bf16(v, hi, lo);
          const int64_t o = int64_t(col_base + c) * planes * D + d;
          pkd[o] = hi;
          if (planes == 2) pkd[o + D] = lo;
        }
      }
      tile[cc][lane] = v;
    }
    __syncthreads();
    // write transposed: warp w handles d = w, w+8, ..; lanes sweep the CB columns
    for (int dd = warp; dd < 32; dd += 8) {
      const int d = d0 + dd;
      if (d >= dend) continue;
#pragma unroll
      for (int cc = lane; cc < CB; cc += 32) {
        const int c = c0 + cc;
        if (c < ncols) {
          const float v = tile[cc][dd];
          dk[int64_t(d) * Kq + col_base + c] = v;
          if (pdk != nullptr) {
            __nv_bfloat16 hi, lo;
            split_bf16(v, hi, lo);
            const int64_t o = int64_t(d) * planes * Kq + col_base + c;
            pdk[o] = hi;
            if (planes == 2) pdk[o + Kq] = lo;
          }
        }
      }
    }
    __syncthreads();
  }
}

// Vectorised variant for even D and 8-byte aligned sources (the normal case): 64 columns x 64 dims per
// block, float2 loads, bf16x2 / float2 stores, so every global access moves 128-256 B per warp.
// Needs the norms pre-pass.  Falls back to scalar stores when the first destination column is odd.
__global__ void __launch_bounds__(256)
enqueue_vec_kernel(int nsamples, int D, EnqueueArgs a, const float* __restrict__ norms,
                   int64_t* __restrict__ queue_ptr, int ptr, int new_ptr, const int32_t* __restrict__ staged) {
  if (staged != nullptr && *staged == 0) return;
  constexpr int CB = 64, DB = 64;
  __shared__ float inv[CB];
  __shared__ int64_t soff[CB];
  __shared__ float tile[CB][DB + 1];     // [column][d]
  const int qi = blockIdx.z;
  const int mult = a.mult[qi];
  const int ncols = nsamples * mult;
  const int c0 = blockIdx.x * CB;
  if (ptr >= 0) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) queue_ptr[0] = new_ptr;
  } else {
    ptr = int(queue_ptr[0]);
  }
  if (c0 >= ncols) return;
  const int Kq = a.Kq[qi];
  const int col_base = ptr * mult;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src = a.src[qi] + enqueue_slot_offset(a);
  if (threadIdx.x < CB) {
    const int c = min(c0 + threadIdx.x, ncols - 1);
    const int smp = c / mult, f = c - smp * mult;
    soff[threadIdx.x] = int64_t(smp) * a.src_stride[qi] + int64_t(f) * D;
    inv[threadIdx.x] = (c0 + threadIdx.x < ncols) ? (a.prenormalised ? 1.0f : 1.0f / norms[a.norm_off[qi] + c0 + threadIdx.x]) : 0.f;
  }
  __syncthreads();
  const int planes = a.planes;
  float* dk = a.dk[qi];
  __nv_bfloat16* pkd = a.pack_kd[qi];
  __nv_bfloat16* pdk = a.pack_dk[qi];
  const int d0 = blockIdx.y * DB;
  const int d = d0 + 2 * lane;           // this lane's two dims in the read phase
  // read: warp w takes columns w, w+8, ...; x * (1/n) is within 1 ulp of F.normalize's x / n
#pragma unroll
  for (int i = 0; i < CB / 8; ++i) {
    const int cc = warp + 8 * i, c = c0 + cc;
    float2 v = make_float2(0.f, 0.f);
    if (c < ncols && d < D) {
      v = __ldg(reinterpret_cast<const float2*>(src + soff[cc] + d));
      const float s = inv[cc];
      v.x *= s; v.y *= s;
      if (pkd != nullptr) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v.x, h0, l0);
        split_bf16(v.y, h1, l1);
        const int64_t o = int64_t(col_base + c) * planes * D + d;
        *reinterpret_cast<__nv_bfloat162*>(pkd + o) = __nv_bfloat162(h0, h1);
        if (planes == 2) *reinterpret_cast<__nv_bfloat162*>(pkd + o + D) = __nv_bfloat162(l0, l1);
      }
    }
    tile[cc][2 * lane] = v.x;
    tile[cc][2 * lane + 1] = v.y;
  }
  __syncthreads();
  // write transposed: warp w takes dims w, w+8, ...; lane l the columns 2l, 2l+1
  const int cpair = c0 + 2 * lane;
  const bool even = ((col_base + c0) & 1) == 0;
#pragma unroll
  for (int i = 0; i < DB / 8; ++i) {
    const int dd = warp + 8 * i, dg = d0 + dd;
    if (dg >= D || cpair >= ncols) continue;
    const float v0 = tile[2 * lane][dd], v1 = tile[2 * lane + 1][dd];
    const bool two = cpair + 1 < ncols;
    float* o = dk + int64_t(dg) * Kq + col_base + cpair;
    if (even && two) *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
    else { o[0] = v0; if (two) o[1] = v1; }
    if (pdk != nullptr) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v0, h0, l0);
      split_bf16(v1, h1, l1);
      __nv_bfloat16* q = pdk + int64_t(dg) * planes * Kq + col_base + cpair;
      if (even && two) {
        *reinterpret_cast<__nv_bfloat162*>(q) = __nv_bfloat162(h0, h1);
        if (planes == 2) *reinterpret_cast<__nv_bfloat162*>(q + Kq) = __nv_bfloat162(l0, l1);
      } else {
        q[0] = h0; if (two) q[1] = h1;
        if (planes == 2) { q[Kq] = l0; if (two) q[Kq + 1] = l1; }
      }
    }
  }
}

// ------------------------------------------------------------------ pack / unpack rows
struct RowPackArgs {
  uint64_t ptrs[8];
  int widths[8];
  int offs[8];
  int n;
  int total;
};

// staged (PACK only): NULL or a device flag set to 1 once the rows are packed: the deferred key exchange marks
// its send buffer "keys staged", the enqueue consumes the mark (hmmc_enqueue_norm)
template <bool PACK>
__global__ void rowpack_kernel(RowPackArgs a, float* __restrict__ packed, int64_t rows, int32_t* staged) {
  const int64_t row = blockIdx.x;
  const int t = blockIdx.y;
  if (PACK && staged != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *staged = 1;
  float* x = reinterpret_cast<float*>(a.ptrs[t]) + row * a.widths[t];
  float* y = packed + row * a.total + a.offs[t];
  const int w = a.widths[t];
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && (w & 3) == 0) {
    float4* x4 = reinterpret_cast<float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (int i = threadIdx.x; i < w / 4; i += blockDim.x) {
      if (PACK) y4[i] = x4[i]; else x4[i] = y4[i];
    }
    return;
  }
  for (int i = threadIdx.x; i < w; i += blockDim.x) {
    if (PACK) y[i] = x[i]; else x[i] = y[i];
  }
}

// ---------------------------------------------------------------- key exchange over peer memory (NVLink)
// The ranks of one node map each other's receive buffers (symmetric memory).  Every rank PUSHES its block of rows
// into all receive buffers with plain 16-byte stores (posted writes: no round trip per access), then raises its
// flag in every peer; a rank's enqueue starts once all flags have reached the current exchange number.  Two slots
// alternate so that a fast rank's next push never lands in a buffer a slow rank still reads (a rank pushes
// exchange e+2 only after it has seen every peer's flag e+1, which that peer raised after its enqueue e).
// All counters live on the device: a captured step replays unchanged.
struct PeerPtrs { uint64_t p[HMMC_MAX_PEERS]; };

// Packing with the enqueue's normalisation folded in: every D-vector of every block is written as
// x * (1 / max(||x||, 1e-12)), the sum of squares accumulated exactly as key_norms_kernel does, so the values are
// bit-identical to what the enqueue would have computed from the raw keys.  Each rank then normalises only its own
// keys (not all W*b of them after the exchange) and the enqueue reads the received rows once.
__global__ void __launch_bounds__(256)
rowpack_norm_kernel(RowPackArgs a, float* __restrict__ packed, int D, int64_t rows, int32_t* staged) {
  // one warp per D-vector, vectors numbered row-major over the packed row (total / D per row)
  if (staged != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *staged = 1;
  const int lane = threadIdx.x & 31;
  const int per_row = a.total / D;
  const int64_t g = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= rows * per_row) return;
  const int64_t row = g / per_row;
  const int off = int(g - row * per_row) * D;          // element offset inside the packed row
  int t = 0;
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (i < a.n && off >= a.offs[i]) t = i;
  const float* xv = reinterpret_cast<const float*>(a.ptrs[t]) + row * a.widths[t] + (off - a.offs[t]);
  float* y = packed + row * a.total + off;
  // the row stays in registers between the two passes (D <= 1024: 32 values per lane); the sum of squares is
  // accumulated in the same order as key_norms_kernel (lane-strided, then the shuffle tree)
  float e[32];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int d = lane + 32 * k;
    e[k] = (d < D) ? xv[d] : 0.f;
    if (d < D) ss = fmaf(e[k], e[k], ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int d = lane + 32 * k;
    if (d < D) y[d] = e[k] * inv;
  }
}


__global__ void __launch_bounds__(256)
peer_push_kernel(const float4* __restrict__ send, int64_t n4, PeerPtrs bufs, PeerPtrs flags, int W, int rank,
                 int64_t slot_stride, const int32_t* __restrict__ epoch, unsigned* done) {
  const int e = *epoch;                                         // exchanges completed so far
  const int peer = (rank + int(blockIdx.y)) % W;                // staggered: no two ranks start on the same peer
  float4* dst = reinterpret_cast<float4*>(bufs.p[peer]) + (int64_t(e & 1) * slot_stride + int64_t(rank) * n4 * 4) / 4;
  // four independent 16-byte loads in flight per thread before the posted stores
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 v0 = send[i], v1 = send[i + stride], v2 = send[i + 2 * stride], v3 = send[i + 3 * stride];
    dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
  }
  for (; i < n4; i += stride) dst[i] = send[i];
  // last block: every block's stores are ordered before its arrival (system-scope fence), the flags after all arrivals
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x < W) {
    volatile int32_t* f = reinterpret_cast<volatile int32_t*>(flags.p[threadIdx.x]) + rank;
    *f = e + 1;
  }
  if (threadIdx.x == 0) *done = 0u;
}

// Waits until every rank's rows of the current exchange have landed here, then counts the exchange.
__global__ void peer_wait_kernel(const volatile int32_t* my_flags, int W, int32_t* epoch) {
  const int e = *epoch + 1;
  if (threadIdx.x < W) {
    const long long t0 = clock64();
    while (my_flags[threadIdx.x] < e)
      if (clock64() - t0 > (1ll << 39)) __trap();             // minutes (a peer may be busy saving a checkpoint): a rank
                                                              // that never pushes is an error in the end, not a hang
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *epoch = e;
}

struct ScaleArgs {
  float* ptrs[8];
  int64_t numels[8];
  int n;
};
// x_t *= scale[0] for up to 8 tensors (backward of the fused heads: the gradients were produced
// with the loss, the upstream gradient arrives later)
__global__ void scale_tensors_kernel(const ScaleArgs a, const float* __restrict__ scale) {
  ptx::grid_dependency_wait();
  const float s = scale[0];
  if (s == 1.0f) return;               // loss.backward() with no upstream scaling: nothing to do
  float* x = a.ptrs[blockIdx.y];
  const int64_t n = a.numels[blockIdx.y];
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 v = x4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    x4[i] = v;
  }
  for (int64_t i = n4 * 4 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    x[i] *= s;
}

static int build_rowpack(RowPackArgs& a, const uint64_t* ptrs, const int32_t* widths, int n) {
  HMMC_REQUIRE(n >= 1 && n <= 8, "pack_rows: between 1 and 8 blocks, got %d", n);
  a.n = n;
  int off = 0;
  for (int i = 0; i < n; ++i) {
    a.ptrs[i] = ptrs[i];
    a.widths[i] = widths[i];
    a.offs[i] = off;
    off += widths[i];
  }
  a.total = off;
  return HMMC_OK;
}

}  // namespace hmmc

using namespace hmmc;

extern "C" {

int hmmc_queue_pack(const hmmc_queue* q, void* stream) {
  HMMC_REQUIRE(q != nullptr && q->dk != nullptr && q->D > 0 && q->Kq > 0, "queue_pack: bad queue");
  HMMC_REQUIRE(q->planes == 1 || q->planes == 2, "queue_pack: planes must be 1 or 2");
  if (q->pack_kd == nullptr && q->pack_dk == nullptr) return HMMC_OK;
  dim3 grid((q->Kq + 31) / 32, (q->D + 31) / 32);
  queue_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q->dk, static_cast<__nv_bfloat16*>(q->pack_kd), static_cast<__nv_bfloat16*>(q->pack_dk), q->D, q->Kq, q->planes);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

}  // extern "C"

namespace hmmc {
// ---------------------------------------------------------------------------------------
// Generic fused InfoNCE driver: several query tensors ("groups"), each contracted against one
// or two queues ("blocks" = GEMM problems), in four launches:
//   prep_rows (normalise + pack) -> S-GEMM with the exp / row-sum / E epilogue (grouped)
//   -> U-GEMM (grouped, split-K, units placed longest-first) -> finish (positives, loss, gradients, loss sums)
struct BlockDesc {
  int group;                 // which query tensor
  const float* keys;
  int pos_mode, Fk;
  const hmmc_queue* queue;
  float coef;                // weight / b
  int kind;                  // loss slot
};
struct GroupDesc {
  const float* q;
  float* dq;
  int rows, Fq;
};

struct InfoNCELayout {       // workspace carving shared by the size query and the run
  LossAcc* acc;
  float* xhat[MAX_GROUPS];
  __nv_bfloat16* packed[MAX_GROUPS];
  float* rowsum_part[MAX_BLOCKS];
  void* E[MAX_BLOCKS];
  float* U_part[MAX_BLOCKS];
  int nparts[MAX_BLOCKS], splits[MAX_BLOCKS], nsplits_eff[MAX_BLOCKS];
  int bn1, bn2;
};

// Split-K of the U-GEMMs (U = E.Q^T: a small output, a long contraction).  The launch is a single wave of
// units (output tile x K slice) placed longest-first on the CTA pairs.  Every tile is cut into slices of `unit`
// k-block steps plus a shorter remainder: the mix of long and short units packs the pairs far more evenly than
// equal slices do (b = 256: makespan 85 steps against a mean of 82, where three equal slices give 128 and two
// give 102), with a charge for every extra partial tile the finish kernel has to read back.
// Costs are in k-block steps of one CTA pair (64 contraction elements of a 256 x 256 tile).
struct SplitChoice { int unit[MAX_BLOCKS]; };     // k-block steps per slice, per problem

static SplitChoice choose_u_splits(const int* rows, const int* Kq, int nb, int D, int planes, int bn2) {
  typedef std::tuple<int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int> Key;
  static std::mutex mu;
  static std::map<Key, SplitChoice> cache;
  int kr[MAX_BLOCKS] = {0}, kk[MAX_BLOCKS] = {0};
  for (int k = 0; k < nb; ++k) { kr[k] = rows[k]; kk[k] = Kq[k]; }
  const int workers = (bn2 == 256) ? sm_count() / 2 : sm_count();
  const Key key(nb, D, planes, bn2, kr[0], kr[1], kr[2], kr[3], kr[4], kr[5], kk[0], kk[1], kk[2], kk[3], kk[4], kk[5]);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  const int nseg = (planes == 2) ? 3 : 1;
  const int m_tile = (bn2 == 256) ? 2 * UMMA_BM : UMMA_BM;
  int tiles[MAX_BLOCKS], total_kb[MAX_BLOCKS], max_kb = 1;
  for (int k = 0; k < nb; ++k) {
    tiles[k] = ((rows[k] + m_tile - 1) / m_tile) * ((D + bn2 - 1) / bn2);
    total_kb[k] = nseg * (Kq[k] / UMMA_BK);
    max_kb = std::max(max_kb, total_kb[k]);
  }
  SplitChoice best;
  double best_cost = 1e30;
  for (int unit = 8; unit <= max_kb; ++unit) {
    std::vector<int> cost;
    int units = 0;
    bool ok = true;
    for (int k = 0; k < nb && ok; ++k) {
      const int u = std::min(unit, total_kb[k]);
      const int n_slices = (total_kb[k] + u - 1) / u;
      ok = n_slices <= 32;
      for (int s = 0; s < n_slices; ++s) {
        const int len = std::min(u, total_kb[k] - s * u);
        for (int t = 0; t < tiles[k]; ++t) cost.push_back(len + UMMA_UNIT_FIXED_COST);
      }
      units += n_slices * tiles[k];
    }
    if (!ok || units > UMMA_MAX_UNITS) continue;
    // every partial tile is written once and read once more by the finish kernel: ~0.25 steps of chip time each
    const double total = double(lpt_makespan(cost, workers)) + 0.25 * units;
    if (total < best_cost) {
      best_cost = total;
      for (int k = 0; k < MAX_BLOCKS; ++k) best.unit[k] = k < nb ? std::min(unit, total_kb[k]) : 1;
    }
  }
  if (best_cost > 1e29)
    for (int k = 0; k < MAX_BLOCKS; ++k) best.unit[k] = k < nb ? total_kb[k] : 1;
  std::lock_guard<std::mutex> lk(mu);
  cache[key] = best;
  return best;
}

static void infonce_layout(Workspace& ws, InfoNCELayout& L, const GroupDesc* groups, int ng, const int* blk_group,
                           const int* blk_Kq, int nb, int D, int prec, bool need_grad) {
  const int planes = planes_of(prec);
  int total_rows = 0;
  for (int i = 0; i < ng; ++i) total_rows += groups[i].rows;
  L.acc = ws.take<LossAcc>(1);
  for (int i = 0; i < ng; ++i) {
    L.xhat[i] = (prec == HMMC_PREC_FP32) ? ws.take<float>(size_t(groups[i].rows) * D) : nullptr;
    L.packed[i] = (prec != HMMC_PREC_FP32) ? ws.take<__nv_bfloat16>(size_t(groups[i].rows) * planes * D) : nullptr;
  }
  L.bn1 = 256;
  for (int k = 0; k < nb; ++k)
    if (blk_Kq[k] % 256 != 0) L.bn1 = 128;
  L.bn2 = (D % 256 == 0) ? 256 : 128;
  const int nseg_layout = (planes == 2) ? 3 : 1;
  SplitChoice sc;
  for (int k = 0; k < MAX_BLOCKS; ++k) sc.unit[k] = k < nb ? nseg_layout * (blk_Kq[k] / UMMA_BK) : 1;
  if (prec != HMMC_PREC_FP32 && need_grad) {
    int rows[MAX_BLOCKS];
    for (int k = 0; k < nb; ++k) rows[k] = groups[blk_group[k]].rows;
    sc = choose_u_splits(rows, blk_Kq, nb, D, planes, L.bn2);
  }
  for (int k = 0; k < nb; ++k) {
    const int R = groups[blk_group[k]].rows;
    const int Kq = blk_Kq[k];
    if (prec == HMMC_PREC_FP32) {
      L.nparts[k] = 1;
      L.splits[k] = L.nsplits_eff[k] = 1;
      L.rowsum_part[k] = ws.take<float>(size_t(R));
      L.E[k] = ws.take<float>(size_t(R) * Kq);
    } else {
      // one partial per epilogue warp column range: the CTA-pair kernel runs four warps per TMEM lane quadrant
      // for the single-plane epilogues, two otherwise (EpiInfoNCE::PARTS_PER_TILE_PAIR)
      const int parts_per_tile = (L.bn1 == 256) ? ((need_grad && planes == 2) ? EpiInfoNCE<2>::PARTS_PER_TILE_PAIR
                                                                              : EpiInfoNCE<1>::PARTS_PER_TILE_PAIR)
                                                : 2;
      L.nparts[k] = parts_per_tile * ((Kq + L.bn1 - 1) / L.bn1);
      const int total_kb = nseg_layout * (Kq / UMMA_BK);
      L.splits[k] = sc.unit[k] > 0 ? sc.unit[k] : total_kb;           // k-block steps per split-K slice
      L.nsplits_eff[k] = (total_kb + L.splits[k] - 1) / L.splits[k];
      L.rowsum_part[k] = ws.take<float>(size_t(L.nparts[k]) * R);
      L.E[k] = need_grad ? ws.take<__nv_bfloat16>(size_t(R) * planes * Kq) : nullptr;
    }
    L.U_part[k] = need_grad ? ws.take<float>(size_t(L.nsplits_eff[k]) * R * D) : nullptr;
  }
}

// sched (may be null = everything, no event): which half of the work to issue, the SMs the GEMM grids leave free
// and the event to record once the queues are no longer read (hmmc_head_schedule in include/hmmc_head.h)
static int run_infonce(const GroupDesc* groups, int ng, const BlockDesc* blocks, int nb, int b, int D, float temperature,
                       int prec, const LossFinal& fin, const hmmc_head_schedule* sched, void* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
  const int phase = sched ? sched->phase : 0;
  const int reserved_sms = sched ? sched->reserved_sms : 0;
  cudaEvent_t release = sched ? static_cast<cudaEvent_t>(sched->queues_released) : nullptr;
  HMMC_REQUIRE(phase >= 0 && phase <= 2, "infonce: schedule phase must be 0, 1 or 2 (got %d)", phase);
  HMMC_REQUIRE(reserved_sms >= 0, "infonce: reserved_sms must be >= 0 (got %d)", reserved_sms);
  HMMC_REQUIRE(ng >= 1 && ng <= MAX_GROUPS && nb >= 1 && nb <= MAX_BLOCKS, "infonce: too many groups/blocks");
  HMMC_REQUIRE(D > 0 && D <= FIN_MAXD, "infonce: D=%d exceeds the supported %d", D, FIN_MAXD);
  HMMC_REQUIRE(prec >= 0 && prec <= 2, "infonce: unknown precision %d", prec);
  // constant-max log-sum-exp: all logits lie in [-1/T, 1/T]; exp(-2/T) must stay a normal fp32
  HMMC_REQUIRE(temperature >= 0.025f, "infonce: temperature %g < 0.025 is not supported by the constant-max LSE",
               temperature);
  const int planes = planes_of(prec);
  bool need_grad = false;
  for (int i = 0; i < ng; ++i) need_grad = need_grad || groups[i].dq != nullptr;
  int blk_group[MAX_BLOCKS], blk_Kq[MAX_BLOCKS];
  for (int k = 0; k < nb; ++k) {
    const hmmc_queue* q = blocks[k].queue;
    HMMC_REQUIRE(q != nullptr && q->dk != nullptr && q->D == D, "infonce: queue %d has D=%d, embeddings D=%d", k, q ? q->D : -1, D);
    blk_group[k] = blocks[k].group;
    blk_Kq[k] = q->Kq;
    if (prec != HMMC_PREC_FP32) {
      HMMC_REQUIRE(q->pack_kd && q->pack_dk, "infonce: queue has no packed operands (call hmmc_queue_pack)");
      HMMC_REQUIRE(q->planes == planes, "infonce: queue packed with %d planes, precision needs %d", q->planes, planes);
      HMMC_REQUIRE(D % UMMA_BK == 0 && q->Kq % 128 == 0,
                   "infonce: tensor-core path needs D %% 64 == 0 and Kq %% 128 == 0 (D=%d Kq=%d)", D, q->Kq);
    }
  }
  Workspace ws(workspace, workspace_bytes);
  InfoNCELayout L;
  infonce_layout(ws, L, groups, ng, blk_group, blk_Kq, nb, D, prec, need_grad);
  if (!ws.ok()) {
    set_error("infonce: workspace too small: need %zu bytes, got %zu", ws.used, workspace_bytes);
    return HMMC_ERR_WORKSPACE;
  }
  const float LOG2E = 1.4426950408889634f;
  const float invT = 1.0f / temperature, cmax = invT;
  // tensor-core path: the packed queries carry log2(e)/T and the epilogue stores 2^acc; the constant max comes
  // back as one factor in the finish kernel
  const float pack_scale = (prec == HMMC_PREC_FP32) ? 1.0f : invT * LOG2E;
  const float kexp = (prec == HMMC_PREC_FP32) ? 1.0f : expf(-cmax);
  // 1. normalise (+ pack) every query tensor
  PrepArgs pa;
  pa.n = ng;
  pa.row_begin[0] = 0;
  for (int i = 0; i < MAX_GROUPS; ++i) {
    const int k = i < ng ? i : 0;
    pa.x[i] = groups[k].q;
    pa.xhat[i] = L.xhat[k];
    pa.packed[i] = L.packed[k];
    pa.row_begin[i + 1] = pa.row_begin[i] + (i < ng ? groups[i].rows : 0);
  }
  const int total_rows = pa.row_begin[ng];
  bool vec_q = (D % 128 == 0) && D <= 1024;
  for (int i = 0; i < ng && vec_q; ++i)
    vec_q = (reinterpret_cast<uintptr_t>(groups[i].q) % 16 == 0) &&
            (groups[i].dq == nullptr || reinterpret_cast<uintptr_t>(groups[i].dq) % 16 == 0);
  int rc;
  if (phase != 2) {
    const dim3 pgrid((total_rows + 7) / 8), pblock(256);
    cudaError_t e;
    if (vec_q && D == 512) e = launch_pdl(prep_rows_kernel<4>, pgrid, pblock, 0, st, pa, D, planes, pack_scale, L.acc);
    else if (vec_q && D == 256) e = launch_pdl(prep_rows_kernel<2>, pgrid, pblock, 0, st, pa, D, planes, pack_scale, L.acc);
    else if (vec_q && D == 128) e = launch_pdl(prep_rows_kernel<1>, pgrid, pblock, 0, st, pa, D, planes, pack_scale, L.acc);
    else if (vec_q && D == 1024) e = launch_pdl(prep_rows_kernel<8>, pgrid, pblock, 0, st, pa, D, planes, pack_scale, L.acc);
    else e = launch_pdl(prep_rows_kernel<0>, pgrid, pblock, 0, st, pa, D, planes, pack_scale, L.acc);
    count_launch();
    HMMC_CHECK_CUDA(e);
    if (prec == HMMC_PREC_FP32) {
      for (int k = 0; k < nb; ++k) {
        const GroupDesc& G = groups[blocks[k].group];
        const int Kq = blk_Kq[k];
        float* S = static_cast<float*>(L.E[k]);
        if ((rc = gemm_f32(L.xhat[blocks[k].group], D, 1, blocks[k].queue->dk, 1, Kq, S, Kq, G.rows, Kq, D, 1.0f, st))) return rc;
        exp_rowsum_kernel<<<G.rows, 256, 0, st>>>(S, Kq, Kq, invT, cmax, L.rowsum_part[k]);
        HMMC_CHECK_LAUNCH();
        if (need_grad && (rc = gemm_f32(S, Kq, 1, blocks[k].queue->dk, Kq, 1, L.U_part[k], D, G.rows, D, Kq, 1.0f, st))) return rc;
      }
    } else {
      GemmProblem<EpiStoreF32> p2[MAX_BLOCKS];
      for (int k = 0; k < nb; ++k) {
        const GroupDesc& G = groups[blocks[k].group];
        const int Kq = blk_Kq[k];
        EpiStoreF32::Params e2{L.U_part[k], int64_t(D), int64_t(G.rows) * D, 1.0f};
        p2[k] = GemmProblem<EpiStoreF32>{L.E[k], int64_t(planes) * Kq, blocks[k].queue->pack_dk, int64_t(planes) * Kq,
                                         G.rows, D, Kq, planes, 1, e2};
        p2[k].kb_per_split = L.splits[k];
      }
      // S-GEMM, epilogue specialised on the number of E planes it writes (0 = forward only).
      // CTA-pair kernels (256 x 256 tiles) whenever the tile width divides the problem.
      auto launch_s = [&](auto epi_tag, bool pair) -> int {
        using Epi = decltype(epi_tag);
        GemmProblem<Epi> p1[MAX_BLOCKS];
        for (int k = 0; k < nb; ++k) {
          const GroupDesc& G = groups[blocks[k].group];
          const int Kq = blk_Kq[k];
          typename Epi::Params e1;
          e1.rowsum_part = L.rowsum_part[k];
          e1.lo_col0 = Kq;
          p1[k] = GemmProblem<Epi>{L.packed[blocks[k].group], int64_t(planes) * D, blocks[k].queue->pack_kd,
                                   int64_t(planes) * D, G.rows, Kq, D, planes, 1, e1,
                                   L.E[k], int64_t(planes) * Kq, int64_t(planes) * Kq};
        }
        if constexpr (Epi::PAIR_WARPS == 8 && Epi::STORE_COLS == 32) {
          if (!pair) return launch_umma_grouped<128, Epi>(p1, nb, st, reserved_sms);
        }
        return launch_umma_grouped_pair<Epi>(p1, nb, st, reserved_sms);
      };
      const int ep_planes = need_grad ? planes : 0;
      if (L.bn1 == 256) {
        if (ep_planes == 0) rc = launch_s(EpiInfoNCE<0>(), true);
        else if (ep_planes == 2) rc = launch_s(EpiInfoNCE<2>(), true);
        else rc = launch_s(EpiInfoNCE<1>(), true);
      } else {
        // 128-wide single-CTA tiles (Kq not a multiple of 256): eight epilogue warps, one chunk per store
        if (ep_planes == 0) rc = launch_s(EpiInfoNCE<0, 8, 6, 1, 2>(), false);
        else if (ep_planes == 2) rc = launch_s(EpiInfoNCE<2, 8, 6, 1, 2>(), false);
        else rc = launch_s(EpiInfoNCE<1, 8, 6, 1, 2>(), false);
      }
      if (rc) return rc;
      if (need_grad) {
        if (L.bn2 == 256) rc = launch_umma_grouped_pair<EpiStoreF32>(p2, nb, st, reserved_sms);
        else rc = launch_umma_grouped<128, EpiStoreF32>(p2, nb, st, reserved_sms);
        if (rc) return rc;
      }
    }
    // every kernel that reads the queues has been issued: let the enqueue start on another stream
  }
  if (release != nullptr) HMMC_CHECK_CUDA(cudaEventRecord(release, st));
  if (phase == 1) return HMMC_OK;
  // 4. positives, loss, gradient, loss sums
  // Row order of the finish kernel: one warp per row, rows of a query tensor that owns a ONE_TO_FRAMES block
  // first.  Those rows walk Fk key rows each (12 dependent rounds of loads) and set the kernel's duration unless
  // they start with the first wave of blocks.
  int slot_of[MAX_GROUPS], group_at[MAX_GROUPS];
  {
    int n = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < ng; ++i) {
        bool heavy = false;
        for (int k = 0; k < nb; ++k) heavy = heavy || (blocks[k].group == i && blocks[k].pos_mode == HMMC_POS_ONE_TO_FRAMES);
        if (heavy == (pass == 0)) { slot_of[i] = n; group_at[n] = i; ++n; }
      }
  }
  FinishArgs fa;
  fa.n = ng;
  fa.row_begin[0] = 0;
  for (int i = 0; i < MAX_GROUPS; ++i) {
    const int gi = i < ng ? group_at[i] : group_at[0];
    RowGroup& R = fa.g[i];
    R.q = groups[gi].q;
    R.dq = groups[gi].dq;
    R.rows = groups[gi].rows;
    R.Fq = groups[gi].Fq;
    R.ncontrib = 0;
    fa.row_begin[i + 1] = fa.row_begin[i] + (i < ng ? groups[gi].rows : 0);
  }
  for (int k = 0; k < nb; ++k) {
    RowGroup& R = fa.g[slot_of[blocks[k].group]];
    HMMC_REQUIRE(R.ncontrib < 2, "infonce: a query tensor may feed at most two queues");
    Contribution& C = R.c[R.ncontrib++];
    C.keys = blocks[k].keys;
    C.rowsum_part = L.rowsum_part[k];
    C.U_part = L.U_part[k];
    C.split_stride = int64_t(R.rows) * D;
    C.pos_mode = blocks[k].pos_mode;
    C.Fk = blocks[k].Fk;
    C.n_parts = L.nparts[k];
    C.n_splits = L.nsplits_eff[k];
    C.coef = blocks[k].coef;
    C.kind = blocks[k].kind;
    vec_q = vec_q && (reinterpret_cast<uintptr_t>(C.keys) % 16 == 0);
  }
  for (int i = 0; i < MAX_GROUPS; ++i)
    if (fa.g[i].ncontrib < 2) fa.g[i].c[1] = fa.g[i].c[0];
  // vector kernel: the rows of the leading (ONE_TO_FRAMES) tensors get a whole block each (finish_row_block)
  bool vec_kernel = vec_q && (D == 128 || D == 256 || D == 512 || D == 1024);
  for (int i = 0; i < ng; ++i) vec_kernel = vec_kernel && (groups[i].rows % FIN_WARPS == 0);   // a block = one tensor
  // Worth it while the grid stays within two waves: then the kernel lasts as long as its slowest row.  Beyond
  // that it is bound by block slots and the extra blocks cost more than they save (b = 128: -2.9 us, b = 256:
  // +4.4 us, profiles/r2_finish_kernel.md).
  fa.heavy_rows = 0;
  if (vec_kernel && HMMC_FIN_HEAVY) {
    int heavy_rows = 0;
    for (int i = 0; i < ng; ++i) {
      bool heavy = false;
      for (int k = 0; k < nb; ++k)
        heavy = heavy || (blocks[k].group == group_at[i] && blocks[k].pos_mode == HMMC_POS_ONE_TO_FRAMES);
      if (!heavy) break;
      heavy_rows = fa.row_begin[i + 1];
    }
    const int grid_blocks = heavy_rows + (total_rows - heavy_rows + FIN_WARPS - 1) / FIN_WARPS;
    const int slots = sm_count() * HMMC_FIN_OCC * 8 / FIN_WARPS;
    if (HMMC_FIN_HEAVY == 2 || grid_blocks <= 2 * slots) fa.heavy_rows = heavy_rows;
  }
  const int fin_warps = vec_kernel ? FIN_WARPS : 8;
  const int fin_blocks = fa.heavy_rows + (total_rows - fa.heavy_rows + fin_warps - 1) / fin_warps;
  const dim3 fgrid(fin_blocks), fblock(32 * fin_warps);
  const int ne = (D + 31) / 32;
  cudaError_t e;
  if (vec_kernel && D == 512) e = launch_pdl(infonce_finish_vec_kernel<4>, fgrid, fblock, 0, st, fa, invT, cmax, kexp, fin, L.acc);
  else if (vec_kernel && D == 256) e = launch_pdl(infonce_finish_vec_kernel<2>, fgrid, fblock, 0, st, fa, invT, cmax, kexp, fin, L.acc);
  else if (vec_kernel && D == 128) e = launch_pdl(infonce_finish_vec_kernel<1>, fgrid, fblock, 0, st, fa, invT, cmax, kexp, fin, L.acc);
  else if (vec_kernel && D == 1024) e = launch_pdl(infonce_finish_vec_kernel<8>, fgrid, fblock, 0, st, fa, invT, cmax, kexp, fin, L.acc);
  else if (ne <= 4) e = launch_pdl(infonce_finish_kernel<4>, fgrid, fblock, 0, st, fa, D, invT, cmax, kexp, fin, L.acc);
  else if (ne <= 16) e = launch_pdl(infonce_finish_kernel<16>, fgrid, fblock, 0, st, fa, D, invT, cmax, kexp, fin, L.acc);
  else if (ne <= 32) e = launch_pdl(infonce_finish_kernel<32>, fgrid, fblock, 0, st, fa, D, invT, cmax, kexp, fin, L.acc);
  else e = launch_pdl(infonce_finish_kernel<FIN_MAXE>, fgrid, fblock, 0, st, fa, D, invT, cmax, kexp, fin, L.acc);
  count_launch();
  HMMC_CHECK_CUDA(e);
  return HMMC_OK;
}

static int check_pos_mode(int pos_mode, int Fq, int Fk) {
  HMMC_REQUIRE(pos_mode >= 0 && pos_mode <= 3, "infonce: unknown pos_mode %d", pos_mode);
  if (pos_mode == HMMC_POS_PAIR) HMMC_REQUIRE(Fq == Fk, "infonce: PAIR needs Fq == Fk");
  if (pos_mode == HMMC_POS_FRAME_NEIGHBOUR) HMMC_REQUIRE(Fq == Fk && Fq >= 2, "infonce: FRAME_NEIGHBOUR needs Fq == Fk >= 2");
  if (pos_mode == HMMC_POS_ONE_TO_FRAMES) HMMC_REQUIRE(Fq == 1, "infonce: ONE_TO_FRAMES needs Fq == 1");
  if (pos_mode == HMMC_POS_FRAMES_TO_ONE) HMMC_REQUIRE(Fk == 1, "infonce: FRAMES_TO_ONE needs Fk == 1");
  return HMMC_OK;
}

}  // namespace hmmc

extern "C" {

size_t hmmc_infonce_workspace_bytes(int64_t R, int D, int Kq, int prec) {
  Workspace ws(nullptr, 0);
  InfoNCELayout L;
  GroupDesc g{nullptr, nullptr, int(R), 1};
  int bg = 0, bk = Kq;
  infonce_layout(ws, L, &g, 1, &bg, &bk, 1, D, prec, true);
  return ws.used + 1024;
}

int hmmc_infonce_queue_fwd_bwd(const float* q, const float* keys, int pos_mode, int b, int Fq, int Fk, int D,
                               const hmmc_queue* queue, float temperature, float weight, int prec, float* loss_out,
                               float* dq, void* workspace, size_t workspace_bytes, void* stream) {
  HMMC_REQUIRE(q && keys && queue && loss_out, "infonce: null argument");
  HMMC_REQUIRE(b > 0 && Fq > 0 && Fk > 0 && D > 0, "infonce: bad sizes b=%d Fq=%d Fk=%d D=%d", b, Fq, Fk, D);
  int rc = check_pos_mode(pos_mode, Fq, Fk);
  if (rc) return rc;
  HMMC_REQUIRE(workspace != nullptr, "infonce: null workspace");
  GroupDesc g{q, dq, b * Fq, Fq};
  BlockDesc blk{0, keys, pos_mode, Fk, queue, weight / float(b), 0};
  LossFinal fin{loss_out, 0, 0.f, 0.f, 0.f, 1};
  return run_infonce(&g, 1, &blk, 1, b, D, temperature, prec, fin, nullptr, workspace, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
}

size_t hmmc_pretrain_head_workspace_bytes(int b, int F, int D, int K, int prec) {
  Workspace ws(nullptr, 0);
  InfoNCELayout L;
  GroupDesc g[4] = {{nullptr, nullptr, b * F, F}, {nullptr, nullptr, b, 1}, {nullptr, nullptr, b, 1}, {nullptr, nullptr, b * F, F}};
  const int bg[5] = {0, 1, 2, 2, 3};
  const int bk[5] = {K * F, K, K, K * F, K};
  infonce_layout(ws, L, g, 4, bg, bk, 5, D, prec, true);
  return ws.used + 1024;
}

int hmmc_pretrain_head_fwd_bwd(const hmmc_pretrain_io* io, int b, int F, int D, const hmmc_queue* q_v,
                               const hmmc_queue* q_title, const hmmc_queue* q_frame_proj,
                               const hmmc_queue* q_frame_cross, float temperature, float w_fam, float w_vtm,
                               float w_ftm, int use_frame_fea, int prec, float* losses_out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  return hmmc_pretrain_head_fwd_bwd_sched(io, b, F, D, q_v, q_title, q_frame_proj, q_frame_cross, temperature, w_fam,
                                          w_vtm, w_ftm, use_frame_fea, prec, losses_out, nullptr, workspace,
                                          workspace_bytes, stream);
}

int hmmc_pretrain_head_fwd_bwd_sched(const hmmc_pretrain_io* io, int b, int F, int D, const hmmc_queue* q_v,
                                     const hmmc_queue* q_title, const hmmc_queue* q_frame_proj,
                                     const hmmc_queue* q_frame_cross, float temperature, float w_fam, float w_vtm,
                                     float w_ftm, int use_frame_fea, int prec, float* losses_out,
                                     const hmmc_head_schedule* sched, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  HMMC_REQUIRE(io && q_v && q_title && q_frame_proj && q_frame_cross && losses_out, "pretrain_head: null argument");
  const int phase = sched ? sched->phase : 0;
  HMMC_REQUIRE(io->v_fea && io->title_fea && io->frame_fea && io->frame_pred, "pretrain_head: null embedding pointer");
  HMMC_REQUIRE(phase == 1 || (io->v_fea_k && io->title_fea_k && io->frame_fea_k && io->frame_proj_k),
               "pretrain_head: null key pointer");
  HMMC_REQUIRE(b > 0 && F >= 2 && D > 0, "pretrain_head: bad sizes b=%d F=%d D=%d", b, F, D);
  HMMC_REQUIRE(workspace != nullptr, "pretrain_head: null workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // query tensors: 0 frame_pred, 1 v_fea, 2 title_fea, 3 frame_fea
  GroupDesc groups[4] = {{io->frame_pred, io->d_frame_pred, b * F, F},
                         {io->v_fea, io->d_v_fea, b, 1},
                         {io->title_fea, io->d_title_fea, b, 1},
                         {io->frame_fea, io->d_frame_fea, b * F, F}};
  const float fb = float(b);
  BlockDesc blocks[5] = {
      // FAM, frame_self_loss(frame_pred, frame_proj_k, queue_frame_proj_ng)        modules/modeling.py:385
      {0, io->frame_proj_k, HMMC_POS_FRAME_NEIGHBOUR, F, q_frame_proj, 1.0f / float(F - 1) / fb, 0},
      // VTM, contrastive_loss(v_fea, title_fea_k, queue_title) + (title_fea, v_fea_k, queue_v)   :387-388
      {1, io->title_fea_k, HMMC_POS_PAIR, 1, q_title, 1.0f / fb, 1},
      {2, io->v_fea_k, HMMC_POS_PAIR, 1, q_v, 1.0f / fb, 1},
      // FTM, frame_cross_loss(frame_fea, frame_fea_k, queue_frame_cross, title_fea, title_fea_k, queue_title)  :398
      {2, io->frame_fea_k, HMMC_POS_ONE_TO_FRAMES, F, q_frame_cross, 1.0f / float(F) / fb, 2},
      {3, io->title_fea_k, HMMC_POS_FRAMES_TO_ONE, 1, q_title, 1.0f / float(F) / fb, 2}};
  // gradients carry the loss weights: scale each block's coefficient (the loss slots stay unweighted
  // only when all three weights are applied afterwards, so slots are reported weighted = w * loss)
  const float wk[3] = {w_fam, w_vtm, w_ftm};
  int nb = use_frame_fea ? 5 : 3;
  int ng = use_frame_fea ? 4 : 3;
  for (int k = 0; k < nb; ++k) blocks[k].coef *= wk[blocks[k].kind];
  LossFinal fin{losses_out, 1, w_fam, w_vtm, w_ftm, use_frame_fea};
  int rc = run_infonce(groups, ng, blocks, nb, b, D, temperature, prec, fin, sched, workspace, workspace_bytes, st);
  if (rc) return rc;
  if (phase != 1 && !use_frame_fea && io->d_frame_fea != nullptr)
    HMMC_CHECK_CUDA(cudaMemsetAsync(io->d_frame_fea, 0, sizeof(float) * size_t(b) * F * D, st));
  return HMMC_OK;
}

int hmmc_ema_block_elems(void) { return EMA_BLOCK_ELEMS; }

int hmmc_ema_multi(const uint64_t* pk_ptrs, const uint64_t* p_ptrs, const int64_t* numels, const int32_t* dtypes,
                   const int64_t* block_offsets, int n, int64_t total_blocks, float m, float one_minus_m, void* stream) {
  HMMC_REQUIRE(pk_ptrs && p_ptrs && numels && dtypes && block_offsets, "ema_multi: null table");
  if (n <= 0 || total_blocks <= 0) return HMMC_OK;
  HMMC_REQUIRE(total_blocks < (int64_t(1) << 31), "ema_multi: too many blocks");
  ema_multi_kernel<<<unsigned(total_blocks), EMA_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      pk_ptrs, p_ptrs, numels, dtypes, block_offsets, n, m, one_minus_m);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

// after the scatter: advance the device-side pointer (advance != 0) and consume the "keys staged" mark
__global__ void advance_ptr_kernel(int64_t* queue_ptr, int B, int K, int advance, int32_t* staged) {
  if (staged != nullptr) {
    if (*staged == 0) return;
    *staged = 0;
  }
  if (advance) queue_ptr[0] = (queue_ptr[0] + B) % K;
}

static int enqueue_common(const float* const* src5, const int64_t* stride5, int B, int F, int D,
                          const hmmc_queue* queues5, int64_t* queue_ptr, int64_t ptr_host, int K, float* scratch,
                          int32_t* staged, const int32_t* slot_epoch, int64_t slot_stride, int prenormalised,
                          cudaStream_t st) {
  HMMC_REQUIRE(queues5 && queue_ptr, "enqueue: null argument");
  // the reference's slice assignment raises when the batch does not fit (modules/modeling.py:273-280)
  const bool device_ptr = ptr_host < 0;     // pointer lives on the device only (graph replay)
  if (device_ptr)
    HMMC_REQUIRE(B <= K && K % B == 0, "enqueue (device pointer): queue size %d must be a multiple of the batch %d", K, B);
  else
    HMMC_REQUIRE(ptr_host + B <= K, "enqueue: ptr %lld + batch %d exceeds queue size %d", (long long)ptr_host, B, K);
  EnqueueArgs a;
  const int mult[5] = {1, 1, 1, F, F};
  a.planes = queues5[0].planes;
  a.slot_epoch = slot_epoch;
  a.slot_stride = slot_stride;
  a.prenormalised = prenormalised;
  for (int i = 0; i < 5; ++i) {
    const hmmc_queue& q = queues5[i];
    HMMC_REQUIRE(q.dk != nullptr && q.D == D && q.Kq == K * mult[i], "enqueue: queue %d has shape [%d,%d], expected [%d,%d]",
                 i, q.D, q.Kq, D, K * mult[i]);
    HMMC_REQUIRE(q.planes == a.planes, "enqueue: queues disagree on planes");
    HMMC_REQUIRE(src5[i] != nullptr, "enqueue: null key tensor %d", i);
    a.dk[i] = q.dk;
    a.pack_kd[i] = static_cast<__nv_bfloat16*>(q.pack_kd);
    a.pack_dk[i] = static_cast<__nv_bfloat16*>(q.pack_dk);
    a.Kq[i] = q.Kq;
    a.mult[i] = mult[i];
    a.src[i] = src5[i];
    a.src_stride[i] = stride5[i];
    a.norm_off[i] = (i == 0) ? 0 : a.norm_off[i - 1] + B * mult[i - 1];
  }
  if (scratch != nullptr && !prenormalised) {
    // norms once (one warp per key vector) so the transposing kernel can use small, numerous blocks
    key_norms_kernel<<<dim3((B * F + 7) / 8, 5), 256, 0, st>>>(B, D, a, scratch, staged);
    HMMC_CHECK_LAUNCH();
  }
  const bool have_norms = scratch != nullptr || prenormalised;
  const int dchunk = have_norms ? 64 : ENQ_DCHUNK;
  const int dchunks = (D + dchunk - 1) / dchunk;
  const int p_arg = device_ptr ? -1 : int(ptr_host), np_arg = device_ptr ? 0 : int((ptr_host + B) % K);
  bool vec_ok = have_norms && (D % 2) == 0;
  for (int i = 0; i < 5 && vec_ok; ++i) {
    vec_ok = (reinterpret_cast<uintptr_t>(a.src[i]) % 8 == 0) && (a.src_stride[i] % 2 == 0) && (a.Kq[i] % 2 == 0) &&
             (reinterpret_cast<uintptr_t>(a.dk[i]) % 8 == 0) && (reinterpret_cast<uintptr_t>(a.pack_kd[i]) % 4 == 0) &&
             (reinterpret_cast<uintptr_t>(a.pack_dk[i]) % 4 == 0);
  }
  if (vec_ok) {
    dim3 grid((B * F + 63) / 64, (D + 63) / 64, 5);
    enqueue_vec_kernel<<<grid, 256, 0, st>>>(B, D, a, scratch, queue_ptr, p_arg, np_arg, staged);
  } else {
    dim3 grid((B * F + 31) / 32, dchunks, 5);
    enqueue_kernel<32><<<grid, 256, 0, st>>>(B, D, dchunk, a, scratch, queue_ptr, p_arg, np_arg, staged);
  }
  HMMC_CHECK_LAUNCH();
  if (device_ptr || staged != nullptr) {
    advance_ptr_kernel<<<1, 1, 0, st>>>(queue_ptr, B, K, device_ptr ? 1 : 0, staged);
    HMMC_CHECK_LAUNCH();
  }
  return HMMC_OK;
}

int hmmc_enqueue_norm(const float* gathered, int W, int b, int F, int D, const hmmc_queue* queues5, int64_t* queue_ptr,
                      int64_t ptr_host, int K, float* scratch, int32_t* staged, const int32_t* slot_epoch,
                      int64_t slot_stride, int prenormalised, void* stream) {
  HMMC_REQUIRE(gathered != nullptr, "enqueue: null gathered buffer");
  const int64_t row = int64_t(3 + 2 * F) * D;
  const float* src[5] = {gathered, gathered + D, gathered + 2 * D, gathered + 3 * D, gathered + 3 * D + int64_t(F) * D};
  const int64_t stride[5] = {row, row, row, row, row};
  HMMC_REQUIRE(staged == nullptr || ptr_host < 0, "enqueue: the staged mark needs the device-side queue pointer (ptr_host < 0)");
  HMMC_REQUIRE(slot_epoch == nullptr || slot_stride >= int64_t(W) * b * row, "enqueue: slot stride smaller than a slot");
  return enqueue_common(src, stride, W * b, F, D, queues5, queue_ptr, ptr_host, K, scratch, staged, slot_epoch,
                        slot_stride, prenormalised, static_cast<cudaStream_t>(stream));
}

int hmmc_enqueue_norm_direct(const float* v_k, const float* tag_k, const float* title_k, const float* frame_fea_k,
                             const float* frame_proj_k, int B, int F, int D, const hmmc_queue* queues5,
                             int64_t* queue_ptr, int64_t ptr_host, int K, float* scratch, void* stream) {
  const float* src[5] = {v_k, tag_k, title_k, frame_fea_k, frame_proj_k};
  const int64_t stride[5] = {D, D, D, int64_t(F) * D, int64_t(F) * D};
  return enqueue_common(src, stride, B, F, D, queues5, queue_ptr, ptr_host, K, scratch, nullptr, nullptr, 0, 0,
                        static_cast<cudaStream_t>(stream));
}

int hmmc_scale_tensors(const uint64_t* ptrs_host, const int64_t* numels_host, int n, const float* scale, void* stream) {
  HMMC_REQUIRE(ptrs_host && numels_host && scale && n >= 1 && n <= 8, "scale_tensors: bad arguments");
  ScaleArgs a;
  a.n = n;
  int64_t mx = 0;
  for (int i = 0; i < n; ++i) {
    a.ptrs[i] = reinterpret_cast<float*>(ptrs_host[i]);
    a.numels[i] = numels_host[i];
    mx = numels_host[i] > mx ? numels_host[i] : mx;
  }
  if (mx <= 0) return HMMC_OK;
  // two blocks per SM over all tensors: in the usual case (upstream gradient 1) every block reads the scalar and
  // exits, and a grid of thousands of blocks costs microseconds just to drain
  int gx = int((mx / 4 + 255) / 256);
  const int cap = (2 * sm_count() + n - 1) / n;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  count_launch();
  HMMC_CHECK_CUDA(launch_pdl(scale_tensors_kernel, dim3(gx, n), dim3(256), 0, static_cast<cudaStream_t>(stream), a, scale));
  return HMMC_OK;
}

int hmmc_pack_rows(const uint64_t* src_ptrs_host, const int32_t* widths_host, int n, int64_t rows, float* dst,
                   int32_t* staged, int norm_dim, void* stream) {
  RowPackArgs a;
  int rc = build_rowpack(a, src_ptrs_host, widths_host, n);
  if (rc) return rc;
  if (rows <= 0) return HMMC_OK;
  if (norm_dim > 0) {
    for (int i = 0; i < n; ++i)
      HMMC_REQUIRE(widths_host[i] % norm_dim == 0, "pack_rows: width %d is not a multiple of the vector length %d",
                   widths_host[i], norm_dim);
    HMMC_REQUIRE(norm_dim <= 1024, "pack_rows: vectors longer than 1024 are not supported (got %d)", norm_dim);
    const int64_t vectors = rows * (a.total / norm_dim);
    rowpack_norm_kernel<<<unsigned((vectors + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, dst, norm_dim, rows, staged);
    HMMC_CHECK_LAUNCH();
    return HMMC_OK;
  }
  rowpack_kernel<true><<<dim3(unsigned(rows), a.n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, dst, rows, staged);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_unpack_rows(const float* src, const uint64_t* dst_ptrs_host, const int32_t* widths_host, int n, int64_t rows,
                     void* stream) {
  RowPackArgs a;
  int rc = build_rowpack(a, dst_ptrs_host, widths_host, n);
  if (rc) return rc;
  if (rows <= 0) return HMMC_OK;
  rowpack_kernel<false><<<dim3(unsigned(rows), a.n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, const_cast<float*>(src), rows,
                                                                                                   nullptr);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_peer_push_rows(const float* send, int64_t elems, const uint64_t* peer_bufs_host, const uint64_t* peer_flags_host,
                        int W, int rank, int64_t slot_stride, const int32_t* epoch, uint32_t* done_counter,
                        void* stream) {
  HMMC_REQUIRE(send && peer_bufs_host && peer_flags_host && epoch && done_counter, "peer_push: null argument");
  HMMC_REQUIRE(W >= 1 && W <= HMMC_MAX_PEERS && rank >= 0 && rank < W, "peer_push: world %d (max %d), rank %d", W,
               HMMC_MAX_PEERS, rank);
  HMMC_REQUIRE(elems > 0 && elems % 4 == 0 && slot_stride % 4 == 0 && slot_stride >= int64_t(W) * elems,
               "peer_push: %lld elements per rank, slot stride %lld", (long long)elems, (long long)slot_stride);
  HMMC_REQUIRE(reinterpret_cast<uintptr_t>(send) % 16 == 0, "peer_push: send buffer not 16-byte aligned");
  PeerPtrs bufs, flags;
  for (int i = 0; i < HMMC_MAX_PEERS; ++i) {
    bufs.p[i] = i < W ? peer_bufs_host[i] : 0;
    flags.p[i] = i < W ? peer_flags_host[i] : 0;
    HMMC_REQUIRE(i >= W || (bufs.p[i] % 16 == 0 && flags.p[i] != 0), "peer_push: bad peer pointer %d", i);
  }
  const int64_t n4 = elems / 4;
  // 32 blocks per destination.  A wider grid finishes the copy sooner (64 instead of 160 us for 52 MB beside the
  // momentum update) but the step gets slower, at 2 and at 8 ranks (0.451 -> 0.472 ms): the copy only has to be
  // done before the update is, and its blocks take slots and memory requests away from it
  const int gx = int(std::min<int64_t>((n4 + 1023) / 1024, 32));
  count_launch();
  peer_push_kernel<<<dim3(gx, W), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(send), n4, bufs, flags, W, rank, slot_stride, epoch, done_counter);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_peer_wait(const int32_t* my_flags, int W, int32_t* epoch, void* stream) {
  HMMC_REQUIRE(my_flags && epoch && W >= 1 && W <= HMMC_MAX_PEERS, "peer_wait: bad arguments");
  count_launch();
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(my_flags, W, epoch);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

}  // extern "C"

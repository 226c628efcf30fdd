// Pre-train (MoCo) head: InfoNCE against the negative queues, queue packing, enqueue,
// momentum EMA and the pack/unpack helpers of the key all-gather.
#include "common.cuh"
#include "umma_gemm.cuh"
#include <stdlib.h>

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

struct PrepArgs {
  const float* x[MAX_GROUPS];
  float* xhat[MAX_GROUPS];              // fp32 normalised copy (FP32 path) or nullptr
  __nv_bfloat16* packed[MAX_GROUPS];    // bf16 planes (tensor-core path) or nullptr
  int row_begin[MAX_GROUPS + 1];
  int n;
};

// one warp per row, all query tensors of the call in one launch (F.normalize, eps 1e-12)
__global__ void prep_rows_kernel(PrepArgs a, int D, int planes) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.row_begin[a.n]) return;
  int gi = 0;
#pragma unroll
  for (int i = 1; i < MAX_GROUPS; ++i)
    if (i < a.n && row >= a.row_begin[i]) gi = i;
  const int r = row - a.row_begin[gi];
  const float* xr = a.x[gi] + int64_t(r) * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = xr[d]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float n = fmaxf(sqrtf(ss), 1e-12f);
  float* xh = a.xhat[gi];
  __nv_bfloat16* pk = a.packed[gi];
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d] / n;
    if (xh != nullptr) xh[int64_t(r) * D + d] = v;
    if (pk != nullptr) {
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      pk[int64_t(r) * planes * D + d] = hi;
      if (planes == 2) pk[int64_t(r) * planes * D + D + d] = lo;
    }
  }
}

// ------------------------------------------------------------------ finish kernel
// One WARP per query row r = n*Fq + f of a query tensor.  For every (queue) contribution of
// that tensor it consumes the negatives' row sum S_r and U_r = sum_j e_rj Q_j, evaluates the
// positive terms selected by pos_mode, and writes the row's loss shares and dL/dq_r
// (SURVEY.md 8a').  A tensor that is the query of two losses (title_fea: VTM and FTM) gets
// both contributions here, so its gradient is written once.
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
};

template <int NE, int OCC>   // NE = elements per lane = D / 32 rounded up; OCC = blocks per SM to compile for
__global__ void __launch_bounds__(256, OCC)
infonce_finish_kernel(const __grid_constant__ FinishArgs a, int D, float invT, float cmax,
                      float* __restrict__ row_loss /* [3][total_rows] */) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int total_rows = a.row_begin[a.n];
  if (row >= total_rows) return;
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

  float loss_kind[3] = {0.f, 0.f, 0.f};
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
    S = warp_sum(S + S1);
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
      // g_hat += coef/T * (sum_t 1/Z_t) U_r ;  U_r = sum over the split-K partials (loads batched per split)
      const float wu = scale * sum_invZ;
      // two splits per iteration: 2*NE independent loads in flight per lane
      for (int sidx = 0; sidx < C.n_splits; sidx += 2) {
        const float* up0 = C.U_part + int64_t(sidx) * C.split_stride + int64_t(r) * D;
        const bool two = sidx + 1 < C.n_splits;
        const float* up1 = two ? up0 + C.split_stride : up0;
        float u0[NE], u1[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          const int d = lane + i * 32;
          u0[i] = (d < D) ? __ldg(up0 + d) : 0.f;
          u1[i] = (d < D && two) ? __ldg(up1 + d) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NE; ++i) g[i] = fmaf(wu, u0[i] + u1[i], g[i]);
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) row_loss[int64_t(k) * total_rows + row] = loss_kind[k];
  }
  if (G.dq == nullptr) return;
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

// per-kind sums in a fixed order (deterministic); kind_out[k] (+)= sum_r row_loss[k][r]
__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, int total_rows, float* __restrict__ kind_out,
                                   int accumulate) {
  __shared__ float red[32];
  for (int k = 0; k < 3; ++k) {
    const float* v = row_loss + int64_t(k) * total_rows;
    float acc = 0.f;
    for (int i = threadIdx.x; i < total_rows; i += 4 * blockDim.x) {
      const int i1 = i + blockDim.x, i2 = i + 2 * blockDim.x, i3 = i + 3 * blockDim.x;
      const float a0 = v[i];
      const float a1 = i1 < total_rows ? v[i1] : 0.f;
      const float a2 = i2 < total_rows ? v[i2] : 0.f;
      const float a3 = i3 < total_rows ? v[i3] : 0.f;
      acc += (a0 + a1) + (a2 + a3);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) kind_out[k] = accumulate ? kind_out[k] + acc : acc;
    __syncthreads();
  }
}

__global__ void add_scalar_kernel(float* __restrict__ out, const float* __restrict__ kinds) {
  out[0] += kinds[0] + kinds[1] + kinds[2];
}
// losses_out = [total, FAM, VTM, FTM]; the slots arrive weighted (w * loss), report them unweighted too
__global__ void head_losses_kernel(float* __restrict__ out, const float* __restrict__ kinds, float w_fam, float w_vtm,
                                   float w_ftm, int use_frame_fea) {
  const float fam = kinds[0], vtm = kinds[1], ftm = use_frame_fea ? kinds[2] : 0.f;
  out[0] = fam + vtm + ftm;
  out[1] = (w_fam != 0.f) ? fam / w_fam : 0.f;
  out[2] = (w_vtm != 0.f) ? vtm / w_vtm : 0.f;
  out[3] = (w_ftm != 0.f) ? ftm / w_ftm : 0.f;
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
  int planes;
};

constexpr int ENQ_DCHUNK = 128;

// Block = CB consecutive queue columns x ENQ_DCHUNK embedding dims of one queue.  Phase 1: the
// warps compute max(||x||,1e-12) of the block's CB key vectors.  Phase 2: 32(d) x CB(col) tiles go
// through shared memory so the [D,Kq] layouts are written CB columns (CB*4 contiguous bytes) at a
// time.  CB = 32 for small batches (more blocks), 128 for large ones (longer DRAM bursts).
// ptr >= 0: the host-tracked pointer (block 0 stores new_ptr);  ptr < 0: read the pointer from
// queue_ptr[0] on the device (CUDA-graph replay: no host value can be baked in) and leave the
// advance to advance_ptr_kernel.
template <int CB>
__global__ void __launch_bounds__(256)
enqueue_kernel(int nsamples, int D, EnqueueArgs a, int64_t* __restrict__ queue_ptr, int ptr, int new_ptr) {
  __shared__ float nrm[CB];
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
  auto src_of = [&](int c) -> const float* {     // c = local column = sample*mult + f
    const int smp = c / mult, f = c - smp * mult;
    return a.src[qi] + int64_t(smp) * a.src_stride[qi] + int64_t(f) * D;
  };
  constexpr int CPW = CB / 8;                    // columns per warp
  for (int i = 0; i < CPW; i += 2) {             // two vectors at a time: more loads in flight
    const int ca = c0 + warp * CPW + i, cb = ca + 1;
    float sa = 0.f, sb = 0.f;
    const float* xa = (ca < ncols) ? src_of(ca) : nullptr;
    const float* xb = (cb < ncols && i + 1 < CPW) ? src_of(cb) : nullptr;
    for (int d = lane; d < D; d += 32) {
      const float va = xa ? xa[d] : 0.f;
      const float vb = xb ? xb[d] : 0.f;
      sa = fmaf(va, va, sa);
      sb = fmaf(vb, vb, sb);
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    if (lane == 0) {
      nrm[warp * CPW + i] = fmaxf(sqrtf(sa), 1e-12f);
      if (i + 1 < CPW) nrm[warp * CPW + i + 1] = fmaxf(sqrtf(sb), 1e-12f);
    }
  }
  __syncthreads();
  const int planes = a.planes;
  float* dk = a.dk[qi];
  __nv_bfloat16* pkd = a.pack_kd[qi];
  __nv_bfloat16* pdk = a.pack_dk[qi];
  const int dbeg = blockIdx.y * ENQ_DCHUNK;
  const int dend = min(dbeg + ENQ_DCHUNK, D);
  for (int d0 = dbeg; d0 < dend; d0 += 32) {
    // read: warp w handles columns w, w+8, ..; lane = d
#pragma unroll 4
    for (int cc = warp; cc < CB; cc += 8) {
      const int c = c0 + cc, d = d0 + lane;
      float v = 0.f;
      if (c < ncols && d < dend) {
        v = src_of(c)[d] / nrm[cc];
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

__global__ void advance_ptr_kernel(int64_t* queue_ptr, int B, int K) { queue_ptr[0] = (queue_ptr[0] + B) % K; }

static int enqueue_common(const float* const* src5, const int64_t* stride5, int B, int F, int D,
                          const hmmc_queue* queues5, int64_t* queue_ptr, int64_t ptr_host, int K, cudaStream_t st) {
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
  }
  const int dchunks = (D + ENQ_DCHUNK - 1) / ENQ_DCHUNK;
  const int p_arg = device_ptr ? -1 : int(ptr_host), np_arg = device_ptr ? 0 : int((ptr_host + B) % K);
  if (B * F >= 4096) {
    dim3 grid((B * F + 127) / 128, dchunks, 5);
    enqueue_kernel<128><<<grid, 256, 0, st>>>(B, D, a, queue_ptr, p_arg, np_arg);
  } else {
    dim3 grid((B * F + 31) / 32, dchunks, 5);
    enqueue_kernel<32><<<grid, 256, 0, st>>>(B, D, a, queue_ptr, p_arg, np_arg);
  }
  HMMC_CHECK_LAUNCH();
  if (device_ptr) {
    advance_ptr_kernel<<<1, 1, 0, st>>>(queue_ptr, B, K);
    HMMC_CHECK_LAUNCH();
  }
  return HMMC_OK;
}

int hmmc_enqueue_norm(const float* gathered, int W, int b, int F, int D, const hmmc_queue* queues5, int64_t* queue_ptr,
                      int64_t ptr_host, int K, void* stream) {
  HMMC_REQUIRE(gathered != nullptr, "enqueue: null gathered buffer");
  const int64_t row = int64_t(3 + 2 * F) * D;
  const float* src[5] = {gathered, gathered + D, gathered + 2 * D, gathered + 3 * D, gathered + 3 * D + int64_t(F) * D};
  const int64_t stride[5] = {row, row, row, row, row};
  return enqueue_common(src, stride, W * b, F, D, queues5, queue_ptr, ptr_host, K, static_cast<cudaStream_t>(stream));
}

int hmmc_enqueue_norm_direct(const float* v_k, const float* tag_k, const float* title_k, const float* frame_fea_k,
                             const float* frame_proj_k, int B, int F, int D, const hmmc_queue* queues5,
                             int64_t* queue_ptr, int64_t ptr_host, int K, void* stream) {
  const float* src[5] = {v_k, tag_k, title_k, frame_fea_k, frame_proj_k};
  const int64_t stride[5] = {D, D, D, int64_t(F) * D, int64_t(F) * D};
  return enqueue_common(src, stride, B, F, D, queues5, queue_ptr, ptr_host, K, static_cast<cudaStream_t>(stream));
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
  int gx = int((mx / 4 + 255) / 256);
  if (gx < 1) gx = 1;
  if (gx > 4 * sm_count()) gx = 4 * sm_count();
  scale_tensors_kernel<<<dim3(gx, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, scale);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_pack_rows(const uint64_t* src_ptrs_host, const int32_t* widths_host, int n, int64_t rows, float* dst,
                   void* stream) {
  RowPackArgs a;
  int rc = build_rowpack(a, src_ptrs_host, widths_host, n);
  if (rc) return rc;
  if (rows <= 0) return HMMC_OK;
  rowpack_kernel<true><<<dim3(unsigned(rows), a.n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, dst, rows);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_unpack_rows(const float* src, const uint64_t* dst_ptrs_host, const int32_t* widths_host, int n, int64_t rows,
                     void* stream) {
  RowPackArgs a;
  int rc = build_rowpack(a, dst_ptrs_host, widths_host, n);
  if (rc) return rc;
  if (rows <= 0) return HMMC_OK;
  rowpack_kernel<false><<<dim3(unsigned(rows), a.n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, const_cast<float*>(src), rows);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

}  // extern "C"

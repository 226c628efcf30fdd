// Pre-train (MoCo) head: InfoNCE against the negative queues, queue packing, enqueue,
// momentum EMA and the pack/unpack helpers of the key all-gather.
#include "common.cuh"
#include "umma_gemm.cuh"

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

// ------------------------------------------------------------------ finish kernel
// One block per query row r = n*Fq + f.  Consumes the negatives' row sum S_r and
// U_r = sum_j e_rj Q_j, evaluates the positive terms selected by pos_mode and writes the
// per-row loss and dL/dq_r (SURVEY.md 8a').
constexpr int FIN_THREADS = 128;
constexpr int FIN_MAXE = 8;   // D <= FIN_THREADS * FIN_MAXE

__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { red[w] = a; red[8 + w] = b; }
  __syncthreads();
  float ta = 0.f, tb = 0.f;
#pragma unroll
  for (int i = 0; i < FIN_THREADS / 32; ++i) { ta += red[i]; tb += red[8 + i]; }
  a = ta;
  b = tb;
}

__global__ void __launch_bounds__(FIN_THREADS)
infonce_finish_kernel(const float* __restrict__ q, const float* __restrict__ keys, int pos_mode, int b, int Fq, int Fk,
                      int D, const float* __restrict__ rowsum_part, int n_parts, int R,
                      const float* __restrict__ U_part, int n_splits, int64_t split_stride, float invT, float cmax,
                      float coef, float* __restrict__ dq, float* __restrict__ row_loss) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int n = r / Fq, f = r - n * Fq;
  const int tid = threadIdx.x;

  float qv[FIN_MAXE], g[FIN_MAXE];
  float ss = 0.f, dummy = 0.f;
#pragma unroll
  for (int i = 0; i < FIN_MAXE; ++i) {
    const int d = tid + i * FIN_THREADS;
    qv[i] = (d < D) ? q[int64_t(r) * D + d] : 0.f;
    g[i] = 0.f;
    ss = fmaf(qv[i], qv[i], ss);
  }
  block_sum2(ss, dummy, red);
  const float nq_raw = sqrtf(ss);
  const float nq = fmaxf(nq_raw, 1e-12f);
#pragma unroll
  for (int i = 0; i < FIN_MAXE; ++i) qv[i] = qv[i] / nq;   // q_hat

  // S_r
  float S = 0.f;
  for (int p = tid; p < n_parts; p += FIN_THREADS) S += rowsum_part[int64_t(p) * R + r];
  dummy = 0.f;
  block_sum2(S, dummy, red);

  // positive terms
  int nterm, kbase, kstep;
  if (pos_mode == HMMC_POS_PAIR) { nterm = 1; kbase = r; kstep = 0; }
  else if (pos_mode == HMMC_POS_FRAME_NEIGHBOUR) { nterm = 2; kbase = 0; kstep = 0; }
  else if (pos_mode == HMMC_POS_ONE_TO_FRAMES) { nterm = Fk; kbase = n * Fk; kstep = 1; }
  else { nterm = 1; kbase = n; kstep = 0; }

  float loss = 0.f, sum_invZ = 0.f;
  for (int t = 0; t < nterm; ++t) {
    int kr = kbase + t * kstep;
    if (pos_mode == HMMC_POS_FRAME_NEIGHBOUR) {
      const int fk = (t == 0) ? f + 1 : f - 1;      // pairs (i, i+1) and (i+1, i) of frame_self_loss
      if (fk < 0 || fk >= Fk) continue;             // block-uniform
      kr = n * Fk + fk;
    }
    float kv[FIN_MAXE];
    float kk = 0.f, qk = 0.f;
#pragma unroll
    for (int i = 0; i < FIN_MAXE; ++i) {
      const int d = tid + i * FIN_THREADS;
      kv[i] = (d < D) ? keys[int64_t(kr) * D + d] : 0.f;
      kk = fmaf(kv[i], kv[i], kk);
      qk = fmaf(kv[i], qv[i], qk);
    }
    block_sum2(kk, qk, red);
    const float nk = fmaxf(sqrtf(kk), 1e-12f);
    const float lpos = (qk / nk) * invT;
    const float epos = expf(lpos - cmax);
    const float Z = epos + S;
    loss += logf(Z) + cmax - lpos;
    const float ppos = epos / Z;
    sum_invZ += 1.0f / Z;
    const float w = (ppos - 1.0f) / nk;
#pragma unroll
    for (int i = 0; i < FIN_MAXE; ++i) g[i] = fmaf(w, kv[i], g[i]);
  }
  if (tid == 0) row_loss[r] = coef * loss;
  if (dq == nullptr) return;

  // g_hat = coef/T * (sum_t (p+ - 1) k_hat_t + (sum_t 1/Z_t) U_r);   dq = (g_hat - q_hat (q_hat.g_hat)) / ||q||
  const float scale = coef * invT;
  float qg = 0.f;
#pragma unroll
  for (int i = 0; i < FIN_MAXE; ++i) {
    const int d = tid + i * FIN_THREADS;
    float u = 0.f;
    if (d < D)
      for (int sidx = 0; sidx < n_splits; ++sidx) u += U_part[int64_t(sidx) * split_stride + int64_t(r) * D + d];
    g[i] = scale * fmaf(sum_invZ, u, g[i]);
    qg = fmaf(qv[i], g[i], qg);
  }
  dummy = 0.f;
  block_sum2(qg, dummy, red);
  const bool clamped = nq_raw < 1e-12f;   // F.normalize clamps: q_hat = q/eps is then linear in q
#pragma unroll
  for (int i = 0; i < FIN_MAXE; ++i) {
    const int d = tid + i * FIN_THREADS;
    if (d < D) dq[int64_t(r) * D + d] = clamped ? g[i] / nq : (g[i] - qv[i] * qg) / nq;
  }
}

// loss_out[0] += sum_r row_loss[r], fixed summation order (deterministic)
__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, int R, float* __restrict__ loss_out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) acc += row_loss[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss_out[0] += acc;
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
  int Kq[5];        // columns of each queue
  int mult[5];      // columns per sample: 1 or F
  int src_off[5];   // element offset of the queue's block inside one gathered row
  int planes;
};

// Block = 32 consecutive queue columns of one queue.  Phase 1: one warp per 4 columns computes
// 1/max(||x||,1e-12).  Phase 2: 32(d) x 32(col) tiles go through shared memory so the [D,Kq]
// layouts are written 32 columns (128 B) at a time.
__global__ void __launch_bounds__(256)
enqueue_kernel(const float* __restrict__ gathered, int nsamples, int F, int D, int row_elems, EnqueueArgs a,
               const int64_t* __restrict__ queue_ptr) {
  __shared__ float inv[32];
  __shared__ float tile[32][33];
  const int qi = blockIdx.y;
  const int mult = a.mult[qi];
  const int ncols = nsamples * mult;
  const int c0 = blockIdx.x * 32;
  if (c0 >= ncols) return;
  const int ptr = int(queue_ptr[0]);
  const int Kq = a.Kq[qi];
  const int col_base = ptr * mult;       // first destination column
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto src_of = [&](int c) -> const float* {     // c = local column = sample*mult + f
    const int smp = c / mult, f = c - smp * mult;
    return gathered + int64_t(smp) * row_elems + a.src_off[qi] + int64_t(f) * D;
  };
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + warp * 4 + i;
    float ss = 0.f;
    if (c < ncols) {
      const float* x = src_of(c);
      for (int d = lane; d < D; d += 32) { const float v = x[d]; ss = fmaf(v, v, ss); }
    }
    ss = warp_sum(ss);
    if (lane == 0) inv[warp * 4 + i] = fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  const int planes = a.planes;
  float* dk = a.dk[qi];
  __nv_bfloat16* pkd = a.pack_kd[qi];
  __nv_bfloat16* pdk = a.pack_dk[qi];
  for (int d0 = 0; d0 < D; d0 += 32) {
    // read: warp w handles columns w, w+8, ..; lane = d
    for (int cc = warp; cc < 32; cc += 8) {
      const int c = c0 + cc, d = d0 + lane;
      float v = 0.f;
      if (c < ncols && d < D) {
        v = src_of(c)[d] / inv[cc];
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
    // write transposed: warp w handles d = w, w+8, ..; lane = column
    for (int dd = warp; dd < 32; dd += 8) {
      const int d = d0 + dd, c = c0 + lane;
      if (d < D && c < ncols) {
        const float v = tile[lane][dd];
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
    __syncthreads();
  }
}

__global__ void advance_ptr_kernel(int64_t* queue_ptr, int B, int K) {
  queue_ptr[0] = (queue_ptr[0] + B) % K;
}

// ------------------------------------------------------------------ pack / unpack rows
struct RowPackArgs {
  uint64_t ptrs[8];
  int widths[8];
  int offs[8];
  int n;
  int total;
};

template <bool PACK>
__global__ void rowpack_kernel(RowPackArgs a, float* __restrict__ packed, int64_t rows) {
  const int64_t row = blockIdx.x;
  float* prow = packed + row * a.total;
  for (int t = 0; t < a.n; ++t) {
    float* x = reinterpret_cast<float*>(a.ptrs[t]) + row * a.widths[t];
    float* y = prow + a.offs[t];
    for (int i = threadIdx.x; i < a.widths[t]; i += blockDim.x) {
      if (PACK) y[i] = x[i]; else x[i] = y[i];
    }
  }
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

static int infonce_bn(int Kq) { return (Kq % 256 == 0) ? 256 : 128; }

struct InfoNCEPlan {
  int planes, bn1, nparts, splits2, bn2;
};
static InfoNCEPlan infonce_plan(int64_t R, int D, int Kq, int prec) {
  InfoNCEPlan p;
  p.planes = planes_of(prec);
  if (prec == HMMC_PREC_FP32) {
    p.bn1 = p.bn2 = 0;
    p.nparts = 1;
    p.splits2 = 1;
  } else {
    p.bn1 = infonce_bn(Kq);
    p.nparts = (Kq + p.bn1 - 1) / p.bn1;
    p.bn2 = (D % 256 == 0) ? 256 : 128;
    const int nseg = (p.planes == 2) ? 3 : 1;
    p.splits2 = pick_splits(int(R), D, p.bn2, nseg * (Kq / UMMA_BK));
  }
  return p;
}

struct InfoNCEWs {
  float* qhat;
  __nv_bfloat16* qpack;
  float* rowsum_part;
  void* E;
  float* U_part;
  float* row_loss;
};
static void infonce_carve(Workspace& ws, InfoNCEWs& w, const InfoNCEPlan& p, int64_t R, int D, int Kq, int prec) {
  w.row_loss = ws.take<float>(size_t(R));
  w.rowsum_part = ws.take<float>(size_t(p.nparts) * R);
  w.U_part = ws.take<float>(size_t(p.splits2) * R * D);
  if (prec == HMMC_PREC_FP32) {
    w.qhat = ws.take<float>(size_t(R) * D);
    w.qpack = nullptr;
    w.E = ws.take<float>(size_t(R) * Kq);
  } else {
    w.qhat = nullptr;
    w.qpack = ws.take<__nv_bfloat16>(size_t(R) * p.planes * D);
    w.E = ws.take<__nv_bfloat16>(size_t(R) * p.planes * Kq);
  }
}

size_t hmmc_infonce_workspace_bytes(int64_t R, int D, int Kq, int prec) {
  Workspace ws(nullptr, 0);
  InfoNCEWs w;
  infonce_carve(ws, w, infonce_plan(R, D, Kq, prec), R, D, Kq, prec);
  return ws.used + 256;
}

int hmmc_infonce_queue_fwd_bwd(const float* q, const float* keys, int pos_mode, int b, int Fq, int Fk, int D,
                               const hmmc_queue* queue, float temperature, float weight, int prec, float* loss_out,
                               float* dq, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_REQUIRE(q && keys && queue && loss_out, "infonce: null argument");
  HMMC_REQUIRE(b > 0 && Fq > 0 && Fk > 0 && D > 0, "infonce: bad sizes b=%d Fq=%d Fk=%d D=%d", b, Fq, Fk, D);
  HMMC_REQUIRE(D <= FIN_THREADS * FIN_MAXE, "infonce: D=%d exceeds the supported %d", D, FIN_THREADS * FIN_MAXE);
  HMMC_REQUIRE(queue->D == D, "infonce: queue D=%d but embeddings D=%d", queue->D, D);
  HMMC_REQUIRE(pos_mode >= 0 && pos_mode <= 3, "infonce: unknown pos_mode %d", pos_mode);
  HMMC_REQUIRE(prec >= 0 && prec <= 2, "infonce: unknown precision %d", prec);
  if (pos_mode == HMMC_POS_PAIR) HMMC_REQUIRE(Fq == Fk, "infonce: PAIR needs Fq == Fk");
  if (pos_mode == HMMC_POS_FRAME_NEIGHBOUR) HMMC_REQUIRE(Fq == Fk && Fq >= 2, "infonce: FRAME_NEIGHBOUR needs Fq == Fk >= 2");
  if (pos_mode == HMMC_POS_ONE_TO_FRAMES) HMMC_REQUIRE(Fq == 1, "infonce: ONE_TO_FRAMES needs Fq == 1");
  if (pos_mode == HMMC_POS_FRAMES_TO_ONE) HMMC_REQUIRE(Fk == 1, "infonce: FRAMES_TO_ONE needs Fk == 1");
  // constant-max log-sum-exp: all logits lie in [-1/T, 1/T]; exp(-2/T) must stay a normal fp32
  HMMC_REQUIRE(temperature >= 0.025f, "infonce: temperature %g < 0.025 is not supported by the constant-max LSE", temperature);
  const int Kq = queue->Kq;
  const int64_t R = int64_t(b) * Fq;
  const InfoNCEPlan plan = infonce_plan(R, D, Kq, prec);
  Workspace ws(workspace, workspace_bytes);
  InfoNCEWs w;
  infonce_carve(ws, w, plan, R, D, Kq, prec);
  if (!ws.ok()) {
    set_error("infonce: workspace too small: need %zu bytes, got %zu", ws.used, workspace_bytes);
    return HMMC_ERR_WORKSPACE;
  }
  const float invT = 1.0f / temperature;
  const float cmax = invT;
  const bool need_grad = dq != nullptr;
  int rc;
  if (prec == HMMC_PREC_FP32) {
    rc = rownorm_pack(q, R, D, D, 1e-12f, 1, w.qhat, nullptr, nullptr, 0, st);
    if (rc) return rc;
    float* S = static_cast<float*>(w.E);
    rc = gemm_f32(w.qhat, D, 1, queue->dk, 1, Kq, S, Kq, int(R), Kq, D, 1.0f, st);
    if (rc) return rc;
    exp_rowsum_kernel<<<unsigned(R), 256, 0, st>>>(S, Kq, Kq, invT, cmax, w.rowsum_part);
    HMMC_CHECK_LAUNCH();
    if (need_grad) {
      rc = gemm_f32(S, Kq, 1, queue->dk, Kq, 1, w.U_part, D, int(R), D, Kq, 1.0f, st);
      if (rc) return rc;
    }
  } else {
    HMMC_REQUIRE(queue->pack_kd && queue->pack_dk, "infonce: queue has no packed operands (call hmmc_queue_pack)");
    HMMC_REQUIRE(queue->planes == plan.planes, "infonce: queue packed with %d planes, precision needs %d", queue->planes, plan.planes);
    HMMC_REQUIRE(D % UMMA_BK == 0 && Kq % 128 == 0, "infonce: tensor-core path needs D %% 64 == 0 and Kq %% 128 == 0 (D=%d Kq=%d)", D, Kq);
    rc = rownorm_pack(q, R, D, D, 1e-12f, plan.planes, nullptr, nullptr, w.qpack, int64_t(plan.planes) * D, st);
    if (rc) return rc;
    const float LOG2E = 1.4426950408889634f;
    EpiInfoNCE::Params ep;
    ep.a2 = invT * LOG2E;
    ep.c2 = cmax * LOG2E;
    ep.rowsum_part = w.rowsum_part;
    ep.E = need_grad ? static_cast<__nv_bfloat16*>(w.E) : nullptr;
    ep.ldE = int64_t(plan.planes) * Kq;
    ep.e_planes = plan.planes;
    if (plan.bn1 == 256)
      rc = launch_umma_gemm<256, EpiInfoNCE>(w.qpack, int64_t(plan.planes) * D, queue->pack_kd, int64_t(plan.planes) * D,
                                             int(R), Kq, D, plan.planes, 1, ep, st);
    else
      rc = launch_umma_gemm<128, EpiInfoNCE>(w.qpack, int64_t(plan.planes) * D, queue->pack_kd, int64_t(plan.planes) * D,
                                             int(R), Kq, D, plan.planes, 1, ep, st);
    if (rc) return rc;
    if (need_grad) {
      rc = umma_gemm_store(w.E, int64_t(plan.planes) * Kq, queue->pack_dk, int64_t(plan.planes) * Kq, w.U_part, D,
                           R * D, int(R), D, Kq, plan.planes, plan.splits2, 1.0f, st);
      if (rc) return rc;
    }
  }
  // how many split-K partials umma_gemm_store really produced (mirrors launch_umma_gemm)
  int n_splits = 1;
  if (prec != HMMC_PREC_FP32) {
    const int total_kb = ((plan.planes == 2) ? 3 : 1) * (Kq / UMMA_BK);
    int sp = plan.splits2 < 1 ? 1 : (plan.splits2 > total_kb ? total_kb : plan.splits2);
    const int per = (total_kb + sp - 1) / sp;
    n_splits = (total_kb + per - 1) / per;
  }
  infonce_finish_kernel<<<unsigned(R), FIN_THREADS, 0, st>>>(q, keys, pos_mode, b, Fq, Fk, D, w.rowsum_part, plan.nparts,
                                                           int(R), w.U_part, n_splits, R * D, invT, cmax,
                                                           weight / float(b), dq, w.row_loss);
  HMMC_CHECK_LAUNCH();
  loss_reduce_kernel<<<1, 256, 0, st>>>(w.row_loss, int(R), loss_out);
  HMMC_CHECK_LAUNCH();
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

int hmmc_enqueue_norm(const float* gathered, int W, int b, int F, int D, const hmmc_queue* queues5, int64_t* queue_ptr,
                      int64_t ptr_host, int K, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_REQUIRE(gathered && queues5 && queue_ptr, "enqueue: null argument");
  const int B = W * b;
  // the reference's slice assignment raises when the batch does not fit (modules/modeling.py:273-280)
  HMMC_REQUIRE(ptr_host >= 0 && ptr_host + B <= K, "enqueue: ptr %lld + batch %d exceeds queue size %d",
               (long long)ptr_host, B, K);
  EnqueueArgs a;
  const int mult[5] = {1, 1, 1, F, F};
  const int off[5] = {0, D, 2 * D, 3 * D, 3 * D + F * D};
  a.planes = queues5[0].planes;
  for (int i = 0; i < 5; ++i) {
    const hmmc_queue& q = queues5[i];
    HMMC_REQUIRE(q.dk != nullptr && q.D == D && q.Kq == K * mult[i], "enqueue: queue %d has shape [%d,%d], expected [%d,%d]",
                 i, q.D, q.Kq, D, K * mult[i]);
    HMMC_REQUIRE(q.planes == a.planes, "enqueue: queues disagree on planes");
    a.dk[i] = q.dk;
    a.pack_kd[i] = static_cast<__nv_bfloat16*>(q.pack_kd);
    a.pack_dk[i] = static_cast<__nv_bfloat16*>(q.pack_dk);
    a.Kq[i] = q.Kq;
    a.mult[i] = mult[i];
    a.src_off[i] = off[i];
  }
  const int row_elems = (3 + 2 * F) * D;
  dim3 grid((B * F + 31) / 32, 5);
  enqueue_kernel<<<grid, 256, 0, st>>>(gathered, B, F, D, row_elems, a, queue_ptr);
  HMMC_CHECK_LAUNCH();
  advance_ptr_kernel<<<1, 1, 0, st>>>(queue_ptr, B, K);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_pack_rows(const uint64_t* src_ptrs_host, const int32_t* widths_host, int n, int64_t rows, float* dst,
                   void* stream) {
  RowPackArgs a;
  int rc = build_rowpack(a, src_ptrs_host, widths_host, n);
  if (rc) return rc;
  if (rows <= 0) return HMMC_OK;
  rowpack_kernel<true><<<unsigned(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, dst, rows);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_unpack_rows(const float* src, const uint64_t* dst_ptrs_host, const int32_t* widths_host, int n, int64_t rows,
                     void* stream) {
  RowPackArgs a;
  int rc = build_rowpack(a, dst_ptrs_host, widths_host, n);
  if (rc) return rc;
  if (rows <= 0) return HMMC_OK;
  rowpack_kernel<false><<<unsigned(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, const_cast<float*>(src), rows);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

}  // extern "C"

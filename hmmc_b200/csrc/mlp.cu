// Projector / predictor MLP of the pre-train model:
//   MLP.forward  modules/modeling.py:788-807, num_layers = 2:
//     Linear(Din, Dh) -> (Sync)BatchNorm1d(Dh) -> ReLU -> Linear(Dh, Dout)
//   applied to the [b*F, 512] frame features right before the head (modeling.py:355-377).
// Forward and backward on the tcgen05 GEMM engine.  Batch normalisation needs the statistics of ALL
// ranks' rows (the reference converts the MLPs to SyncBatchNorm, modeling.py:127-129), so each
// direction is split in two phases around one small exchange the host performs:
//   fwd_a: H = X W1^T, local column sums (sum h, sum h^2)            -> all-reduce 2*Dh doubles
//   fwd_b: mean / invstd, running statistics, A = relu(BN(H)) written straight as bf16 operand
//          packs, Y = A W2^T + b2
//   bwd_a: dA = dY W2, dW2 = dY^T A, db2, dZ = dA * [z > 0], local (sum dZ, sum dZ*xhat) -> all-reduce
//   bwd_b: dgamma / dbeta (local sums), dH through the batch statistics, dW1 = dH^T X, db1, dX = dH W1
// The first Linear's bias cancels inside a training-mode BatchNorm, so H is kept without it and the
// bias enters only the running mean (and the eval-mode path).
#include "common.cuh"

namespace hmmc {

// ------------------------------------------------------------------ column reductions
// Column sums over the rows of an [M, N] matrix, two levels so that the grid fills the machine:
// block (x, y) = 32 columns x one slab of COL_SLAB rows -> partial sums part[y][c]; a second, tiny
// kernel adds the slabs in order (double).  Deterministic, no atomics.
//   MODE 0: (sum h, sum h^2)              forward statistics of H
//   MODE 1: (sum v)                       db2 = column sums of dY
//   MODE 2: dZ = dA * [z > 0] in place, (sum dZ, sum dZ * xhat)       backward sums
constexpr int COL_SLAB = 128;
template <int MODE>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(float* __restrict__ A, const float* __restrict__ H, int M, int N, const float* __restrict__ mean,
                      const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                      double* __restrict__ part0, double* __restrict__ part1) {
  __shared__ double r0[8][32], r1[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int row_begin = blockIdx.y * COL_SLAB, row_end = min(row_begin + COL_SLAB, M);
  // fp64 accumulators: the variance is a difference of two sums (sum h^2 / n - mean^2)
  double a0 = 0.0, a1 = 0.0;
  if (c < N) {
    float mu = 0.f, is = 0.f, g = 0.f, b = 0.f;
    if (MODE == 2) { mu = mean[c]; is = invstd[c]; g = gamma[c]; b = beta[c]; }
#pragma unroll 4
    for (int r = row_begin + warp; r < row_end; r += 8) {
      const int64_t o = int64_t(r) * N + c;
      if (MODE == 2) {
        const float xh = (H[o] - mu) * is;
        const float dz = (fmaf(xh, g, b) > 0.f) ? A[o] : 0.f;
        A[o] = dz;
        a0 += double(dz);
        a1 += double(dz) * double(xh);
      } else {
        const double v = double(A[o]);
        a0 += v;
        if (MODE == 0) a1 += v * v;
      }
    }
  }
  r0[warp][lane] = a0;
  r1[warp][lane] = a1;
  __syncthreads();
  if (warp == 0 && c < N) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { t0 += r0[w][lane]; t1 += r1[w][lane]; }
    part0[int64_t(blockIdx.y) * N + c] = t0;
    if (MODE != 1) part1[int64_t(blockIdx.y) * N + c] = t1;
  }
}

// slabs added in order; results as doubles into up to two destinations (local copy + exchange buffer)
__global__ void colsum_finish_kernel(const double* __restrict__ part0, const double* __restrict__ part1, int n_slabs, int N,
                                     double* __restrict__ dst_a, double* __restrict__ dst_b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  double s0 = 0.0, s1 = 0.0;
  for (int y = 0; y < n_slabs; ++y) {
    s0 += part0[int64_t(y) * N + c];
    if (part1 != nullptr) s1 += part1[int64_t(y) * N + c];
  }
  dst_a[c] = s0;
  if (dst_b != nullptr) dst_b[c] = s0;
  if (part1 != nullptr) {
    dst_a[N + c] = s1;
    if (dst_b != nullptr) dst_b[N + c] = s1;
  }
}

// mean / invstd of the batch from the (global) sums; running statistics as nn.BatchNorm1d updates them
// (biased variance for the normalisation, unbiased for running_var, the Linear bias folded back in).
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int N, double count, float eps, float momentum,
                                   const float* __restrict__ b1, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const double m = stats[c] / count;
  double var = stats[N + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = float(m);
  invstd[c] = float(1.0 / sqrt(var + double(eps)));
  if (running_mean != nullptr) {
    const float bm = float(m) + (b1 ? b1[c] : 0.f);
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * bm;
  }
  if (running_var != nullptr) {
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * float(unb);
  }
}

// eval mode: normalise with the running statistics; H carries no bias, so mean_eff = running_mean - b1
__global__ void bn_eval_prepare_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                       const float* __restrict__ b1, int N, float eps, float* __restrict__ mean,
                                       float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  mean[c] = running_mean[c] - (b1 ? b1[c] : 0.f);
  invstd[c] = 1.0f / sqrtf(running_var[c] + eps);
}

// A = relu((h - mean) * invstd * gamma + beta), written only as the bf16 plane packs the GEMMs read:
// Ap [M, planes*N] and (for the backward) ATp [N, planes*M].  32x32 tiles through shared memory.
__global__ void __launch_bounds__(256)
bn_relu_pack_kernel(const float* __restrict__ H, int M, int N, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int planes, __nv_bfloat16* __restrict__ Ap, __nv_bfloat16* __restrict__ ATp) {
  __shared__ float tile[32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int c = c0 + lane;
  float mu = 0.f, is = 0.f, g = 0.f, b = 0.f;
  if (c < N) { mu = mean[c]; is = invstd[c]; g = gamma[c]; b = beta[c]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = warp * 4 + k, r = r0 + rr;
    float a = 0.f;
    if (r < M && c < N) {
      a = fmaxf(fmaf((H[int64_t(r) * N + c] - mu) * is, g, b), 0.f);
      __nv_bfloat16 hi, lo;
      split_bf16(a, hi, lo);
      const int64_t o = int64_t(r) * planes * N + c;
      Ap[o] = hi;
      if (planes == 2) Ap[o + N] = lo;
    }
    tile[rr][lane] = a;
  }
  if (ATp == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cc = warp * 4 + k, col = c0 + cc, r = r0 + lane;
    if (col < N && r < M) {
      __nv_bfloat16 hi, lo;
      split_bf16(tile[lane][cc], hi, lo);
      const int64_t o = int64_t(col) * planes * M + r;
      ATp[o] = hi;
      if (planes == 2) ATp[o + M] = lo;
    }
  }
}

// out = fixed-order sum of the split-K partials (+ bias).  The long contractions (K = Dh) run as
// several shorter accumulation chains: tcgen05 truncates when it accumulates, so the error grows with
// the chain length (2e-5 at K = 4096 in the bf16x3 mode, 5e-6 at 512), and the few output tiles of
// these GEMMs need the extra parallelism anyway.
__global__ void sum_partials_bias_kernel(const float* __restrict__ parts, int n_splits, int64_t stride, int64_t total,
                                         int N, const float* __restrict__ bias, float* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float acc = 0.f;
  for (int s0 = 0; s0 < n_splits; ++s0) acc += parts[int64_t(s0) * stride + i];
  out[i] = acc + (bias ? bias[i % N] : 0.f);
}

// Parameter gradients of the normalisation from the LOCAL sums (DDP reduces parameter gradients
// itself), and db1 = column sums of dH in closed form.
__global__ void bn_param_grads_kernel(const double* __restrict__ sums_local, const double* __restrict__ sums_global,
                                      const double* __restrict__ stats_local, int N, int M, double count,
                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                      const float* __restrict__ gamma, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, float* __restrict__ db1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  if (dbeta != nullptr) dbeta[c] = float(sums_local[c]);
  if (dgamma != nullptr) dgamma[c] = float(sums_local[N + c]);
  if (db1 != nullptr) {
    const double m1 = sums_global[c] / count, m2 = sums_global[N + c] / count;
    const double sum_xhat = (stats_local[c] - double(M) * double(mean[c])) * double(invstd[c]);
    db1[c] = float(double(gamma[c]) * double(invstd[c]) * (sums_local[c] - double(M) * m1 - m2 * sum_xhat));
  }
}

// dH = gamma * invstd * (dZ - mean(dZ) - xhat * mean(dZ * xhat)), means over ALL ranks' rows; written
// only as the bf16 packs of the two GEMMs that consume it: dHp [M, planes*N], dHTp [N, planes*M].
__global__ void __launch_bounds__(256)
bn_bwd_pack_kernel(const float* __restrict__ dZ, const float* __restrict__ H, int M, int N,
                   const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                   const double* __restrict__ sums_global, double count, int planes, __nv_bfloat16* __restrict__ dHp,
                   __nv_bfloat16* __restrict__ dHTp) {
  __shared__ float tile[32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int c = c0 + lane;
  float mu = 0.f, is = 0.f, gi = 0.f, m1 = 0.f, m2 = 0.f;
  if (c < N) {
    mu = mean[c]; is = invstd[c]; gi = gamma[c] * is;
    m1 = float(sums_global[c] / count);
    m2 = float(sums_global[N + c] / count);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = warp * 4 + k, r = r0 + rr;
    float dh = 0.f;
    if (r < M && c < N) {
      const int64_t o = int64_t(r) * N + c;
      const float xh = (H[o] - mu) * is;
      dh = gi * (dZ[o] - m1 - xh * m2);
      __nv_bfloat16 hi, lo;
      split_bf16(dh, hi, lo);
      const int64_t p = int64_t(r) * planes * N + c;
      dHp[p] = hi;
      if (planes == 2) dHp[p + N] = lo;
    }
    tile[rr][lane] = dh;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cc = warp * 4 + k, col = c0 + cc, r = r0 + lane;
    if (col < N && r < M) {
      __nv_bfloat16 hi, lo;
      split_bf16(tile[lane][cc], hi, lo);
      const int64_t p = int64_t(col) * planes * M + r;
      dHTp[p] = hi;
      if (planes == 2) dHTp[p + M] = lo;
    }
  }
}

__global__ void double_to_float_kernel(const double* __restrict__ src, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = float(src[i]);
}

// ------------------------------------------------------------------ visual encoder tail
// VisualEncoder.forward, modules/module_cross.py:207-213:
//   h = temporal_out + original   (residual; original alone without the temporal transformer)
//   visual_output[b] = mean_f ( h[b,f] / ||h[b,f]|| )
// One block per sample, one warp per frame row; the F normalised rows meet in shared memory and are
// averaged in frame order.
__global__ void visual_tail_fwd_kernel(const float* __restrict__ temporal, const float* __restrict__ original, int F,
                                       int D, float* __restrict__ out) {
  extern __shared__ float rows[];      // [F][D]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int f = warp; f < F; f += nw) {
    const int64_t o = (int64_t(b) * F + f) * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = original[o + d] + (temporal ? temporal[o + d] : 0.f);
      rows[f * D + d] = v;
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float n = sqrtf(ss);
    for (int d = lane; d < D; d += 32) rows[f * D + d] /= n;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc += rows[f * D + d];
    out[int64_t(b) * D + d] = acc / float(F);
  }
}

// dh[b,f] = (g[b] - hhat (hhat . g[b])) / (F ||h||): one warp per frame row; the same gradient reaches
// the temporal branch and the residual branch.
__global__ void visual_tail_bwd_kernel(const float* __restrict__ temporal, const float* __restrict__ original,
                                       const float* __restrict__ dout, int64_t rows, int F, int D,
                                       float* __restrict__ dh) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t o = r * D;
  const float* g = dout + (r / F) * D;
  float ss = 0.f, hg = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = original[o + d] + (temporal ? temporal[o + d] : 0.f);
    ss = fmaf(v, v, ss);
    hg = fmaf(v, g[d], hg);
  }
  ss = warp_sum(ss);
  hg = warp_sum(hg);
  const float n = sqrtf(ss);
  const float proj = hg / (n * n), scale = 1.0f / (n * float(F));
  for (int d = lane; d < D; d += 32) {
    const float v = original[o + d] + (temporal ? temporal[o + d] : 0.f);
    dh[o + d] = (g[d] - v * proj) * scale;
  }
}

// ------------------------------------------------------------------ context layout
constexpr int MLP_MAX_SPLITS = 8;
static int mlp_splits(int K) {
  int sp = K / 512;
  return sp < 1 ? 1 : (sp > MLP_MAX_SPLITS ? MLP_MAX_SPLITS : sp);
}

struct MlpCtx {
  __nv_bfloat16 *Xp, *XTp, *W1p, *W1Tp, *W2p, *W2Tp, *Ap, *ATp, *dYp, *dYTp, *dHp, *dHTp;
  float *H, *dA, *mean, *invstd, *parts;
  double *cpart0, *cpart1;
  double *stats_local, *stats, *sums_local, *sums, *colsum;
};
static void mlp_carve(Workspace& ws, MlpCtx& c, int M, int Din, int Dh, int Dout, int P, bool need_grad) {
  const size_t p = size_t(P);
  c.stats_local = ws.take<double>(size_t(2) * Dh);
  c.stats = ws.take<double>(size_t(2) * Dh);
  c.mean = ws.take<float>(size_t(Dh));
  c.invstd = ws.take<float>(size_t(Dh));
  c.Xp = ws.take<__nv_bfloat16>(size_t(M) * p * Din);
  c.W1p = ws.take<__nv_bfloat16>(size_t(Dh) * p * Din);
  c.W2p = ws.take<__nv_bfloat16>(size_t(Dout) * p * Dh);
  c.H = ws.take<float>(size_t(M) * Dh);
  c.Ap = ws.take<__nv_bfloat16>(size_t(M) * p * Dh);
  c.parts = ws.take<float>(size_t(MLP_MAX_SPLITS) * M * (Dout > Din ? Dout : Din));
  const size_t slabs = size_t((M + COL_SLAB - 1) / COL_SLAB);
  c.cpart0 = ws.take<double>(slabs * (Dh > Dout ? Dh : Dout));
  c.cpart1 = ws.take<double>(slabs * (Dh > Dout ? Dh : Dout));
  c.XTp = c.W1Tp = c.W2Tp = c.ATp = c.dYp = c.dYTp = c.dHp = c.dHTp = nullptr;
  c.dA = nullptr;
  c.sums_local = c.sums = c.colsum = nullptr;
  if (need_grad) {
    c.XTp = ws.take<__nv_bfloat16>(size_t(Din) * p * M);
    c.W1Tp = ws.take<__nv_bfloat16>(size_t(Din) * p * Dh);
    c.W2Tp = ws.take<__nv_bfloat16>(size_t(Dh) * p * Dout);
    c.ATp = ws.take<__nv_bfloat16>(size_t(Dh) * p * M);
    c.dYp = ws.take<__nv_bfloat16>(size_t(M) * p * Dout);
    c.dYTp = ws.take<__nv_bfloat16>(size_t(Dout) * p * M);
    c.dA = ws.take<float>(size_t(M) * Dh);
    c.dHp = ws.take<__nv_bfloat16>(size_t(M) * p * Dh);
    c.dHTp = ws.take<__nv_bfloat16>(size_t(Dh) * p * M);
    c.sums_local = ws.take<double>(size_t(2) * Dh);
    c.sums = ws.take<double>(size_t(2) * Dh);
    c.colsum = ws.take<double>(size_t(Dout));
  }
}

static int mlp_check(int M, int Din, int Dh, int Dout, int prec, bool need_grad) {
  HMMC_REQUIRE(prec == HMMC_PREC_BF16 || prec == HMMC_PREC_BF16X3,
               "mlp: runs on the tensor cores only (precision bf16 or bf16x3)");
  HMMC_REQUIRE(M > 0 && Din > 0 && Dh > 0 && Dout > 0, "mlp: bad shape");
  HMMC_REQUIRE(Din % 64 == 0 && Dh % 64 == 0 && Dout % 64 == 0, "mlp: feature sizes must be multiples of 64 (%d, %d, %d)",
               Din, Dh, Dout);
  HMMC_REQUIRE(!need_grad || M % 64 == 0, "mlp: the backward needs a multiple of 64 rows (got %d)", M);
  return HMMC_OK;
}

}  // namespace hmmc

using namespace hmmc;

extern "C" {

size_t hmmc_mlp_ctx_bytes(int M, int Din, int Dh, int Dout, int prec, int need_grad) {
  Workspace ws(nullptr, 0);
  MlpCtx c;
  mlp_carve(ws, c, M, Din, Dh, Dout, planes_of(prec), need_grad != 0);
  return ws.used + 256;
}

int hmmc_mlp_fwd_a(const float* x, int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, int prec, int need_grad,
                   void* ctx, size_t ctx_bytes, double** stats_out, void* stream) {
  int rc;
  if ((rc = mlp_check(M, Din, Dh, Dout, prec, need_grad != 0))) return rc;
  HMMC_REQUIRE(x && p && p->W1 && p->W2 && p->gamma && p->beta, "mlp_fwd_a: null argument");
  const int P = planes_of(prec);
  Workspace ws(ctx, ctx_bytes);
  MlpCtx c;
  mlp_carve(ws, c, M, Din, Dh, Dout, P, need_grad != 0);
  HMMC_REQUIRE(ws.ok(), "mlp: context too small (%zu needed, %zu given)", ws.used, ctx_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((rc = pack_dual(x, M, Din, P, c.Xp, c.XTp, st))) return rc;
  if ((rc = pack_dual(p->W1, Dh, Din, P, c.W1p, c.W1Tp, st))) return rc;
  if ((rc = pack_dual(p->W2, Dout, Dh, P, c.W2p, c.W2Tp, st))) return rc;
  if ((rc = umma_gemm_store(c.Xp, int64_t(P) * Din, c.W1p, int64_t(P) * Din, c.H, Dh, 0, M, Dh, Din, P, 1, 1.0f, st))) return rc;
  const int slabs = (M + COL_SLAB - 1) / COL_SLAB;
  colsum_partial_kernel<0><<<dim3((Dh + 31) / 32, slabs), 256, 0, st>>>(c.H, nullptr, M, Dh, nullptr, nullptr, nullptr,
                                                                        nullptr, c.cpart0, c.cpart1);
  HMMC_CHECK_LAUNCH();
  colsum_finish_kernel<<<(Dh + 255) / 256, 256, 0, st>>>(c.cpart0, c.cpart1, slabs, Dh, c.stats_local, c.stats);
  HMMC_CHECK_LAUNCH();
  if (stats_out != nullptr) *stats_out = c.stats;
  return HMMC_OK;
}

int hmmc_mlp_fwd_b(int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, float eps, float momentum, double count,
                   int training, int prec, int need_grad, void* ctx, size_t ctx_bytes, float* y, void* stream) {
  int rc;
  if ((rc = mlp_check(M, Din, Dh, Dout, prec, need_grad != 0))) return rc;
  HMMC_REQUIRE(p && y && count >= 1.0, "mlp_fwd_b: bad argument");
  HMMC_REQUIRE(training || (p->running_mean && p->running_var), "mlp_fwd_b: eval mode needs the running statistics");
  const int P = planes_of(prec);
  Workspace ws(ctx, ctx_bytes);
  MlpCtx c;
  mlp_carve(ws, c, M, Din, Dh, Dout, P, need_grad != 0);
  HMMC_REQUIRE(ws.ok(), "mlp: context too small (%zu needed, %zu given)", ws.used, ctx_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (training)
    bn_finalize_kernel<<<(Dh + 255) / 256, 256, 0, st>>>(c.stats, Dh, count, eps, momentum, p->b1, c.mean, c.invstd,
                                                         p->running_mean, p->running_var);
  else
    bn_eval_prepare_kernel<<<(Dh + 255) / 256, 256, 0, st>>>(p->running_mean, p->running_var, p->b1, Dh, eps, c.mean,
                                                             c.invstd);
  HMMC_CHECK_LAUNCH();
  bn_relu_pack_kernel<<<dim3((Dh + 31) / 32, (M + 31) / 32), 256, 0, st>>>(c.H, M, Dh, c.mean, c.invstd, p->gamma, p->beta,
                                                                           P, c.Ap, c.ATp);
  HMMC_CHECK_LAUNCH();
  const int sp = mlp_splits(Dh);
  const int64_t total = int64_t(M) * Dout;
  if ((rc = umma_gemm_store(c.Ap, int64_t(P) * Dh, c.W2p, int64_t(P) * Dh, c.parts, Dout, total, M, Dout, Dh, P, sp, 1.0f, st))) return rc;
  sum_partials_bias_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(c.parts, umma_effective_splits(Dh, P, sp), total,
                                                                         total, Dout, p->b2, y);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_mlp_bwd_a(const float* dy, int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, int prec, void* ctx,
                   size_t ctx_bytes, float* dW2, float* db2, double** sums_out, void* stream) {
  int rc;
  if ((rc = mlp_check(M, Din, Dh, Dout, prec, true))) return rc;
  HMMC_REQUIRE(dy && p, "mlp_bwd_a: null argument");
  const int P = planes_of(prec);
  Workspace ws(ctx, ctx_bytes);
  MlpCtx c;
  mlp_carve(ws, c, M, Din, Dh, Dout, P, true);
  HMMC_REQUIRE(ws.ok(), "mlp: context too small (%zu needed, %zu given)", ws.used, ctx_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int slabs = (M + COL_SLAB - 1) / COL_SLAB;
  if ((rc = pack_dual(dy, M, Dout, P, c.dYp, c.dYTp, st))) return rc;
  // dA = dY W2 [M, Dh]  and  dW2 = dY^T A [Dout, Dh]  in one grouped launch
  StoreGemm g[2];
  int n = 0;
  g[n++] = StoreGemm{c.dYp, int64_t(P) * Dout, c.W2Tp, int64_t(P) * Dout, c.dA, Dh, 0, M, Dh, Dout, 1};
  if (dW2 != nullptr) g[n++] = StoreGemm{c.dYTp, int64_t(P) * M, c.ATp, int64_t(P) * M, dW2, Dh, 0, Dout, Dh, M, 1};
  if ((rc = umma_gemm_store_grouped(g, n, P, 1.0f, st))) return rc;
  if (db2 != nullptr) {
    colsum_partial_kernel<1><<<dim3((Dout + 31) / 32, slabs), 256, 0, st>>>(const_cast<float*>(dy), nullptr, M, Dout, nullptr,
                                                                            nullptr, nullptr, nullptr, c.cpart0, nullptr);
    HMMC_CHECK_LAUNCH();
    colsum_finish_kernel<<<(Dout + 255) / 256, 256, 0, st>>>(c.cpart0, nullptr, slabs, Dout, c.colsum, nullptr);
    HMMC_CHECK_LAUNCH();
    double_to_float_kernel<<<(Dout + 255) / 256, 256, 0, st>>>(c.colsum, Dout, db2);
    HMMC_CHECK_LAUNCH();
  }
  colsum_partial_kernel<2><<<dim3((Dh + 31) / 32, slabs), 256, 0, st>>>(c.dA, c.H, M, Dh, c.mean, c.invstd, p->gamma, p->beta,
                                                                        c.cpart0, c.cpart1);
  HMMC_CHECK_LAUNCH();
  colsum_finish_kernel<<<(Dh + 255) / 256, 256, 0, st>>>(c.cpart0, c.cpart1, slabs, Dh, c.sums_local, c.sums);
  HMMC_CHECK_LAUNCH();
  if (sums_out != nullptr) *sums_out = c.sums;
  return HMMC_OK;
}

int hmmc_mlp_bwd_b(int M, int Din, int Dh, int Dout, const hmmc_mlp_params* p, double count, int prec, void* ctx,
                   size_t ctx_bytes, float* dx, float* dW1, float* db1, float* dgamma, float* dbeta, void* stream) {
  int rc;
  if ((rc = mlp_check(M, Din, Dh, Dout, prec, true))) return rc;
  HMMC_REQUIRE(p && count >= 1.0, "mlp_bwd_b: bad argument");
  const int P = planes_of(prec);
  Workspace ws(ctx, ctx_bytes);
  MlpCtx c;
  mlp_carve(ws, c, M, Din, Dh, Dout, P, true);
  HMMC_REQUIRE(ws.ok(), "mlp: context too small (%zu needed, %zu given)", ws.used, ctx_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bn_param_grads_kernel<<<(Dh + 255) / 256, 256, 0, st>>>(c.sums_local, c.sums, c.stats_local, Dh, M, count, c.mean,
                                                          c.invstd, p->gamma, dgamma, dbeta, db1);
  HMMC_CHECK_LAUNCH();
  bn_bwd_pack_kernel<<<dim3((Dh + 31) / 32, (M + 31) / 32), 256, 0, st>>>(c.dA, c.H, M, Dh, c.mean, c.invstd, p->gamma,
                                                                          c.sums, count, P, c.dHp, c.dHTp);
  HMMC_CHECK_LAUNCH();
  // dW1 = dH^T X [Dh, Din]  and  dX = dH W1 [M, Din]  in one grouped launch
  StoreGemm g[2];
  int n = 0;
  if (dW1 != nullptr) g[n++] = StoreGemm{c.dHTp, int64_t(P) * M, c.XTp, int64_t(P) * M, dW1, Din, 0, Dh, Din, M, 1};
  const int sp = mlp_splits(Dh);
  const int64_t total = int64_t(M) * Din;
  if (dx != nullptr) g[n++] = StoreGemm{c.dHp, int64_t(P) * Dh, c.W1Tp, int64_t(P) * Dh, c.parts, Din, total, M, Din, Dh, sp};
  if ((rc = umma_gemm_store_grouped(g, n, P, 1.0f, st))) return rc;
  if (dx != nullptr) {
    sum_partials_bias_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(c.parts, umma_effective_splits(Dh, P, sp), total,
                                                                           total, Din, nullptr, dx);
    HMMC_CHECK_LAUNCH();
  }
  return HMMC_OK;
}

int hmmc_visual_tail_fwd(const float* temporal, const float* original, int B, int F, int D, float* out, void* stream) {
  HMMC_REQUIRE(original && out && B > 0 && F > 0 && D > 0, "visual_tail_fwd: bad arguments");
  const size_t smem = size_t(F) * D * sizeof(float);
  HMMC_REQUIRE(smem <= 200 * 1024, "visual_tail_fwd: F*D = %d floats exceed the shared-memory staging", F * D);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    HMMC_CHECK_CUDA(cudaFuncSetAttribute(visual_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    smem_set = smem;
  }
  const int warps = F < 16 ? F : 16;
  visual_tail_fwd_kernel<<<B, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(temporal, original, F, D, out);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_visual_tail_bwd(const float* temporal, const float* original, const float* dout, int B, int F, int D,
                         float* dhidden, void* stream) {
  HMMC_REQUIRE(original && dout && dhidden && B > 0 && F > 0 && D > 0, "visual_tail_bwd: bad arguments");
  const int64_t rows = int64_t(B) * F;
  visual_tail_bwd_kernel<<<unsigned((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(temporal, original, dout,
                                                                                                  rows, F, D, dhidden);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

}  // extern "C"

// Optimizer step right after the head's backward:
//   torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)   main_pretrain.py:277
//   BertAdam.step()                                           modules/optimization.py:103-168
// The reference walks the parameters in Python (~15 elementwise launches per tensor, a per-tensor
// clip inside the loop).  Here: one pass for the squared gradient norms, one small reduction that
// turns them into the global and per-tensor clip coefficients, one multi-tensor update pass.
// Rounding follows the reference's op sequence (which ops fuse into an FMA was pinned against the
// reference run on CPU, see oracle/optim_oracle.py).
#include "common.cuh"

namespace hmmc {

constexpr int OPT_THREADS = 256;
constexpr int OPT_BLOCK_ELEMS = 8192;   // == hmmc_ema_block_elems(): one block table serves both

// tensor owning block `blk`: last t with block_offsets[t] <= blk
__device__ __forceinline__ int owner_of_block(const int64_t* __restrict__ block_offsets, int n, int64_t blk) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (block_offsets[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---------------------------------------------------------------- pass 1: sum of squares per block
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_kernel(const uint64_t* __restrict__ g_ptrs, const int64_t* __restrict__ numels,
                   const int64_t* __restrict__ block_offsets, int n, const float* __restrict__ inv_scale,
                   float* __restrict__ partials) {
  __shared__ float red[32];
  // AMP: the norms are those of the UNSCALED gradients fl(g * 1/scale), the values GradScaler.unscale_ would
  // have left in place (main_pretrain.py:282-284 -> torch.amp.GradScaler.step)
  const float inv = inv_scale != nullptr ? __ldg(inv_scale) : 1.0f;
  const int64_t blk = blockIdx.x;
  const int t = owner_of_block(block_offsets, n, blk);
  const int64_t begin = (blk - block_offsets[t]) * OPT_BLOCK_ELEMS;
  const int64_t end = min(begin + int64_t(OPT_BLOCK_ELEMS), numels[t]);
  const float* g = reinterpret_cast<const float*>(g_ptrs[t]);
  float ss = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int64_t nvec = (end - begin) / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g + begin);
    for (int64_t i = threadIdx.x; i < nvec; i += OPT_THREADS) {
      float4 v = __ldg(g4 + i);
      v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv); v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    for (int64_t i = begin + nvec * 4 + threadIdx.x; i < end; i += OPT_THREADS) {
      const float x = __fmul_rn(g[i], inv);
      ss = fmaf(x, x, ss);
    }
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += OPT_THREADS) {
      const float x = __fmul_rn(g[i], inv);
      ss = fmaf(x, x, ss);
    }
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) partials[blk] = ss;
}

// ---------------------------------------------------------------- pass 2: norms -> clip coefficients
// Block t sums tensor t's partials (fixed order, double).  The last block to finish turns the n
// sums into  total_norm, the global coefficient  min(1, G/(total+1e-6))  and per tensor
// min(1, max_grad_norm_t/(||g_t||*cg + 1e-6))  — torch's clip_grad_norm_ applied twice, once over
// all parameters and once per parameter inside BertAdam.step (optimization.py:135-136).
__global__ void __launch_bounds__(OPT_THREADS)
clip_coefs_kernel(const float* __restrict__ partials, const int64_t* __restrict__ block_offsets, int n,
                  const float* __restrict__ hyper, float global_max_norm, double* __restrict__ sq,
                  float* __restrict__ coefs, float* __restrict__ norms_out, unsigned int* __restrict__ ticket) {
  __shared__ double redd[OPT_THREADS / 32];
  __shared__ bool last;
  __shared__ float cg_s;
  const int t = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  {
    double acc = 0.0;
    for (int64_t i = block_offsets[t] + threadIdx.x; i < block_offsets[t + 1]; i += OPT_THREADS) acc += double(partials[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) redd[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < OPT_THREADS / 32; ++i) s += redd[i];
      sq[t] = s;
      __threadfence();
      last = (atomicAdd(ticket, 1u) == unsigned(n - 1));
    }
    __syncthreads();
  }
  if (!last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += OPT_THREADS) acc += __ldcg(sq + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __syncthreads();
  if (lane == 0) redd[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < OPT_THREADS / 32; ++i) s += redd[i];
    const float total = float(sqrt(s));
    float cg = 1.0f;
    if (global_max_norm > 0.f) cg = fminf(__fdiv_rn(global_max_norm, __fadd_rn(total, 1e-6f)), 1.0f);
    cg_s = cg;
    if (norms_out != nullptr) norms_out[n] = total;
  }
  __syncthreads();
  const float cg = cg_s;
  for (int i = threadIdx.x; i < n; i += OPT_THREADS) {
    const float nt = float(sqrt(__ldcg(sq + i)));
    const float mg = hyper[i * 8 + 7];
    float ct = 1.0f;
    if (mg > 0.f) ct = fminf(__fdiv_rn(mg, __fadd_rn(__fmul_rn(nt, cg), 1e-6f)), 1.0f);
    coefs[2 * i] = cg;
    coefs[2 * i + 1] = ct;
    if (norms_out != nullptr) norms_out[i] = nt;
  }
}

// ---------------------------------------------------------------- pass 3: the update
struct AdamHyper { float lr, wd, b1, omb1, b2, omb2, eps, cg, ct, inv; };

// One element of BertAdam.step, each line one rounded op of the reference (optimization.py:141-166);
// add_(grad, alpha) and addcmul_ are single fused multiply-adds there, everything else rounds separately.
__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamHyper& h) {
  g = __fmul_rn(g, h.inv);                                       // GradScaler's unscale (1 without AMP: exact)
  g = __fmul_rn(__fmul_rn(g, h.cg), h.ct);                       // clip_grad_norm_ twice (global, per tensor)
  m = __fmaf_rn(h.omb1, g, __fmul_rn(m, h.b1));                  // next_m.mul_(b1).add_(grad, alpha=1-b1)
  v = __fmaf_rn(__fmul_rn(h.omb2, g), g, __fmul_rn(v, h.b2));    // next_v.mul_(b2).addcmul_(grad, grad, value=1-b2)
  float u = __fdiv_rn(m, __fadd_rn(__fsqrt_rn(v), h.eps));       // next_m / (next_v.sqrt() + e)
  if (h.wd > 0.f) u = __fadd_rn(u, __fmul_rn(h.wd, p));          // update += weight_decay * p
  p = __fadd_rn(p, -__fmul_rn(h.lr, u));                         // p.add_(-(lr_scheduled * update))
}

template <bool WRITE_G>
__global__ void __launch_bounds__(OPT_THREADS)
bert_adam_kernel(const uint64_t* __restrict__ p_ptrs, const uint64_t* __restrict__ g_ptrs,
                 const uint64_t* __restrict__ m_ptrs, const uint64_t* __restrict__ v_ptrs,
                 const int64_t* __restrict__ numels, const int64_t* __restrict__ block_offsets, int n,
                 const float* __restrict__ hyper, const float* __restrict__ coefs,
                 const float* __restrict__ inv_scale) {
  const int64_t blk = blockIdx.x;
  const int t = owner_of_block(block_offsets, n, blk);
  const int64_t begin = (blk - block_offsets[t]) * OPT_BLOCK_ELEMS;
  const int64_t end = min(begin + int64_t(OPT_BLOCK_ELEMS), numels[t]);
  AdamHyper h;
  h.inv = inv_scale != nullptr ? __ldg(inv_scale) : 1.0f;
  h.lr = hyper[t * 8 + 0]; h.wd = hyper[t * 8 + 1]; h.b1 = hyper[t * 8 + 2]; h.omb1 = hyper[t * 8 + 3];
  h.b2 = hyper[t * 8 + 4]; h.omb2 = hyper[t * 8 + 5]; h.eps = hyper[t * 8 + 6];
  h.cg = coefs[2 * t]; h.ct = coefs[2 * t + 1];
  float* p = reinterpret_cast<float*>(p_ptrs[t]);
  float* g = reinterpret_cast<float*>(g_ptrs[t]);
  float* m = reinterpret_cast<float*>(m_ptrs[t]);
  float* v = reinterpret_cast<float*>(v_ptrs[t]);
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int64_t tail = begin;
  if (aligned) {
    const int64_t nvec = (end - begin) / 4;
    float4* p4 = reinterpret_cast<float4*>(p + begin);
    float4* g4 = reinterpret_cast<float4*>(g + begin);
    float4* m4 = reinterpret_cast<float4*>(m + begin);
    float4* v4 = reinterpret_cast<float4*>(v + begin);
    for (int64_t i = threadIdx.x; i < nvec; i += OPT_THREADS) {
      float4 pp = p4[i], gg = __ldcs(g4 + i), mm = m4[i], vv = v4[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, h);
      adam_one(pp.y, gg.y, mm.y, vv.y, h);
      adam_one(pp.z, gg.z, mm.z, vv.z, h);
      adam_one(pp.w, gg.w, mm.w, vv.w, h);
      p4[i] = pp; m4[i] = mm; v4[i] = vv;
      if (WRITE_G) g4[i] = gg;
    }
    tail = begin + nvec * 4;
  }
  for (int64_t i = tail + threadIdx.x; i < end; i += OPT_THREADS) {
    float pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    adam_one(pp, gg, mm, vv, h);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (WRITE_G) g[i] = gg;
  }
}

struct AdamWs {
  float* partials;
  double* sq;
  float* coefs;
  unsigned int* ticket;
};
static void adam_carve(Workspace& ws, AdamWs& w, int n, int64_t total_blocks) {
  w.ticket = ws.take<unsigned int>(64);
  w.partials = ws.take<float>(size_t(total_blocks));
  w.sq = ws.take<double>(size_t(n));
  w.coefs = ws.take<float>(size_t(n) * 2);
}


// passes 1 + 2
static int clip_coefs(const uint64_t* g_ptrs, const int64_t* numels, const int64_t* block_offsets, int n,
                      int64_t total_blocks, const float* hyper, float global_max_norm, const float* inv_scale,
                      float* norms_out, const AdamWs& w, cudaStream_t st) {
  HMMC_CHECK_CUDA(cudaMemsetAsync(w.ticket, 0, sizeof(unsigned int), st));
  grad_sqnorm_kernel<<<unsigned(total_blocks), OPT_THREADS, 0, st>>>(g_ptrs, numels, block_offsets, n, inv_scale,
                                                                     w.partials);
  HMMC_CHECK_LAUNCH();
  clip_coefs_kernel<<<unsigned(n), OPT_THREADS, 0, st>>>(w.partials, block_offsets, n, hyper, global_max_norm, w.sq,
                                                        w.coefs, norms_out, w.ticket);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

// g *= coefs[2t] (the global coefficient) for clip_grad_norm_ used on its own
__global__ void __launch_bounds__(OPT_THREADS)
grad_scale_kernel(const uint64_t* __restrict__ g_ptrs, const int64_t* __restrict__ numels,
                  const int64_t* __restrict__ block_offsets, int n, const float* __restrict__ coefs) {
  const int64_t blk = blockIdx.x;
  const int t = owner_of_block(block_offsets, n, blk);
  const int64_t begin = (blk - block_offsets[t]) * OPT_BLOCK_ELEMS;
  const int64_t end = min(begin + int64_t(OPT_BLOCK_ELEMS), numels[t]);
  const float c = coefs[2 * t];
  float* g = reinterpret_cast<float*>(g_ptrs[t]);
  for (int64_t i = begin + threadIdx.x; i < end; i += OPT_THREADS) g[i] = __fmul_rn(g[i], c);
}

}  // namespace hmmc

using namespace hmmc;

extern "C" {

size_t hmmc_bert_adam_workspace_bytes(int n, int64_t total_blocks) {
  Workspace ws(nullptr, 0);
  AdamWs w;
  adam_carve(ws, w, n > 0 ? n : 0, total_blocks > 0 ? total_blocks : 0);
  ws.take<float>(size_t(n > 0 ? n : 0) * 8);   // hyper table of hmmc_clip_grad_norm_multi
  return ws.used + 256;
}

int hmmc_bert_adam_multi(const uint64_t* p_ptrs, const uint64_t* g_ptrs, const uint64_t* m_ptrs,
                         const uint64_t* v_ptrs, const int64_t* numels, const int32_t* dtypes,
                         const int64_t* block_offsets, int n, int64_t total_blocks, const float* hyper,
                         float global_max_norm, int write_back_grads, const float* inv_scale, float* norms_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  (void)dtypes;   // fp32 only; the host checks, the table keeps the EMA layout
  if (n <= 0 || total_blocks <= 0) return HMMC_OK;
  HMMC_REQUIRE(p_ptrs && g_ptrs && m_ptrs && v_ptrs && numels && block_offsets && hyper, "bert_adam: null table");
  HMMC_REQUIRE(total_blocks < (int64_t(1) << 31), "bert_adam: too many blocks");
  Workspace ws(workspace, workspace_bytes);
  AdamWs w;
  adam_carve(ws, w, n, total_blocks);
  HMMC_REQUIRE(ws.ok(), "bert_adam: workspace too small (%zu needed, %zu given)", ws.used, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = clip_coefs(g_ptrs, numels, block_offsets, n, total_blocks, hyper, global_max_norm, inv_scale, norms_out, w, st);
  if (rc) return rc;
  if (write_back_grads)
    bert_adam_kernel<true><<<unsigned(total_blocks), OPT_THREADS, 0, st>>>(p_ptrs, g_ptrs, m_ptrs, v_ptrs, numels,
                                                                            block_offsets, n, hyper, w.coefs, inv_scale);
  else
    bert_adam_kernel<false><<<unsigned(total_blocks), OPT_THREADS, 0, st>>>(p_ptrs, g_ptrs, m_ptrs, v_ptrs, numels,
                                                                             block_offsets, n, hyper, w.coefs, inv_scale);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_clip_grad_norm_multi(const uint64_t* g_ptrs, const int64_t* numels, const int64_t* block_offsets, int n,
                              int64_t total_blocks, float max_norm, float* norms_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (n <= 0 || total_blocks <= 0) return HMMC_OK;
  HMMC_REQUIRE(g_ptrs && numels && block_offsets, "clip_grad_norm: null table");
  HMMC_REQUIRE(total_blocks < (int64_t(1) << 31), "clip_grad_norm: too many blocks");
  HMMC_REQUIRE(max_norm > 0.f, "clip_grad_norm: max_norm must be positive");
  Workspace ws(workspace, workspace_bytes);
  AdamWs w;
  adam_carve(ws, w, n, total_blocks);
  float* hyper = ws.take<float>(size_t(n) * 8);    // all zero: no per-tensor clip
  HMMC_REQUIRE(ws.ok(), "clip_grad_norm: workspace too small (%zu needed, %zu given)", ws.used, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_CHECK_CUDA(cudaMemsetAsync(hyper, 0, size_t(n) * 8 * sizeof(float), st));
  int rc = clip_coefs(g_ptrs, numels, block_offsets, n, total_blocks, hyper, max_norm, nullptr, norms_out, w, st);
  if (rc) return rc;
  grad_scale_kernel<<<unsigned(total_blocks), OPT_THREADS, 0, st>>>(g_ptrs, numels, block_offsets, n, w.coefs);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

}  // extern "C"

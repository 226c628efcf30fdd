// Shared helpers of libhmmc_head.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/hmmc_head.h"

namespace hmmc {

void set_error(const char* fmt, ...);
void count_launch();

#define HMMC_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      hmmc::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return HMMC_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

// every kernel launch of this library goes through here (also feeds hmmc_launch_count())
#define HMMC_CHECK_LAUNCH()                  \
  do {                                       \
    hmmc::count_launch();                    \
    HMMC_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

#define HMMC_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      hmmc::set_error(__VA_ARGS__);      \
      return HMMC_ERR_ARG;               \
    }                                    \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Launch with programmatic stream serialisation: the kernel may start while the previous kernel of the stream
// drains, and must execute griddepcontrol.wait (ptx::grid_dependency_wait) before it touches global memory.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int sm_count();

// Bump allocator over the caller's workspace.
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t off = align_up(used, 256);
    size_t bytes = count * sizeof(T);
    if (base == nullptr || off + bytes > size) {
      used = off + bytes;  // keep counting so the caller can report the need
      return nullptr;
    }
    used = off + bytes;
    return reinterpret_cast<T*>(base + off);
  }
  bool ok() const { return base != nullptr && used <= size; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) t = warp_max(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// bf16 hi/lo split of an fp32 value: hi = bf16(x), lo = bf16(x - hi).
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

static inline int planes_of(int prec) { return prec == HMMC_PREC_BF16X3 ? 2 : 1; }

}  // namespace hmmc

// ---- internal entry points shared between translation units (defined in core.cu)
namespace hmmc {
int gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
             int64_t ldc, int M, int N, int K, float alpha, cudaStream_t st);
int rownorm_pack(const float* x, int64_t R, int D, int64_t ldx, float eps, int planes, float* xhat, float* inv_norm,
                 void* packed, int64_t ldp, cudaStream_t st);
// tcgen05 GEMM, C[split] = alpha * A.B^T (fp32 out); splits > 1 writes partial sums split_stride apart.
// tiling: 0 = chosen from the shape; 128 / 256 = single-CTA kernel with that tile width; 512 = CTA-pair
// kernel (256 x 256 tiles) -- explicit values are for hmmc_umma_gemm_nt_tiled (tools/gemm_bench.py)
int umma_gemm_store(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                    int64_t split_stride, int M, int N, int K, int planes, int splits, float alpha,
                    cudaStream_t st, int tiling = 0);
// several store-GEMMs (each with its own split count) in one grouped launch
struct StoreGemm {
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  float* C; int64_t ldc; int64_t split_stride;
  int M, N, K, splits;
};
int umma_gemm_store_grouped(const StoreGemm* g, int n, int planes, float alpha, cudaStream_t st);
// number of split-K partials umma_gemm_store(splits) really writes
int umma_effective_splits(int K, int planes, int splits);
// src [rows, cols] fp32 -> straight [rows, planes*cols] and transposed [cols, planes*rows] bf16 plane
// packs (either output may be null); defined in pretrain.cu
int pack_dual(const float* src, int rows, int cols, int planes, void* straight, void* transposed, cudaStream_t st);
// fused-tile eval (eval_fused.cu): gallery pack (1 + 12 rows per video, 208-row tiles) and the materialising
// sweep that writes sim / fsim (or their sum) from the register top-k, no [Nt, Nv*F] intermediate
size_t eval_gallery_pack_rows(int64_t Nv);
int eval_pack_gallery(const float* video, const float* frames, int64_t Nv, int F, int D, int planes, void* out,
                      cudaStream_t st, const float* text = nullptr, int64_t Nt = 0, void* text_out = nullptr);
int eval_sim_write(const void* text_packed, const void* gallery_packed, int64_t Nt, int64_t Nv, int D, int prec,
                   float scale, int top_k, float* sim, float* fsim, int64_t ld_out, int combine, cudaStream_t st);
}  // namespace hmmc

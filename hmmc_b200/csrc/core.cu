// Error state, device checks, tensor-map construction and the two GEMM primitives.
#include "common.cuh"
#include "umma_gemm.cuh"
#include <cudaTypedefs.h>
#include <atomic>
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace hmmc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return HMMC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};   // bytes, dimension 1
  if (box_cols != 64 && box_cols != 32) {
    set_error("make_tmap_bf16: box of %u columns (64 = operand loads, 32 = epilogue store tiles)", box_cols);
    return HMMC_ERR_ARG;
  }
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  // the swizzle span equals the box's inner extent: 128 B for the operand tiles, 64 B for the store tiles
  const CUtensorMapSwizzle sw = (box_cols == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box_rows=%u)", int(r),
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return HMMC_ERR_CUDA;
  }
  return HMMC_OK;
}

// ---------------------------------------------------------------- rownorm + pack
// One warp per row.  Sum of squares in fp32 (lane-strided, then shuffle tree).
__global__ void rownorm_pack_kernel(const float* __restrict__ x, int64_t R, int D, int64_t ldx, float eps, int planes,
                                    float* __restrict__ xhat, float* __restrict__ inv_norm,
                                    __nv_bfloat16* __restrict__ packed, int64_t ldp) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* xr = x + row * ldx;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  float n = sqrtf(ss);
  if (eps > 0.f) n = fmaxf(n, eps);
  const float inv = 1.0f / n;    // eps == 0 and a zero row: inf, x*inf = NaN, like the reference's 0/0
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d] / n;
    if (xhat != nullptr) xhat[row * int64_t(D) + d] = v;
    if (packed != nullptr) {
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      packed[row * ldp + d] = hi;
      if (planes == 2) packed[row * ldp + D + d] = lo;
    }
  }
}

// ---------------------------------------------------------------- SIMT fp32 GEMM
// C[m,n] = alpha * sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk]; 64x64 tile, 16-deep k-slab,
// 256 threads each owning a 4x4 micro-tile.  Reference-grade path (HMMC_PREC_FP32).
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbn,
                int64_t sbk, float* __restrict__ C, int64_t ldc, int M, int N, int K, float alpha) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    // 64 x 16 elements per operand, 256 threads -> 4 each
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      int r, k;
      if (sak == 1) { k = idx & 15; r = idx >> 4; } else { r = idx & 63; k = idx >> 6; }
      const int gm = m0 + r, gk = k0 + k;
      As[k][r] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      int r, k;
      if (sbk == 1) { k = idx & 15; r = idx >> 4; } else { r = idx & 63; k = idx >> 6; }
      const int gn = n0 + r, gk = k0 + k;
      Bs[k][r] = (gn < N && gk < K) ? B[gn * sbn + gk * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) C[gm * ldc + gn] = alpha * acc[i][j];
    }
  }
}

int gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
             int64_t ldc, int M, int N, int K, float alpha, cudaStream_t st) {
  if (M <= 0 || N <= 0) return HMMC_OK;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, alpha);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int rownorm_pack(const float* x, int64_t R, int D, int64_t ldx, float eps, int planes, float* xhat, float* inv_norm,
                 void* packed, int64_t ldp, cudaStream_t st) {
  if (R <= 0) return HMMC_OK;
  const int warps = 8;
  rownorm_pack_kernel<<<unsigned((R + warps - 1) / warps), warps * 32, 0, st>>>(
      x, R, D, ldx, eps, planes, xhat, inv_norm, static_cast<__nv_bfloat16*>(packed), ldp);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int umma_gemm_store(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                    int64_t split_stride, int M, int N, int K, int planes, int splits, float alpha,
                    cudaStream_t st, int tiling) {
  EpiStoreF32::Params ep{C, ldc, split_stride, alpha};
  if (tiling == 128) return launch_umma_gemm<128, EpiStoreF32>(A, lda, B, ldb, M, N, K, planes, splits, ep, st);
  if (tiling == 512) {     // CTA-pair kernel, 256 x 256 tiles
    GemmProblem<EpiStoreF32> pr{A, lda, B, ldb, M, N, K, planes, splits, ep};
    return launch_umma_grouped_pair<EpiStoreF32>(&pr, 1, st);
  }
  if (tiling == 256) return launch_umma_gemm<256, EpiStoreF32>(A, lda, B, ldb, M, N, K, planes, splits, ep, st);
  if (N % 256 == 0 || N > 1024) return launch_umma_gemm<256, EpiStoreF32>(A, lda, B, ldb, M, N, K, planes, splits, ep, st);
  return launch_umma_gemm<128, EpiStoreF32>(A, lda, B, ldb, M, N, K, planes, splits, ep, st);
}

int umma_gemm_store_grouped(const StoreGemm* g, int n, int planes, float alpha, cudaStream_t st) {
  if (n <= 0) return HMMC_OK;
  bool wide = n <= UMMA_MAX_PROBLEMS;
  for (int i = 0; i < n; ++i) wide = wide && (g[i].N % 256 == 0);
  if (!wide) {           // shapes the 256-wide tile does not divide: one launch each
    for (int i = 0; i < n; ++i) {
      const int rc = umma_gemm_store(g[i].A, g[i].lda, g[i].B, g[i].ldb, g[i].C, g[i].ldc, g[i].split_stride, g[i].M,
                                     g[i].N, g[i].K, planes, g[i].splits, alpha, st);
      if (rc) return rc;
    }
    return HMMC_OK;
  }
  GemmProblem<EpiStoreF32> pr[UMMA_MAX_PROBLEMS];
  for (int i = 0; i < n; ++i)
    pr[i] = GemmProblem<EpiStoreF32>{g[i].A, g[i].lda, g[i].B, g[i].ldb, g[i].M, g[i].N, g[i].K, planes, g[i].splits,
                                     EpiStoreF32::Params{g[i].C, g[i].ldc, g[i].split_stride, alpha}};
  return launch_umma_grouped<256, EpiStoreF32>(pr, n, st);
}

int umma_effective_splits(int K, int planes, int splits) { return effective_splits(K, planes, splits); }

}  // namespace hmmc

using namespace hmmc;

extern "C" {

const char* hmmc_last_error(void) { return g_err; }
int hmmc_version(void) { return 100; }
unsigned long long hmmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int hmmc_device_check(void) {
  int dev = 0;
  HMMC_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  HMMC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  HMMC_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("libhmmc_head is built for sm_100a (B200) only; device %d is sm_%d%d", dev, major, minor);
    return HMMC_ERR_UNSUPPORTED;
  }
  return HMMC_OK;
}

int hmmc_rownorm_pack(const float* x, int64_t R, int D, int64_t ldx, float eps, int planes, float* xhat,
                      float* inv_norm, void* packed, int64_t ld_packed, void* stream) {
  HMMC_REQUIRE(x != nullptr && D > 0 && ldx >= D, "rownorm_pack: bad arguments");
  HMMC_REQUIRE(planes == 1 || planes == 2, "rownorm_pack: planes must be 1 or 2");
  HMMC_REQUIRE(packed == nullptr || ld_packed >= int64_t(planes) * D, "rownorm_pack: ld_packed too small");
  return rownorm_pack(x, R, D, ldx, eps, planes, xhat, inv_norm, packed, ld_packed, static_cast<cudaStream_t>(stream));
}

int hmmc_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
                  int64_t ldc, int M, int N, int K, float alpha, void* stream) {
  HMMC_REQUIRE(A && B && C && K > 0, "gemm_f32: bad arguments");
  return gemm_f32(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, alpha, static_cast<cudaStream_t>(stream));
}

int hmmc_umma_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                      int K, int planes, float alpha, void* stream) {
  HMMC_REQUIRE(A && B && C, "umma_gemm_nt: null operand");
  return umma_gemm_store(A, lda, B, ldb, C, ldc, 0, M, N, K, planes, 1, alpha, static_cast<cudaStream_t>(stream));
}

int hmmc_umma_gemm_nt_tiled(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M,
                            int N, int K, int planes, float alpha, int tiling, void* stream) {
  HMMC_REQUIRE(A && B && C, "umma_gemm_nt_tiled: null operand");
  HMMC_REQUIRE(tiling == 0 || tiling == 128 || tiling == 256 || tiling == 512,
               "umma_gemm_nt_tiled: tiling must be 0, 128, 256 or 512 (got %d)", tiling);
  return umma_gemm_store(A, lda, B, ldb, C, ldc, 0, M, N, K, planes, 1, alpha, static_cast<cudaStream_t>(stream),
                         tiling);
}

}  // extern "C"

// Eval: similarity + top-k frame pooling, and rank counting for the retrieval metrics.
#include "common.cuh"
#include "umma_gemm.cuh"

namespace hmmc {

// mean of the k largest of F values (torch.topk(..., k)[0].mean): one thread per (text, video).
// F <= 32.  Selection keeps a small sorted register array (descending), the sum runs in
// descending order like torch's reduction over the top-k output.
template <int MAXK>
__device__ __forceinline__ float topk_mean(const float* v, int F, int k) {
  float best[MAXK];
#pragma unroll
  for (int i = 0; i < MAXK; ++i) best[i] = -INFINITY;
  for (int f = 0; f < F; ++f) {
    float x = v[f];
#pragma unroll
    for (int i = 0; i < MAXK; ++i) {
      if (i < k && x > best[i]) { const float t = best[i]; best[i] = x; x = t; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXK; ++i) if (i < k) s += best[i];
  return s / float(k);
}

constexpr int TOPK_MAX = 16;

__global__ void topk_frames_kernel(const float* __restrict__ SF, int64_t Nt, int64_t Nv, int F, int k,
                                   float* __restrict__ fsim, int64_t ld_out) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= Nt * Nv) return;
  const int64_t t = idx / Nv, v = idx - t * Nv;
  const float* p = SF + (t * Nv + v) * F;
  float buf[32];
  for (int f = 0; f < F; ++f) buf[f] = p[f];
  fsim[t * ld_out + v] = topk_mean<TOPK_MAX>(buf, F, k);
}

// ------------------------------------------------------------------ rank counting
// t2v: one warp per text row.
__global__ void rank_t2v_kernel(const float* __restrict__ sim, int64_t lds, int Nt, int Nv,
                                const int32_t* __restrict__ gt, int32_t* __restrict__ t2v) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t s = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= Nt) return;
  const float* row = sim + s * lds;
  const float ref = row[gt[s]];
  int cnt = 0;
  for (int j = lane; j < Nv; j += 32) cnt += (row[j] > ref) ? 1 : 0;
  cnt = warp_sum_int(cnt);
  if (lane == 0) t2v[s] = cnt;
}

// theta[j] = max over the captions of video j of sim[s,j]   (NaN -> -inf, metrics.py:83)
// also clears the v2t counters the next kernel accumulates into
__global__ void rank_theta_kernel(const float* __restrict__ sim, int64_t lds, int Nv,
                                  const int32_t* __restrict__ group_start, float* __restrict__ theta,
                                  int32_t* __restrict__ v2t) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Nv) return;
  v2t[j] = 0;
  float m = -INFINITY;
  for (int s = group_start[j]; s < group_start[j + 1]; ++s) {
    float v = sim[int64_t(s) * lds + j];
    if (v != v) v = -INFINITY;
    m = fmaxf(m, v);
  }
  theta[j] = m;
}

// t2v counts and the v2t thresholds in one launch (they are independent): blocks [0, t2v_blocks) count, the rest
// compute theta
__global__ void rank_t2v_theta_kernel(const float* __restrict__ sim, int64_t lds, int Nt, int Nv,
                                      const int32_t* __restrict__ gt, int32_t* __restrict__ t2v, int t2v_blocks,
                                      const int32_t* __restrict__ group_start, float* __restrict__ theta,
                                      int32_t* __restrict__ v2t) {
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  if (int(blockIdx.x) < t2v_blocks) {
    const int lane = threadIdx.x & 31;
    const int64_t s = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= Nt) return;
    const float* row = sim + s * lds;
    const float ref = row[gt[s]];
    int cnt = 0;
    for (int j = lane; j < Nv; j += 32) cnt += (row[j] > ref) ? 1 : 0;
    cnt = warp_sum_int(cnt);
    if (lane == 0) t2v[s] = cnt;
    return;
  }
  const int j = (int(blockIdx.x) - t2v_blocks) * blockDim.x + threadIdx.x;
  if (j >= Nv) return;
  v2t[j] = 0;
  float m = -INFINITY;
  for (int s = group_start[j]; s < group_start[j + 1]; ++s) {
    float v = sim[int64_t(s) * lds + j];
    if (v != v) v = -INFINITY;
    m = fmaxf(m, v);
  }
  theta[j] = m;
}

// v2t: block = 32 video columns x a slice of caption groups; warps stride the groups.
__global__ void __launch_bounds__(256)
rank_v2t_kernel(const float* __restrict__ sim, int64_t lds, int Nv, const int32_t* __restrict__ group_start,
                const float* __restrict__ theta, int groups_per_block, int32_t* __restrict__ v2t) {
  __shared__ int cnts[8][32];
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int g0 = blockIdx.y * groups_per_block;
  const int g1 = min(g0 + groups_per_block, Nv);
  int cnt = 0;
  if (j < Nv) {
    const float th = theta[j];
    for (int g = g0 + warp; g < g1; g += 8) {
      float m = -INFINITY;
      for (int s = group_start[g]; s < group_start[g + 1]; ++s) {
        float v = sim[int64_t(s) * lds + j];
        if (v != v) v = -INFINITY;
        m = fmaxf(m, v);
      }
      cnt += (m > th) ? 1 : 0;
    }
  }
  cnts[warp][lane] = cnt;
  __syncthreads();
  if (warp == 0 && j < Nv) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += cnts[w][lane];
    if (t) atomicAdd(&v2t[j], t);
  }
}

}  // namespace hmmc

using namespace hmmc;

extern "C" {

size_t hmmc_sim_topk_workspace_bytes(int64_t Nt, int64_t Nv, int F, int D, int prec) {
  Workspace ws(nullptr, 0);
  const int planes = planes_of(prec);
  if (prec == HMMC_PREC_FP32) {
    ws.take<float>(size_t(Nt) * D);
    ws.take<float>(size_t(Nv) * D);
    ws.take<float>(size_t(Nv) * F * D);
  } else if (hmmc_eval_fused_supported(F, D, 1)) {
    // fused tiles: packed text + packed gallery only (top_k is checked at call time; sized for either path)
    ws.take<__nv_bfloat16>(size_t(Nt) * planes * D);
    ws.take<__nv_bfloat16>(eval_gallery_pack_rows(Nv) * planes * D);
    Workspace alt(nullptr, 0);
    alt.take<__nv_bfloat16>(size_t(Nt) * planes * D);
    alt.take<__nv_bfloat16>(size_t(Nv) * planes * D);
    alt.take<__nv_bfloat16>(size_t(Nv) * F * planes * D);
    alt.take<float>(size_t(Nt) * Nv * F);
    (void)alt;
    return ws.used + 256;
  } else {
    ws.take<__nv_bfloat16>(size_t(Nt) * planes * D);
    ws.take<__nv_bfloat16>(size_t(Nv) * planes * D);
    ws.take<__nv_bfloat16>(size_t(Nv) * F * planes * D);
  }
  ws.take<float>(size_t(Nt) * Nv * F);
  return ws.used + 256;
}

int hmmc_sim_topk_fwd(const float* text, int64_t Nt, const float* video, const float* frames, int64_t Nv, int F, int D,
                      float scale, int top_k, int prec, float* sim, float* fsim, int64_t ld_out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_REQUIRE(text && Nt > 0 && Nv > 0 && D > 0, "sim_topk: bad arguments");
  HMMC_REQUIRE(sim == nullptr || video != nullptr, "sim_topk: sim requested without video embeddings");
  HMMC_REQUIRE(fsim == nullptr || (frames != nullptr && F > 0 && F <= 32), "sim_topk: fsim needs frames with 1 <= F <= 32");
  // torch.topk raises when k is out of range (main_task_retrieval.py:335)
  HMMC_REQUIRE(fsim == nullptr || (top_k >= 1 && top_k <= F && top_k <= TOPK_MAX), "sim_topk: top_k=%d out of range for F=%d", top_k, F);
  HMMC_REQUIRE(ld_out >= Nv, "sim_topk: ld_out too small");
  HMMC_REQUIRE(prec >= 0 && prec <= 2, "sim_topk: unknown precision %d", prec);
  HMMC_REQUIRE(Nt * Nv * int64_t(F > 0 ? F : 1) < (int64_t(1) << 40), "sim_topk: tile too large, split the gallery");
  Workspace ws(workspace, workspace_bytes);
  const int planes = planes_of(prec);
  int rc;
  float* SF = nullptr;
  if (prec == HMMC_PREC_FP32) {
    float* th = ws.take<float>(size_t(Nt) * D);
    float* vh = ws.take<float>(size_t(Nv) * D);
    float* fh = ws.take<float>(size_t(Nv) * F * D);
    SF = ws.take<float>(size_t(Nt) * Nv * F);
    if (!ws.ok()) { set_error("sim_topk: workspace too small (%zu > %zu)", ws.used, workspace_bytes); return HMMC_ERR_WORKSPACE; }
    if ((rc = rownorm_pack(text, Nt, D, D, 0.f, 1, th, nullptr, nullptr, 0, st))) return rc;
    if (sim != nullptr) {
      if ((rc = rownorm_pack(video, Nv, D, D, 0.f, 1, vh, nullptr, nullptr, 0, st))) return rc;
      if ((rc = gemm_f32(th, D, 1, vh, D, 1, sim, ld_out, int(Nt), int(Nv), D, scale, st))) return rc;
    }
    if (fsim != nullptr) {
      if ((rc = rownorm_pack(frames, Nv * F, D, D, 0.f, 1, fh, nullptr, nullptr, 0, st))) return rc;
      if ((rc = gemm_f32(th, D, 1, fh, D, 1, SF, Nv * F, int(Nt), int(Nv * F), D, scale, st))) return rc;
    }
  } else if (frames != nullptr && video != nullptr && hmmc_eval_fused_supported(F, D, top_k) && (sim != nullptr || fsim != nullptr)) {
    // fused tiles (F = 12, top_k <= 4): one GEMM sweep over [video | 12 frames] column groups, top-k pooling in
    // registers, scores written once - no [Nt, Nv*F] frame-similarity matrix, three launches in all
    __nv_bfloat16* tp = ws.take<__nv_bfloat16>(size_t(Nt) * planes * D);
    __nv_bfloat16* gp = ws.take<__nv_bfloat16>(eval_gallery_pack_rows(Nv) * planes * D);
    if (!ws.ok()) { set_error("sim_topk: workspace too small (%zu > %zu)", ws.used, workspace_bytes); return HMMC_ERR_WORKSPACE; }
    const int64_t ldp = int64_t(planes) * D;
    (void)ldp;
    if ((rc = eval_pack_gallery(video, frames, Nv, F, D, planes, gp, st, text, Nt, tp))) return rc;   // both operands
    // sim == fsim (one buffer): the caller wants the sum (main_task_retrieval.py:512-513)
    const int combine = (sim != nullptr && sim == fsim) ? 1 : 0;
    return eval_sim_write(tp, gp, Nt, Nv, D, prec, scale, top_k, sim, combine ? nullptr : fsim, ld_out, combine, st);
  } else {
    HMMC_REQUIRE(D % 64 == 0, "sim_topk: tensor-core path needs D %% 64 == 0 (D=%d)", D);
    __nv_bfloat16* tp = ws.take<__nv_bfloat16>(size_t(Nt) * planes * D);
    __nv_bfloat16* vp = ws.take<__nv_bfloat16>(size_t(Nv) * planes * D);
    __nv_bfloat16* fp = ws.take<__nv_bfloat16>(size_t(Nv) * F * planes * D);
    SF = ws.take<float>(size_t(Nt) * Nv * F);
    if (!ws.ok()) { set_error("sim_topk: workspace too small (%zu > %zu)", ws.used, workspace_bytes); return HMMC_ERR_WORKSPACE; }
    const int64_t ldp = int64_t(planes) * D;
    if ((rc = rownorm_pack(text, Nt, D, D, 0.f, planes, nullptr, nullptr, tp, ldp, st))) return rc;
    if (sim != nullptr) {
      if ((rc = rownorm_pack(video, Nv, D, D, 0.f, planes, nullptr, nullptr, vp, ldp, st))) return rc;
      if ((rc = umma_gemm_store(tp, ldp, vp, ldp, sim, ld_out, 0, int(Nt), int(Nv), D, planes, 1, scale, st))) return rc;
    }
    if (fsim != nullptr) {
      if ((rc = rownorm_pack(frames, Nv * F, D, D, 0.f, planes, nullptr, nullptr, fp, ldp, st))) return rc;
      if ((rc = umma_gemm_store(tp, ldp, fp, ldp, SF, Nv * F, 0, int(Nt), int(Nv * F), D, planes, 1, scale, st))) return rc;
    }
  }
  if (fsim != nullptr) {
    const int64_t n = Nt * Nv;
    topk_frames_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(SF, Nt, Nv, F, top_k, fsim, ld_out);
    HMMC_CHECK_LAUNCH();
  }
  return HMMC_OK;
}

int hmmc_rank_count(const float* sim, int64_t lds, int Nt, int Nv, const int32_t* gt, const int32_t* group_start,
                    int32_t* t2v, int32_t* v2t, float* theta_scratch, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_REQUIRE(sim && Nt > 0 && Nv > 0 && lds >= Nv, "rank_count: bad arguments");
  HMMC_REQUIRE(t2v == nullptr || gt != nullptr, "rank_count: t2v needs gt");
  HMMC_REQUIRE(v2t == nullptr || (group_start != nullptr && theta_scratch != nullptr),
               "rank_count: v2t needs group_start and theta_scratch");
  if (t2v != nullptr && v2t != nullptr) {
    const int tb = (Nt + 7) / 8;
    count_launch();
    HMMC_CHECK_CUDA(launch_pdl(rank_t2v_theta_kernel, dim3(unsigned(tb + (Nv + 255) / 256)), dim3(256), 0, st, sim, lds, Nt,
                               Nv, gt, t2v, tb, group_start, theta_scratch, v2t));
  } else if (t2v != nullptr) {
    count_launch();
    HMMC_CHECK_CUDA(launch_pdl(rank_t2v_kernel, dim3(unsigned((Nt + 7) / 8)), dim3(256), 0, st, sim, lds, Nt, Nv, gt, t2v));
  }
  if (v2t != nullptr) {
    if (t2v == nullptr) {
      count_launch();
      HMMC_CHECK_CUDA(launch_pdl(rank_theta_kernel, dim3(unsigned((Nv + 255) / 256)), dim3(256), 0, st, sim, lds, Nv,
                                 group_start, theta_scratch, v2t));
    }
    const int col_blocks = (Nv + 31) / 32;
    int gy = (4 * sm_count() + col_blocks - 1) / col_blocks;   // enough blocks to fill the machine
    if (gy < 1) gy = 1;
    int gpb = (Nv + gy - 1) / gy;
    if (gpb < 8) gpb = 8;
    gy = (Nv + gpb - 1) / gpb;
    dim3 grid(col_blocks, gy);
    count_launch();
    HMMC_CHECK_CUDA(launch_pdl(rank_v2t_kernel, grid, dim3(256), 0, st, sim, lds, Nv, group_start, theta_scratch, gpb, v2t));
  }
  return HMMC_OK;
}

}  // extern "C"

namespace hmmc {
// tensor_video_to_text_sim (metrics.py:79-86): out[j, g] = max_{s in group g} sim[s, j], NaN -> -inf
__global__ void group_max_kernel(const float* __restrict__ sim, int64_t lds, int Nv, int G,
                                 const int32_t* __restrict__ group_start, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (j >= Nv || g >= G) return;
  float m = -INFINITY;
  for (int s = group_start[g]; s < group_start[g + 1]; ++s) {
    float v = sim[int64_t(s) * lds + j];
    if (v != v) v = -INFINITY;
    m = fmaxf(m, v);
  }
  out[int64_t(j) * G + g] = m;
}
}  // namespace hmmc

extern "C" int hmmc_group_max(const float* sim, int64_t lds, int Nv, int G, const int32_t* group_start, float* out,
                              void* stream) {
  HMMC_REQUIRE(sim && group_start && out && Nv > 0 && G > 0 && G < 65536, "group_max: bad arguments");
  dim3 grid((Nv + 127) / 128, G);
  hmmc::group_max_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(sim, lds, Nv, G, group_start, out);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

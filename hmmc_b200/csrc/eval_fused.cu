// Large-gallery retrieval eval without materialising the similarity matrix (BASELINE config 5).
//
//   score[s, j] = scale * t_hat_s . v_hat_j  +  mean of the top_k of { scale * t_hat_s . f_hat_{j,f} }_f
//   t2v_cnt[s]  = #{ j != gt(s) : score[s, j] > score[s, gt(s)] }
//   v2t_cnt[j]  = #{ groups g != j : max_{s in g} score[s, j] > theta_j },  theta_j = max_{s in group j} score[s, j]
//
// (the multi-sentence ranks of metrics.py:49-86; a square set is the special case of one caption per
// video).  The gallery is packed as 1+F consecutive rows per video (video embedding, then its frames)
// so one 128 x 208 accumulator tile (F = 12: 16 videos x 13 columns) holds everything the top-k
// pooling of 16 videos needs in ONE thread's row: no shuffles, no logits in HBM.
//
// Texts are packed group-aligned: the captions of one video never straddle a 128-row tile, so the
// "any caption of group g beats theta_j" reduction stays inside one CTA tile.
//
// Same warp roles and TMA / tcgen05 pipeline as umma_gemm.cuh; the scheduling differs: a CTA owns
// M tiles (128 captions) and sweeps all gallery tiles for each, so all CTAs walk the gallery in
// step and every gallery tile is fetched from HBM once and then served from L2, and each thread
// keeps its caption's t2v count in a register for the whole sweep.
#include "common.cuh"
#include "umma_gemm.cuh"
#include <stdlib.h>

namespace hmmc {

constexpr int EV_F = 12;                    // frames per video handled by the fused path
constexpr int EV_COLS = 1 + EV_F;           // accumulator columns per video
constexpr int EV_VPT = 16;                  // videos per tile
constexpr int EV_BN = EV_COLS * EV_VPT;     // 208
constexpr int EV_MAXK = 4;                  // top_frames supported by the fused path

struct EvalArgs {
  int num_m_blk, num_n_blk;       // caption tiles (128 rows) and gallery tiles (16 videos)
  int Nv_local;                   // videos in this shard
  int num_seg, kb_per_seg;
  int a_k0[3], b_k0[3];
  float scale;
  int top_k;
  int mode;                       // 0: ground-truth scores on the listed tiles, 1: counting sweep
  int video_base;                 // global index of this shard's first video
  const int32_t* grp;             // [Nt_pad] global video id each packed caption belongs to, -1 = padding
  float* gt_score;                // [Nt_pad] score[s, gt(s)] (mode 0 writes, mode 1 reads)
  const float* theta;             // [Nv_local] (mode 1)
  int32_t* t2v_cnt;               // [Nt_pad] (mode 1, one atomic per row half per caption tile)
  int32_t* v2t_cnt;               // [Nv_local] (mode 1, atomics)
  const int2* diag_tiles;         // mode 0: (m_blk, n_blk) pairs
  int n_diag_tiles;
  int panel_n_blk;                // mode 1: gallery tiles per L2-resident panel
  // mode 2 (materialise): scores written out instead of counted.  sim = s * t.v, fsim = mean of the top-k frame
  // similarities (either may be NULL); combine != 0: sim receives the sum (main_task_retrieval.py:512-513)
  float* sim_out;
  float* fsim_out;
  int64_t ld_out;
  int Nt;                         // valid text rows (rows beyond are padding of the last tile)
  int combine;
};

// OR-reduction of a predicate over the 256 epilogue threads (named barrier 1); also a barrier.
__device__ __forceinline__ int epi_bar_or(int pred) {
  int r;
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %1, 0;\n\t"
      "bar.red.or.pred q, 1, 256, p;\n\t"
      "selp.b32 %0, 1, 0, q;\n\t"
      "}\n"
      : "=r"(r)
      : "r"(pred)
      : "memory");
  return r;
}

// running top-K of a video's frame similarities: branch-free insertion with min/max
template <int K>
struct VideoAcc {
  float vsim;
  float best[K];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < K; ++i) best[i] = -INFINITY;
  }
  __device__ __forceinline__ void push(float x) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const float lo = fminf(best[i], x);
      best[i] = fmaxf(best[i], x);
      x = lo;
    }
  }
  // scale > 0 commutes with max and mean: one multiply per video instead of one per column
  __device__ __forceinline__ float frame_score(float scale) const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) s += best[i] * scale;        // descending order, like torch.topk(...).mean
    return s / float(K);
  }
  __device__ __forceinline__ float score(float scale) const { return vsim * scale + frame_score(scale); }
};

constexpr int EV_EPI_WARPS = 8;
constexpr int EV_THREADS = (4 + EV_EPI_WARPS) * 32;
constexpr int EV_HALF_COLS = EV_BN / 2;      // 104 columns = 8 videos per epilogue warp
constexpr int EV_HALF_VID = EV_VPT / 2;

// One epilogue warp's share of a tile: walk its 104 accumulator columns (8 videos x 13) in order,
// pool each video (video sim + mean of the top-K frame sims), update the caption's t2v count and
// return the bit mask of videos whose score beats theta (v2t candidates).
// WRITE (mode 2): the eight pooled scores of this thread's caption are kept in registers and written as two
// 16-byte stores per output matrix (one full 32-byte sector per row and tile half).
template <int K, bool WRITE>
__device__ __forceinline__ uint32_t eval_tile_columns(const EvalArgs& a, uint32_t taddr, int row, int n_blk, int half,
                                                      int my_grp, float my_gt, int& my_cnt) {
  const int v0 = n_blk * EV_VPT + half * EV_HALF_VID;     // first local video of this warp's columns
  uint32_t mask = 0;
  VideoAcc<K> va;
  va.reset();
  va.vsim = 0.f;
  float o_sim[WRITE ? EV_HALF_VID : 1], o_fs[WRITE ? EV_HALF_VID : 1];
  auto consume = [&](float x, int c) {
    const int mem = c % EV_COLS;
    if (mem == 0) { va.reset(); va.vsim = x; } else { va.push(x); }
    if (mem == EV_COLS - 1) {
      const int vl = c / EV_COLS;             // video inside this warp's half
      const int vloc = v0 + vl;               // local video index
      if (WRITE) {
        o_sim[WRITE ? vl : 0] = va.vsim * a.scale;
        o_fs[WRITE ? vl : 0] = va.frame_score(a.scale);
        return;
      }
      const float sc = va.score(a.scale);
      const bool own = (my_grp == a.video_base + vloc);
      if (a.mode == 0) {
        if (own && my_grp >= 0 && vloc < a.Nv_local) a.gt_score[row] = sc;
      } else if (my_grp >= 0 && vloc < a.Nv_local && !own) {
        my_cnt += (sc > my_gt) ? 1 : 0;
        if (sc > a.theta[vloc]) mask |= (1u << vl);
      }
    }
  };
#pragma unroll
  for (int c = 0; c < EV_HALF_COLS / 32; ++c) {
    float v[32];
    ptx::tmem_ld_x32(taddr + c * 32, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) consume(v[j], c * 32 + j);
  }
  {
    float v[8];
    ptx::tmem_ld_x8(taddr + (EV_HALF_COLS / 32) * 32, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) consume(v[j], (EV_HALF_COLS / 32) * 32 + j);
  }
  if (WRITE && row < a.Nt) {
    if (a.combine) {
#pragma unroll
      for (int i = 0; i < EV_HALF_VID; ++i) o_sim[WRITE ? i : 0] += o_fs[WRITE ? i : 0];
    }
    auto put = [&](float* out, const float (&o)[WRITE ? EV_HALF_VID : 1]) {
      if (out == nullptr) return;
      float* dst = out + int64_t(row) * a.ld_out + v0;
      if (v0 + EV_HALF_VID <= a.Nv_local && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[WRITE ? 1 : 0], o[WRITE ? 2 : 0], o[WRITE ? 3 : 0]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(o[WRITE ? 4 : 0], o[WRITE ? 5 : 0], o[WRITE ? 6 : 0], o[WRITE ? 7 : 0]);
      } else {
#pragma unroll
        for (int i = 0; i < EV_HALF_VID; ++i)
          if (v0 + i < a.Nv_local) dst[i] = o[WRITE ? i : 0];
      }
    };
    put(a.sim_out, o_sim);
    if (!a.combine) put(a.fsim_out, o_fs);
  }
  return mask;
}

// v2t: count, per video of this tile, the caption GROUPS with at least one hit (all 256 epilogue
// threads of the CTA take part).  The exchange is skipped when no thread saw a hit (the common case).
__device__ __forceinline__ void eval_tile_v2t(const EvalArgs& a, uint32_t mask, int buf, int trow, int half, int n_blk,
                                              int my_grp, uint32_t (*s_mask)[2][128], const int32_t* s_grp,
                                              int32_t (*s_cnt)[EV_VPT]) {
  s_mask[buf][half][trow] = mask;
  if (half == 0 && trow < EV_VPT) s_cnt[buf][trow] = 0;
  const int any = epi_bar_or(mask != 0 ? 1 : 0);
  if (any) {
    const bool head = (half == 0) && (my_grp >= 0) && (trow == 0 || s_grp[trow - 1] != my_grp);
    if (head) {
      uint32_t m = 0;
      for (int t = trow; t < 128 && s_grp[t] == my_grp; ++t)
        m |= s_mask[buf][0][t] | (s_mask[buf][1][t] << EV_HALF_VID);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        atomicAdd(&s_cnt[buf][b], 1);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0 && trow < EV_VPT) {
      const int c = s_cnt[buf][trow];
      const int vloc = n_blk * EV_VPT + trow;
      if (c != 0 && vloc < a.Nv_local) atomicAdd(&a.v2t_cnt[vloc], c);
    }
  }
}

template <int K, bool WRITE = false>
__global__ void __launch_bounds__(EV_THREADS, 1)
eval_rank_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ EvalArgs a) {
  using Cfg = UmmaCfg<EV_BN, EpiStoreF32>;   // ring geometry only; this kernel has its own epilogue
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * Cfg::STAGE_BYTES);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_base = ptx::smem_u32(bars);
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (STAGES + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 2 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  // epilogue scratch: per-row hit masks (double buffered), group ids, per-tile video counters
  __shared__ uint32_t s_mask[2][2][128];     // [buffer][column half][row]
  __shared__ int32_t s_grp[128];
  __shared__ int32_t s_cnt[2][EV_VPT];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(full_bar(i), 1);
      ptx::mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar(i), 1);
      ptx::mbar_init(tempty_bar(i), EV_EPI_WARPS);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_kb = a.num_seg * a.kb_per_seg;
  // tile enumeration shared by the three roles
  //   mode 1: the gallery is swept in panels small enough to stay in L2 while every CTA runs all
  //           of its caption tiles against the panel:
  //             for (panel) for (m = blockIdx.x; m < num_m_blk; m += gridDim.x) for (n in panel)
  //   mode 0: for (t = blockIdx.x; t < n_diag_tiles; t += gridDim.x) (m, n) = diag_tiles[t]
  //   mode 2: for (t = blockIdx.x; t < num_m_blk * num_n_blk; t += gridDim.x) (m, n) = (t % num_m_blk, t / num_m_blk)
  const int my_m = (a.num_m_blk - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int list_tiles = WRITE ? a.num_m_blk * a.num_n_blk : a.n_diag_tiles;
  const long long tiles_per_cta =
      (a.mode == 1) ? (long long)my_m * a.num_n_blk
                    : (long long)((list_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x));
  const int NP = a.panel_n_blk;
  const long long full_span = (long long)my_m * NP * (a.num_n_blk / (NP > 0 ? NP : 1));
  const int rem_n = (NP > 0) ? a.num_n_blk % NP : 0;
  auto tile_of = [&](long long i, int& m_blk, int& n_blk) {
    if (a.mode == 1) {
      int mi, nj;
      if (i < full_span) {
        const long long per_panel = (long long)my_m * NP;
        const int pnl = int(i / per_panel);
        const int r = int(i - pnl * per_panel);
        mi = r / NP;
        nj = pnl * NP + (r - mi * NP);
      } else {
        const int r = int(i - full_span);
        mi = r / rem_n;
        nj = (a.num_n_blk - rem_n) + (r - mi * rem_n);
      }
      m_blk = int(blockIdx.x) + mi * int(gridDim.x);
      n_blk = nj;
    } else if (WRITE) {
      const int t = int(blockIdx.x) + int(i) * int(gridDim.x);
      m_blk = t % a.num_m_blk;
      n_blk = t / a.num_m_blk;
    } else {
      const int2 t = a.diag_tiles[int(blockIdx.x) + int(i) * int(gridDim.x)];
      m_blk = t.x;
      n_blk = t.y;
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long i = 0; i < tiles_per_cta; ++i) {
        int m_blk, n_blk;
        tile_of(i, m_blk, n_blk);
        for (int kbi = 0; kbi < total_kb; ++kbi) {
          const int seg = kbi / a.kb_per_seg, kb = kbi - seg * a.kb_per_seg;
          const int ak = (seg == 0 ? a.a_k0[0] : (seg == 1 ? a.a_k0[1] : a.a_k0[2])) + kb * UMMA_BK;
          const int bk = (seg == 0 ? a.b_k0[0] : (seg == 1 ? a.b_k0[1] : a.b_k0[2])) + kb * UMMA_BK;
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          ptx::tma_load_2d(sa, &tmA, full_bar(stage), ak, m_blk * UMMA_BM);
          ptx::tma_load_2d(sa + Cfg::A_BYTES, &tmB, full_bar(stage), bk, n_blk * EV_BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UMMA_BM, EV_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long i = 0; i < tiles_per_cta; ++i) {
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kbi = 0; kbi < total_kb; ++kbi) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = ptx::umma_desc_k_sw128(sa);
          const uint64_t bdesc = ptx::umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < UMMA_BK / 16; ++k)
            ptx::umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kbi > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;           // videos [8*half, 8*half+8) of each tile
    const int trow = quad * 32 + lane;          // row inside the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_m = -1;
    int my_grp = -1;
    float my_gt = 0.f;
    int my_cnt = 0;
    int buf = 0;
    for (long long i = 0; i < tiles_per_cta; ++i) {
      int m_blk, n_blk;
      tile_of(i, m_blk, n_blk);
      const int row = m_blk * UMMA_BM + trow;
      if (m_blk != cur_m) {
        if (a.mode == 1 && cur_m >= 0 && my_cnt != 0) atomicAdd(&a.t2v_cnt[cur_m * UMMA_BM + trow], my_cnt);
        cur_m = m_blk;
        my_cnt = 0;
        my_grp = WRITE ? -1 : a.grp[row];
        if (a.mode == 1) {
          my_gt = a.gt_score[row];
          // publish the tile's group ids for the cross-row reduction (all 256 epilogue threads)
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (half == 0) s_grp[trow] = my_grp;
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * Cfg::ACC_STRIDE + half * EV_HALF_COLS;
      const uint32_t mask = eval_tile_columns<K, WRITE>(a, taddr, row, n_blk, half, my_grp, my_gt, my_cnt);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      if (a.mode == 1) {
        eval_tile_v2t(a, mask, buf, trow, half, n_blk, my_grp, s_mask, s_grp, s_cnt);
        buf ^= 1;
      }
    }
    if (a.mode == 1 && cur_m >= 0 && my_cnt != 0) atomicAdd(&a.t2v_cnt[cur_m * UMMA_BM + trow], my_cnt);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ CTA-pair, caption-stationary sweep
// Counting sweep (mode 1, one bf16 plane) on CTA pairs: tcgen05 cta_group::2, M = 256 captions per
// pair (128 per CTA), N = 208.  Each CTA keeps ITS caption tile (128 x D bf16, 128 KB at D = 512)
// resident in shared memory for a whole gallery panel, and streams only HALF of every gallery tile
// (104 rows, 13 KB per k-block): the L2 -> SM traffic per tile drops from 336 KB to 104 KB.
//   warp 0 lane 0 (both CTAs) : TMA producer; completion bytes land on the leader's barriers
//   warp 1 lane 0 (leader)    : tcgen05.mma.cta_group::2 issuer; commits are multicast to both CTAs
//   warp 2 (both)             : TMEM allocation (cta_group::2)
//   warps 4..11 (both)        : epilogue over the CTA's own 128 accumulator rows
constexpr int EVP_STAGES = 6;
constexpr uint32_t EVP_B_BYTES = (EV_BN / 2) * UMMA_BK * 2;      // 104 rows x 128 B = 13 KB
constexpr uint32_t EVP_A_SLICE = UMMA_BM * UMMA_BK * 2;          // 16 KB per 64-wide k-block
constexpr int EVP_MAX_KB = 8;                                    // D <= 512
constexpr size_t EVP_SMEM_BYTES = size_t(EVP_MAX_KB) * EVP_A_SLICE + size_t(EVP_STAGES) * EVP_B_BYTES + 1024 + 256;
static_assert(EVP_B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment");

template <int K>
__global__ void __launch_bounds__(EV_THREADS, 1)
eval_rank_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                      const __grid_constant__ EvalArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_base = ptx::smem_u32(smem);
  const uint32_t b_base = a_base + EVP_MAX_KB * EVP_A_SLICE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(EVP_MAX_KB) * EVP_A_SLICE + size_t(EVP_STAGES) * EVP_B_BYTES);
  const uint32_t bar_base = ptx::smem_u32(bars);
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (EVP_STAGES + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * EVP_STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * EVP_STAGES + 2 + i); };
  const uint32_t afull_bar = bar_base + 8u * (2 * EVP_STAGES + 4);
  const uint32_t aempty_bar = bar_base + 8u * (2 * EVP_STAGES + 5);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * EVP_STAGES + 6);
  __shared__ uint32_t s_mask[2][2][128];
  __shared__ int32_t s_grp[128];
  __shared__ int32_t s_cnt[2][EV_VPT];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = ptx::cluster_ctarank();        // 0 = leader of the pair
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  constexpr uint32_t TMEM_COLS = 512, ACC_STRIDE = 256;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmBh);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < EVP_STAGES; ++i) {
      ptx::mbar_init(full_bar(i), 1);       // leader's producer arrives once; bytes come from both CTAs
      ptx::mbar_init(empty_bar(i), 1);      // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar(i), 1);
      ptx::mbar_init(tempty_bar(i), 2 * EV_EPI_WARPS);   // epilogue warps of BOTH CTAs (used on the leader)
    }
    ptx::mbar_init(afull_bar, 1);
    ptx::mbar_init(aempty_bar, 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();                  // barriers of both CTAs initialised before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kbs = a.kb_per_seg;                                     // one plane: K / 64
  const int num_m2 = (a.num_m_blk + 1) / 2;                         // 256-caption super tiles
  const int my_m2 = (num_m2 - pair + num_pairs - 1) / num_pairs;
  const int NP = a.panel_n_blk;
  const int num_panels = (a.num_n_blk + NP - 1) / NP;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      for (int pnl = 0; pnl < num_panels; ++pnl) {
        const int n0 = pnl * NP, n1 = min(n0 + NP, a.num_n_blk);
        for (int mi = 0; mi < my_m2; ++mi) {
          const int m_blk = (pair + mi * num_pairs) * 2 + int(cta);
          // caption tile: wait until the MMAs of the previous block have released it, then load all k-blocks
          ptx::mbar_wait(aempty_bar, aphase ^ 1u);
          if (cta == 0) ptx::mbar_expect_tx(afull_bar, 2u * kbs * EVP_A_SLICE);
          for (int kb = 0; kb < kbs; ++kb)
            ptx::tma_load_2d_pair(a_base + kb * EVP_A_SLICE, &tmA, afull_bar, kb * UMMA_BK, m_blk * UMMA_BM);
          aphase ^= 1u;
          for (int n = n0; n < n1; ++n) {
            for (int kb = 0; kb < kbs; ++kb) {
              ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
              if (cta == 0) ptx::mbar_expect_tx(full_bar(stage), 2u * EVP_B_BYTES);
              ptx::tma_load_2d_pair(b_base + stage * EVP_B_BYTES, &tmBh, full_bar(stage), kb * UMMA_BK,
                                    n * EV_BN + int(cta) * (EV_BN / 2));
              if (++stage == EVP_STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * UMMA_BM, EV_BN);
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pnl = 0; pnl < num_panels; ++pnl) {
        const int n0 = pnl * NP, n1 = min(n0 + NP, a.num_n_blk);
        for (int mi = 0; mi < my_m2; ++mi) {
          ptx::mbar_wait(afull_bar, aphase);
          aphase ^= 1u;
          ptx::tc_fence_after();
          for (int n = n0; n < n1; ++n) {
            ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
            for (int kb = 0; kb < kbs; ++kb) {
              ptx::mbar_wait(full_bar(stage), phase);
              ptx::tc_fence_after();
              const uint64_t adesc = ptx::umma_desc_k_sw128(a_base + kb * EVP_A_SLICE);
              const uint64_t bdesc = ptx::umma_desc_k_sw128(b_base + stage * EVP_B_BYTES);
#pragma unroll
              for (int k = 0; k < UMMA_BK / 16; ++k)
                ptx::umma_bf16_ss_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              ptx::umma_commit_pair(empty_bar(stage));
              if (++stage == EVP_STAGES) { stage = 0; phase ^= 1u; }
            }
            ptx::umma_commit_pair(tfull_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          }
          ptx::umma_commit_pair(aempty_bar);       // caption tiles of both CTAs may be overwritten
        }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int trow = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    for (int pnl = 0; pnl < num_panels; ++pnl) {
      const int n0 = pnl * NP, n1 = min(n0 + NP, a.num_n_blk);
      for (int mi = 0; mi < my_m2; ++mi) {
        const int m_blk = (pair + mi * num_pairs) * 2 + int(cta);
        const bool live = m_blk < a.num_m_blk;                     // odd tile count: the last pair has one idle CTA
        const int row = m_blk * UMMA_BM + trow;
        const int my_grp = live ? a.grp[row] : -1;
        const float my_gt = live ? a.gt_score[row] : 0.f;
        int my_cnt = 0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) s_grp[trow] = my_grp;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int n = n0; n < n1; ++n) {
          ptx::mbar_wait(tfull_bar(acc), acc_phase);
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * ACC_STRIDE + half * EV_HALF_COLS;
          const uint32_t mask = eval_tile_columns<K, false>(a, taddr, row, n, half, my_grp, my_gt, my_cnt);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tempty_bar(acc), 0);   // the leader's MMA thread waits on it
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          eval_tile_v2t(a, mask, buf, trow, half, n, my_grp, s_mask, s_grp, s_cnt);
          buf ^= 1;
        }
        if (live && my_cnt != 0) atomicAdd(&a.t2v_cnt[row], my_cnt);
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();                  // nobody leaves while the peer may still signal or read
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ packing
// texts: packed row i takes source row src_row[i] (or zeros when src_row[i] < 0), L2-normalised without eps
__global__ void eval_pack_text_kernel(const float* __restrict__ text, const int32_t* __restrict__ src_row, int64_t rows,
                                      int D, int planes, __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t ldp = int64_t(planes) * D;
  const int sr = src_row[r];
  if (sr < 0) {
    for (int d = lane; d < planes * D; d += 32) out[r * ldp + d] = __float2bfloat16_rn(0.f);
    return;
  }
  const float* x = text + int64_t(sr) * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = x[d]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float n = sqrtf(ss);
  for (int d = lane; d < D; d += 32) {
    __nv_bfloat16 hi, lo;
    split_bf16(x[d] / n, hi, lo);
    out[r * ldp + d] = hi;
    if (planes == 2) out[r * ldp + D + d] = lo;
  }
}

// gallery: row v*(1+F)+0 = video v, rows v*(1+F)+1+f = frame f of video v; rows beyond Nv are zeros
// text (optional): Nt more rows after the gallery's, normalised and packed straight into text_out [Nt, planes*D]
// (the same arithmetic as rownorm_pack_kernel) - one launch for both operands of the materialised eval
__global__ void eval_pack_gallery_kernel(const float* __restrict__ video, const float* __restrict__ frames, int64_t Nv,
                                         int64_t rows_pad, int F, int D, int planes, __nv_bfloat16* __restrict__ out,
                                         const float* __restrict__ text, int64_t Nt, __nv_bfloat16* __restrict__ text_out) {
  const int lane = threadIdx.x & 31;
  int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows_pad + Nt) return;
  const int64_t ldp = int64_t(planes) * D;
  const float* x;
  if (r >= rows_pad) {
    r -= rows_pad;
    x = text + r * D;
    out = text_out;
  } else {
    const int64_t v = r / (1 + F);
    const int mem = int(r - v * (1 + F));
    if (v >= Nv) {
      for (int d = lane; d < planes * D; d += 32) out[r * ldp + d] = __float2bfloat16_rn(0.f);
      return;
    }
    x = (mem == 0) ? video + v * D : frames + (v * F + (mem - 1)) * D;
  }
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float t = x[d]; ss = fmaf(t, t, ss); }
  ss = warp_sum(ss);
  const float n = sqrtf(ss);
  for (int d = lane; d < D; d += 32) {
    __nv_bfloat16 hi, lo;
    split_bf16(x[d] / n, hi, lo);
    out[r * ldp + d] = hi;
    if (planes == 2) out[r * ldp + D + d] = lo;
  }
}

// theta[j] = max over the packed captions of local video j of gt_score (a group's captions are contiguous)
__global__ void eval_theta_kernel(const float* __restrict__ gt_score, const int32_t* __restrict__ group_start_packed,
                                  const int32_t* __restrict__ group_count, int Nv_local, float* __restrict__ theta) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Nv_local) return;
  float m = -INFINITY;
  const int s0 = group_start_packed[j];
  for (int s = s0; s < s0 + group_count[j]; ++s) m = fmaxf(m, gt_score[s]);
  theta[j] = m;
}

}  // namespace hmmc

using namespace hmmc;

extern "C" {

int hmmc_eval_fused_supported(int F, int D, int top_k) {
  return (F == EV_F && D % UMMA_BK == 0 && top_k >= 1 && top_k <= EV_MAXK) ? 1 : 0;
}

int hmmc_eval_pack_text(const float* text, const int32_t* src_row, int64_t rows_pad, int D, int prec, void* out,
                        void* stream) {
  HMMC_REQUIRE(text && src_row && out && rows_pad > 0 && rows_pad % UMMA_BM == 0, "eval_pack_text: bad arguments");
  HMMC_REQUIRE(prec == HMMC_PREC_BF16 || prec == HMMC_PREC_BF16X3, "eval_pack_text: tensor-core precisions only");
  eval_pack_text_kernel<<<unsigned((rows_pad + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      text, src_row, rows_pad, D, planes_of(prec), static_cast<__nv_bfloat16*>(out));
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_eval_pack_gallery(const float* video, const float* frames, int64_t Nv, int F, int D, int prec, void* out,
                           void* stream) {
  HMMC_REQUIRE(video && frames && out && Nv > 0, "eval_pack_gallery: bad arguments");
  HMMC_REQUIRE(F == EV_F, "eval_pack_gallery: the fused path handles F = %d frames (got %d)", EV_F, F);
  HMMC_REQUIRE(prec == HMMC_PREC_BF16 || prec == HMMC_PREC_BF16X3, "eval_pack_gallery: tensor-core precisions only");
  const int64_t nblk = (Nv + EV_VPT - 1) / EV_VPT;
  const int64_t rows_pad = nblk * EV_BN;
  eval_pack_gallery_kernel<<<unsigned((rows_pad + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      video, frames, Nv, rows_pad, F, D, planes_of(prec), static_cast<__nv_bfloat16*>(out), nullptr, 0, nullptr);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

static int eval_launch(const void* text_packed, const void* gallery_packed, int64_t Nt_pad, int64_t Nv_local, int D,
                       int prec, EvalArgs& a, cudaStream_t st, int64_t text_rows = 0) {
  HMMC_REQUIRE(text_packed && gallery_packed, "eval: null operand");
  HMMC_REQUIRE(Nt_pad > 0 && Nt_pad % UMMA_BM == 0, "eval: Nt_pad must be a positive multiple of %d", UMMA_BM);
  HMMC_REQUIRE(D % UMMA_BK == 0, "eval: D %% 64 != 0");
  HMMC_REQUIRE(prec == HMMC_PREC_BF16 || prec == HMMC_PREC_BF16X3, "eval: tensor-core precisions only");
  HMMC_REQUIRE(a.top_k >= 1 && a.top_k <= EV_MAXK, "eval: fused path supports 1 <= top_k <= %d", EV_MAXK);
  const int planes = planes_of(prec);
  a.num_m_blk = int(Nt_pad / UMMA_BM);
  a.num_n_blk = int((Nv_local + EV_VPT - 1) / EV_VPT);
  a.Nv_local = int(Nv_local);
  GemmShape s;
  fill_segments(s, planes, D);
  a.num_seg = s.num_seg;
  a.kb_per_seg = s.kb_per_seg;
  for (int i = 0; i < 3; ++i) { a.a_k0[i] = s.a_k0[i]; a.b_k0[i] = s.b_k0[i]; }
  CUtensorMap tmA, tmB;
  // text_rows > 0: the operand has fewer rows than the tile grid covers, TMA zero-fills the rest
  int rc = make_tmap_bf16(&tmA, text_packed, uint64_t(text_rows > 0 ? text_rows : Nt_pad), uint64_t(planes) * D,
                          uint64_t(planes) * D, UMMA_BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, gallery_packed, uint64_t(a.num_n_blk) * EV_BN, uint64_t(planes) * D, uint64_t(planes) * D, EV_BN);
  if (rc) return rc;
  using Cfg = UmmaCfg<EV_BN, EpiStoreF32>;   // ring geometry only; this kernel has its own epilogue
  const int work = (a.mode == 1) ? a.num_m_blk : (a.mode == 2 ? a.num_m_blk * a.num_n_blk : a.n_diag_tiles);
  if (work <= 0) return HMMC_OK;
  // counting sweep in one bf16 plane: CTA pairs with the caption tile resident in shared memory
  if (a.mode == 1 && planes == 1 && D / UMMA_BK <= EVP_MAX_KB) {
    CUtensorMap tmBh;
    rc = make_tmap_bf16(&tmBh, gallery_packed, uint64_t(a.num_n_blk) * EV_BN, uint64_t(D), uint64_t(D), EV_BN / 2);
    if (rc) return rc;
    const int num_m2 = (a.num_m_blk + 1) / 2;
    int pairs = sm_count() / 2;
    if (pairs > num_m2) pairs = num_m2;
    auto launch_pair = [&](auto kern) -> int {
      HMMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(EVP_SMEM_BYTES)));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * pairs);
      cfg.blockDim = dim3(EV_THREADS);
      cfg.dynamicSmemBytes = EVP_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      count_launch();
      HMMC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmBh, a));
      return HMMC_OK;
    };
    switch (a.top_k) {
      case 1: return launch_pair(eval_rank_pair_kernel<1>);
      case 2: return launch_pair(eval_rank_pair_kernel<2>);
      case 3: return launch_pair(eval_rank_pair_kernel<3>);
      default: return launch_pair(eval_rank_pair_kernel<4>);
    }
  }
  const int grid = work < sm_count() ? work : sm_count();
  auto launch = [&](auto kern) -> int {
    HMMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM_BYTES)));
    kern<<<grid, EV_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, a);
    HMMC_CHECK_LAUNCH();
    return HMMC_OK;
  };
  if (a.mode == 2) {
    switch (a.top_k) {
      case 1: return launch(eval_rank_kernel<1, true>);
      case 2: return launch(eval_rank_kernel<2, true>);
      case 3: return launch(eval_rank_kernel<3, true>);
      default: return launch(eval_rank_kernel<4, true>);
    }
  }
  switch (a.top_k) {
    case 1: return launch(eval_rank_kernel<1>);
    case 2: return launch(eval_rank_kernel<2>);
    case 3: return launch(eval_rank_kernel<3>);
    default: return launch(eval_rank_kernel<4>);
  }
}

}  // extern "C"

namespace hmmc {
// Materialised scores through the same tiles: sim / fsim [Nt, Nv] from packed operands (text: straight
// [Nt, planes*D] rows, tensor map clipped at Nt; gallery: hmmc_eval_pack_gallery layout).  Internal: called by
// hmmc_sim_topk_fwd (eval.cu) for the shapes the fused tiles cover.
int eval_sim_write(const void* text_packed, const void* gallery_packed, int64_t Nt, int64_t Nv, int D, int prec,
                   float scale, int top_k, float* sim, float* fsim, int64_t ld_out, int combine, cudaStream_t st) {
  EvalArgs a{};
  a.scale = scale;
  a.top_k = top_k;
  a.mode = 2;
  a.sim_out = sim;
  a.fsim_out = fsim;
  a.ld_out = ld_out;
  a.Nt = int(Nt);
  a.combine = combine;
  const int64_t Nt_pad = (Nt + UMMA_BM - 1) / UMMA_BM * UMMA_BM;
  return eval_launch(text_packed, gallery_packed, Nt_pad, Nv, D, prec, a, st, Nt);
}

size_t eval_gallery_pack_rows(int64_t Nv) { return size_t((Nv + EV_VPT - 1) / EV_VPT) * EV_BN; }

int eval_pack_gallery(const float* video, const float* frames, int64_t Nv, int F, int D, int planes, void* out,
                      cudaStream_t st, const float* text, int64_t Nt, void* text_out) {
  const int64_t rows_pad = int64_t(eval_gallery_pack_rows(Nv));
  if (text == nullptr) Nt = 0;
  eval_pack_gallery_kernel<<<unsigned((rows_pad + Nt + 7) / 8), 256, 0, st>>>(video, frames, Nv, rows_pad, F, D, planes,
                                                                             static_cast<__nv_bfloat16*>(out), text, Nt,
                                                                             static_cast<__nv_bfloat16*>(text_out));
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}
}  // namespace hmmc

extern "C" {

int hmmc_eval_gt_scores(const void* text_packed, const void* gallery_packed, int64_t Nt_pad, int64_t Nv_local, int D,
                        int prec, float scale, int top_k, int video_base, const int32_t* grp,
                        const int32_t* diag_tiles, int n_diag_tiles, float* gt_score, void* stream) {
  HMMC_REQUIRE(grp && gt_score && (n_diag_tiles == 0 || diag_tiles), "eval_gt_scores: null argument");
  EvalArgs a{};
  a.scale = scale;
  a.top_k = top_k;
  a.mode = 0;
  a.video_base = video_base;
  a.grp = grp;
  a.gt_score = gt_score;
  a.diag_tiles = reinterpret_cast<const int2*>(diag_tiles);
  a.n_diag_tiles = n_diag_tiles;
  return eval_launch(text_packed, gallery_packed, Nt_pad, Nv_local, D, prec, a, static_cast<cudaStream_t>(stream));
}

int hmmc_eval_theta(const float* gt_score, const int32_t* group_start_packed, const int32_t* group_count, int Nv_local,
                    float* theta, void* stream) {
  HMMC_REQUIRE(gt_score && group_start_packed && group_count && theta && Nv_local > 0, "eval_theta: bad arguments");
  eval_theta_kernel<<<(Nv_local + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gt_score, group_start_packed, group_count, Nv_local, theta);
  HMMC_CHECK_LAUNCH();
  return HMMC_OK;
}

int hmmc_eval_fused_rank(const void* text_packed, const void* gallery_packed, int64_t Nt_pad, int64_t Nv_local, int D,
                         int prec, float scale, int top_k, int video_base, const int32_t* grp, const float* gt_score,
                         const float* theta, int32_t* t2v_cnt, int32_t* v2t_cnt, void* stream) {
  HMMC_REQUIRE(grp && gt_score && theta && t2v_cnt && v2t_cnt, "eval_fused_rank: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMMC_CHECK_CUDA(cudaMemsetAsync(v2t_cnt, 0, sizeof(int32_t) * size_t(Nv_local), st));
  HMMC_CHECK_CUDA(cudaMemsetAsync(t2v_cnt, 0, sizeof(int32_t) * size_t(Nt_pad), st));
  EvalArgs a{};
  a.scale = scale;
  a.top_k = top_k;
  a.mode = 1;
  a.video_base = video_base;
  a.grp = grp;
  a.gt_score = const_cast<float*>(gt_score);
  a.theta = theta;
  a.t2v_cnt = t2v_cnt;
  a.v2t_cnt = v2t_cnt;
  // panel = as many gallery tiles as fit comfortably in L2 next to the live caption tiles (~48 MB)
  {
    const int64_t tile_bytes = int64_t(EV_BN) * planes_of(prec) * D * 2;
    int np = int((int64_t(48) << 20) / tile_bytes);
    if (np < 8) np = 8;
    a.panel_n_blk = np;
  }
  return eval_launch(text_packed, gallery_packed, Nt_pad, Nv_local, D, prec, a, st);
}

}  // extern "C"

// tcgen05 / TMA GEMM engine:  acc[m,n] = sum_k A[m,k] * B[n,k]   (both operands K-major bf16)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma, M=128 x N=BN x K=16, fp32 accumulators in TMEM)
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue      (tcgen05.ld: one accumulator row per thread -> fused epilogue;
//                                  two warps per TMEM lane quadrant, each takes half of the columns)
// Two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// The K loop runs over up to three "segments" (pairs of K offsets into the plane-packed
// operands), which is how the BF16X3 mode (hi*hi + hi*lo + lo*hi) is expressed without
// duplicating data, and over a [k-block) sub-range per tile for split-K.
//
// Work distribution: a grouped launch is a list of units (output tile x split-K slice) of different
// lengths.  The host assigns them to the CTAs (or CTA pairs) longest-first onto the least-loaded one and
// passes the schedule in kernel-parameter space, so a launch that is a single wave of long split-K units
// (U = E.Q^T) finishes everywhere at the same time instead of waiting for the CTAs that drew two long units.
//
// The epilogue is a functor (Epi) that sees 32 consecutive accumulator columns of one row
// at a time; see the Epi* structs below.
//
// Every kernel here is launched with programmatic stream serialisation: barrier set-up, TMEM allocation and
// tensor-map prefetch run while the previous kernel of the stream drains; griddepcontrol.wait precedes the
// first access to global memory.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"
#include <algorithm>
#include <vector>

namespace hmmc {

struct GemmShape {
  int M, N;                 // logical output extent (rows / columns beyond are masked)
  int num_m_blk, num_n_blk, num_splits;
  int num_seg;              // 1 or 3
  int kb_per_seg;           // K / 64
  int kb_per_split;         // k-blocks (over all segments) handled by one split
  int a_k0[3], b_k0[3];     // element offsets of each segment on the packed K axis
};

constexpr int UMMA_BM = 128;
constexpr int UMMA_BK = 64;
constexpr int UMMA_EPI_WARPS = 8;                       // two warps per TMEM lane quadrant, half the columns each
constexpr int UMMA_THREADS = (4 + UMMA_EPI_WARPS) * 32;

template <int BN, class Epi>
struct UmmaCfg {
  static constexpr uint32_t A_BYTES = UMMA_BM * UMMA_BK * 2;
  static constexpr uint32_t B_BYTES = BN * UMMA_BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // ring depth: 6 stages of 32 KB / 4 of 48 KB, fewer when the epilogue's staging area needs the room
  static constexpr int WANT_STAGES = (BN > 128) ? 4 : 6;
  static constexpr int FIT_STAGES = int((232448 - 2048 - UMMA_EPI_WARPS * Epi::STAGE_BYTES) / STAGE_BYTES);
  static constexpr int STAGES = WANT_STAGES < FIT_STAGES ? WANT_STAGES : FIT_STAGES;
  static_assert(STAGES >= 3, "TMA ring too shallow");
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t ACC_STRIDE = (BN <= 128) ? 128 : 256;
  static constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;               // 256 B of mbarriers + the TMEM slot
  static constexpr uint32_t EPI_OFF = BAR_OFF + 1024;                     // epilogue staging, 1024-byte aligned
  static constexpr size_t SMEM_BYTES = size_t(EPI_OFF) + 1024 /*align slack*/ + UMMA_EPI_WARPS * Epi::STAGE_BYTES;
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
};

constexpr int UMMA_MAX_PROBLEMS = 6;
constexpr int UMMA_MAX_UNITS = 2048;      // units a launch can place explicitly; larger launches go round-robin
constexpr int UMMA_MAX_WORKERS = 160;     // CTAs (single-CTA kernel) or CTA pairs of a launch

// Tensor maps of a grouped launch (kept in kernel-parameter space: TMA reads them from there).
// c = the epilogue's output map (TMA store), for epilogues that use one.
struct TmapSet {
  CUtensorMap a[UMMA_MAX_PROBLEMS];
  CUtensorMap b[UMMA_MAX_PROBLEMS];
  CUtensorMap c[UMMA_MAX_PROBLEMS];
};

// Which units worker w (a CTA or a CTA pair) processes, in order.
struct UnitSchedule {
  int explicit_order;                           // 0: worker w takes units w, w + workers, ...
  int num_units;
  uint16_t begin[UMMA_MAX_WORKERS + 1];
  uint16_t order[UMMA_MAX_UNITS];
  __device__ __forceinline__ int count(int w, int workers) const {
    return explicit_order ? int(begin[w + 1]) - int(begin[w]) : (num_units - w + workers - 1) / workers;
  }
  __device__ __forceinline__ int unit(int w, int workers, int i) const {
    return explicit_order ? int(order[int(begin[w]) + i]) : w + i * workers;
  }
};

// A grouped launch = up to UMMA_MAX_PROBLEMS independent GEMMs sharing one persistent grid.
// Units are numbered problem after problem: unit = tile_begin[p] + split * (m tiles * n tiles) + n_blk * m tiles + m_blk.
template <class Epi>
struct GroupedArgs {
  int num_problems;
  int tile_begin[UMMA_MAX_PROBLEMS + 1];
  GemmShape shape[UMMA_MAX_PROBLEMS];
  typename Epi::Params ep[UMMA_MAX_PROBLEMS];
  UnitSchedule sched;
};

template <class Epi>
__device__ __forceinline__ int find_problem(const GroupedArgs<Epi>& g, int t) {
  int p = 0;
#pragma unroll
  for (int i = 1; i < UMMA_MAX_PROBLEMS; ++i)
    if (i < g.num_problems && t >= g.tile_begin[i]) p = i;
  return p;
}

// The epilogue's walk over the 32-column chunks of its half tile, unrolled with the chunk index as a
// compile-time constant (epilogues pick staging buffers by it): the TMEM load of chunk C+1 is in flight while
// chunk C is processed.
template <int C, int NC, class Epi>
__device__ __forceinline__ void epi_chunks(Epi& epi, const typename Epi::Params& ep, const GemmShape& s, int row,
                                           int col_base, uint32_t taddr, float (&v)[2][32]) {
  ptx::tmem_ld_wait();
  if constexpr (C + 1 < NC) ptx::tmem_ld_x32(taddr + (C + 1) * 32, v[(C + 1) & 1]);
  epi.template chunk<C, NC>(ep, s, row, col_base + C * 32, v[C & 1]);
  if constexpr (C + 1 < NC) epi_chunks<C + 1, NC>(epi, ep, s, row, col_base, taddr, v);
}

template <int BN, class Epi>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ TmapSet tm, const __grid_constant__ GroupedArgs<Epi> g) {
  using Cfg = UmmaCfg<BN, Epi>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_base = ptx::smem_u32(bars);
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (STAGES + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 2 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < g.num_problems; ++p) {
      ptx::prefetch_tmap(&tm.a[p]);
      ptx::prefetch_tmap(&tm.b[p]);
      if (Epi::USES_CMAP) ptx::prefetch_tmap(&tm.c[p]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(full_bar(i), 1);
      ptx::mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar(i), 1);
      ptx::mbar_init(tempty_bar(i), UMMA_EPI_WARPS);   // one arrival per epilogue warp
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
  // everything above overlapped the previous kernel's tail; its results are needed from here on
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();

  const int workers = gridDim.x;
  const int w = blockIdx.x;
  const int my_units = g.sched.count(w, workers);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = 0; u < my_units; ++u) {
        const int t = g.sched.unit(w, workers, u);
        const int p = find_problem(g, t);
        const GemmShape& s = g.shape[p];
        const int tl = t - g.tile_begin[p];
        const int tiles_mn = s.num_m_blk * s.num_n_blk;
        const int split = tl / tiles_mn;
        const int mn = tl - split * tiles_mn;
        const int m_blk = mn % s.num_m_blk, n_blk = mn / s.num_m_blk;
        const int kb0 = split * s.kb_per_split;
        const int kb1 = min(kb0 + s.kb_per_split, s.num_seg * s.kb_per_seg);
        for (int i = kb0; i < kb1; ++i) {
          const int seg = i / s.kb_per_seg, kb = i - seg * s.kb_per_seg;
          const int ak = (seg == 0 ? s.a_k0[0] : (seg == 1 ? s.a_k0[1] : s.a_k0[2])) + kb * UMMA_BK;
          const int bk = (seg == 0 ? s.b_k0[0] : (seg == 1 ? s.b_k0[1] : s.b_k0[2])) + kb * UMMA_BK;
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          ptx::tma_load_2d(sa, &tm.a[p], full_bar(stage), ak, m_blk * UMMA_BM);
          ptx::tma_load_2d(sa + Cfg::A_BYTES, &tm.b[p], full_bar(stage), bk, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(UMMA_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = 0; u < my_units; ++u) {
        const int t = g.sched.unit(w, workers, u);
        const int p = find_problem(g, t);
        const GemmShape& s = g.shape[p];
        const int split = (t - g.tile_begin[p]) / (s.num_m_blk * s.num_n_blk);
        const int kb0 = split * s.kb_per_split;
        const int kb1 = min(kb0 + s.kb_per_split, s.num_seg * s.kb_per_seg);
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int i = kb0; i < kb1; ++i) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = ptx::umma_desc_k_sw128(sa);
          const uint64_t bdesc = ptx::umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < UMMA_BK / 16; ++k) {
            // +32 bytes along K inside the 128-byte swizzle span = +2 in the (addr>>4) field
            ptx::umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (i > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));   // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(tfull_bar(acc));       // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;           // which half of the tile's columns
    constexpr int HALF_N = BN / 2;
    static_assert(HALF_N % 32 == 0, "epilogue halves must be whole 32-column chunks");
    int acc = 0;
    uint32_t acc_phase = 0;
    Epi epi;
    // per-warp staging buffer in shared memory (how it is used is the epilogue's business)
    epi.stage = smem + Cfg::EPI_OFF + size_t(warp - 4) * Epi::STAGE_BYTES;
    epi.lane = lane;
    epi.init();
    for (int u = 0; u < my_units; ++u) {
      const int t = g.sched.unit(w, workers, u);
      const int p = find_problem(g, t);
      const GemmShape& s = g.shape[p];
      const typename Epi::Params& ep = g.ep[p];
      const int tl = t - g.tile_begin[p];
      const int tiles_mn = s.num_m_blk * s.num_n_blk;
      const int split = tl / tiles_mn;
      const int mn = tl - split * tiles_mn;
      const int m_blk = mn % s.num_m_blk, n_blk = mn / s.num_m_blk;
      const int row = m_blk * UMMA_BM + quad * 32 + lane;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * Cfg::ACC_STRIDE + half * HALF_N;
      epi.begin_tile(ep, s, row, n_blk, split, &tm.c[p]);
      {
        float v[2][32];
        ptx::tmem_ld_x32(taddr, v[0]);
        epi_chunks<0, HALF_N / 32>(epi, ep, s, row, n_blk * BN + half * HALF_N, taddr, v);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      epi.end_tile(ep, s, row, n_blk * 2 + half, split);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    epi.finish();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Same roles and epilogues, but two CTAs of a cluster share every 256(M) x 256(N) tile: each CTA loads
// its own 128 rows of A and HALF of the B tile (128 rows), the leader's thread issues
// tcgen05.mma.cta_group::2 (M = 256) and multicasts its commits to both CTAs, and each CTA's epilogue
// warps drain their own 128 accumulator rows.  A stage is 32 KB instead of 48 KB, so six stages fit:
// half the L2 -> SM bytes per FLOP and 1.5x more latency tolerance for the TMA ring.
// The number of epilogue warps (8, or 16 = four per TMEM lane quadrant for epilogues that are latency-bound at
// two warps per scheduler) and the depth of the TMA ring are properties of the epilogue (Epi::PAIR_WARPS,
// Epi::PAIR_STAGES): a deeper ring or a larger store staging area, whichever the 227 KB pay for best.
constexpr int UMMA_PAIR_BN = 256;
constexpr uint32_t UMMA_PAIR_A_BYTES = UMMA_BM * UMMA_BK * 2;               // 16 KB
constexpr uint32_t UMMA_PAIR_B_BYTES = (UMMA_PAIR_BN / 2) * UMMA_BK * 2;    // 16 KB
constexpr uint32_t UMMA_PAIR_STAGE_BYTES = UMMA_PAIR_A_BYTES + UMMA_PAIR_B_BYTES;
template <class Epi>
struct UmmaPairCfg {
  static constexpr int STAGES = Epi::PAIR_STAGES;
  static constexpr int EPI_WARPS = Epi::PAIR_WARPS;
  static constexpr int THREADS = (4 + EPI_WARPS) * 32;
  static constexpr uint32_t BAR_OFF = STAGES * UMMA_PAIR_STAGE_BYTES;
  static constexpr uint32_t EPI_OFF = BAR_OFF + 1024;
  static constexpr size_t SMEM_BYTES = size_t(EPI_OFF) + 1024 + EPI_WARPS * Epi::STAGE_BYTES;
  static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "two or four epilogue warps per TMEM lane quadrant");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
};

template <class Epi>
__global__ void __launch_bounds__(UmmaPairCfg<Epi>::THREADS, 1)
umma_gemm_pair_kernel(const __grid_constant__ TmapSet tm, const __grid_constant__ GroupedArgs<Epi> g) {
  using Cfg = UmmaPairCfg<Epi>;
  constexpr int BN = UMMA_PAIR_BN;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int EPI_WARPS = Cfg::EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_base = ptx::smem_u32(bars);
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (STAGES + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * STAGES + 2 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  constexpr uint32_t TMEM_COLS = 512, ACC_STRIDE = 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = ptx::cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < g.num_problems; ++p) {
      ptx::prefetch_tmap(&tm.a[p]);
      ptx::prefetch_tmap(&tm.b[p]);
      if (Epi::USES_CMAP) ptx::prefetch_tmap(&tm.c[p]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(full_bar(i), 1);      // the leader's producer arrives; bytes come from both CTAs
      ptx::mbar_init(empty_bar(i), 1);     // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar(i), 1);
      ptx::mbar_init(tempty_bar(i), 2 * EPI_WARPS);        // epilogue warps of both CTAs (leader's copy is used)
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();

  // unit t of a problem = (m2, n_blk, split) with m2 indexing 256-row super tiles; tile_begin[] was
  // built with num_m_blk = number of super tiles (see launch_umma_grouped_pair)
  const int my_units = g.sched.count(pair, num_pairs);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = 0; u < my_units; ++u) {
        const int t = g.sched.unit(pair, num_pairs, u);
        const int p = find_problem(g, t);
        const GemmShape& s = g.shape[p];
        const int tl = t - g.tile_begin[p];
        const int tiles_mn = s.num_m_blk * s.num_n_blk;
        const int split = tl / tiles_mn;
        const int mn = tl - split * tiles_mn;
        const int m2 = mn % s.num_m_blk, n_blk = mn / s.num_m_blk;
        const int kb0 = split * s.kb_per_split;
        const int kb1 = min(kb0 + s.kb_per_split, s.num_seg * s.kb_per_seg);
        for (int i = kb0; i < kb1; ++i) {
          const int seg = i / s.kb_per_seg, kb = i - seg * s.kb_per_seg;
          const int ak = (seg == 0 ? s.a_k0[0] : (seg == 1 ? s.a_k0[1] : s.a_k0[2])) + kb * UMMA_BK;
          const int bk = (seg == 0 ? s.b_k0[0] : (seg == 1 ? s.b_k0[1] : s.b_k0[2])) + kb * UMMA_BK;
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (cta == 0) ptx::mbar_expect_tx(full_bar(stage), 2u * UMMA_PAIR_STAGE_BYTES);
          const uint32_t sa = smem_base + stage * UMMA_PAIR_STAGE_BYTES;
          ptx::tma_load_2d_pair(sa, &tm.a[p], full_bar(stage), ak, (m2 * 2 + int(cta)) * UMMA_BM);
          ptx::tma_load_2d_pair(sa + UMMA_PAIR_A_BYTES, &tm.b[p], full_bar(stage), bk, n_blk * BN + int(cta) * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * UMMA_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = 0; u < my_units; ++u) {
        const int t = g.sched.unit(pair, num_pairs, u);
        const int p = find_problem(g, t);
        const GemmShape& s = g.shape[p];
        const int split = (t - g.tile_begin[p]) / (s.num_m_blk * s.num_n_blk);
        const int kb0 = split * s.kb_per_split;
        const int kb1 = min(kb0 + s.kb_per_split, s.num_seg * s.kb_per_seg);
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int i = kb0; i < kb1; ++i) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * UMMA_PAIR_STAGE_BYTES;
          const uint64_t adesc = ptx::umma_desc_k_sw128(sa);
          const uint64_t bdesc = ptx::umma_desc_k_sw128(sa + UMMA_PAIR_A_BYTES);
#pragma unroll
          for (int k = 0; k < UMMA_BK / 16; ++k)
            ptx::umma_bf16_ss_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (i > kb0 || k > 0) ? 1u : 0u);
          ptx::umma_commit_pair(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit_pair(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
    constexpr int PARTS = EPI_WARPS / 4;        // warps per quadrant: each takes BN / PARTS columns
    const int part = (warp - 4) >> 2;
    constexpr int PART_N = BN / PARTS;
    int acc = 0;
    uint32_t acc_phase = 0;
    Epi epi;
    epi.stage = smem + Cfg::EPI_OFF + size_t(warp - 4) * Epi::STAGE_BYTES;
    epi.lane = lane;
    epi.init();
    for (int u = 0; u < my_units; ++u) {
      const int t = g.sched.unit(pair, num_pairs, u);
      const int p = find_problem(g, t);
      const GemmShape& s = g.shape[p];
      const typename Epi::Params& ep = g.ep[p];
      const int tl = t - g.tile_begin[p];
      const int tiles_mn = s.num_m_blk * s.num_n_blk;
      const int split = tl / tiles_mn;
      const int mn = tl - split * tiles_mn;
      const int m2 = mn % s.num_m_blk, n_blk = mn / s.num_m_blk;
      const int row = (m2 * 2 + int(cta)) * UMMA_BM + quad * 32 + lane;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * ACC_STRIDE + part * PART_N;
      epi.begin_tile(ep, s, row, n_blk, split, &tm.c[p]);
      {
        float v[2][32];
        ptx::tmem_ld_x32(taddr, v[0]);
        epi_chunks<0, PART_N / 32>(epi, ep, s, row, n_blk * BN + part * PART_N, taddr, v);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tempty_bar(acc), 0);
      epi.end_tile(ep, s, row, n_blk * PARTS + part, split);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    epi.finish();
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ epilogues

// C[split][m, n] = alpha * acc
struct EpiStoreF32 {
  struct Params {
    float* C;
    int64_t ldc;
    int64_t split_stride;
    float alpha;
  };
  static constexpr uint32_t STAGE_BYTES = 32 * 80;    // per warp: 32 rows x (64 + 16 pad) bytes
  static constexpr bool USES_CMAP = false;
  static constexpr int PAIR_WARPS = 8, PAIR_STAGES = 6;
  static constexpr int PARTS_PER_TILE_PAIR = 2;       // partial results per 256-wide tile (CTA-pair kernel)
  static constexpr int STORE_COLS = 64;               // (no TMA stores)
  uint8_t* stage;
  int lane;
  float* tile_out;   // &C[split][row of lane 0 of this warp, 0]
  int row0;
  bool vec_ok;
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ void finish() {}
  __device__ __forceinline__ void begin_tile(const Params& p, const GemmShape& s, int row, int n_blk, int split,
                                             const CUtensorMap*) {
    row0 = row - lane;
    tile_out = p.C + int64_t(split) * p.split_stride + int64_t(row0) * p.ldc;
    vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.split_stride & 3) == 0);
  }
  template <int C, int NC>
  __device__ __forceinline__ void chunk(const Params& p, const GemmShape& s, int row, int col0, float (&v)[32]) {
    // two 16-column halves: rows staged at an 80-byte pitch, then each instruction stores eight
    // 64-byte row segments
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4* srow = reinterpret_cast<float4*>(stage + lane * 80);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        srow[j] = make_float4(v[16 * h + 4 * j] * p.alpha, v[16 * h + 4 * j + 1] * p.alpha,
                              v[16 * h + 4 * j + 2] * p.alpha, v[16 * h + 4 * j + 3] * p.alpha);
      __syncwarp();
      const int c4 = (lane & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        const float4 x = *reinterpret_cast<const float4*>(stage + r * 80 + c4 * 4);
        const int grow = row0 + r, gcol = col0 + 16 * h + c4;
        if (grow < s.M) {
          float* o = tile_out + int64_t(r) * p.ldc + gcol;
          if (vec_ok && gcol + 3 < s.N) {
            *reinterpret_cast<float4*>(o) = x;
          } else {
            if (gcol < s.N) o[0] = x.x;
            if (gcol + 1 < s.N) o[1] = x.y;
            if (gcol + 2 < s.N) o[2] = x.z;
            if (gcol + 3 < s.N) o[3] = x.w;
          }
        }
      }
      __syncwarp();
    }
  }
  __device__ __forceinline__ void end_tile(const Params&, const GemmShape&, int, int, int) {}
};

// InfoNCE negatives.  The A operand (the normalised queries) was scaled by log2(e)/T when it was packed, so
// the accumulator is the logit in base-2 units and e = 2^acc needs one MUFU per element and nothing else; the
// constant maximum of the log-sum-exp (all logits <= 1/T) is applied by the finish kernel as one factor
// 2^(-c2) on the row sums and on U (2^(2/T log2 e) still fits fp32 and bf16 comfortably for T >= 0.025).
// Per-row partial sums -> rowsum_part[part][row]; bf16 hi(/lo) planes of e -> E through TMA stores: a warp
// packs GROUP 32-column chunks into a 32-row staging tile (64B / 128B swizzle, conflict-free 16-byte stores,
// NBUF tiles in flight) and one lane hands it to the copy engine, which also clips rows beyond M.
// No logits are written.  Requires N % 64 == 0 (checked by the host).
// PLANES: 0 = forward only (no E), 1 = bf16, 2 = bf16x3 (hi and lo planes).
// Shape of the epilogue in the CTA-pair kernel (measured on B200 at b = 256, profiles/r2_epilogue_sweep.md):
// WARPS epilogue warps, STAGES TMA ring stages, GROUP chunks per store, NBUF staging tiles per warp.
#ifndef HMMC_S_WARPS
#define HMMC_S_WARPS 8
#define HMMC_S_STAGES 6
#define HMMC_S_GROUP 1
#define HMMC_S_NBUF 2
#endif
template <int PLANES, int WARPS = (PLANES == 2 ? 8 : HMMC_S_WARPS), int STAGES = HMMC_S_STAGES,
          int GROUP = (PLANES == 2 ? 1 : HMMC_S_GROUP), int NBUF = (PLANES == 2 ? 2 : HMMC_S_NBUF)>
struct EpiInfoNCE {
  struct Params {
    float* rowsum_part;       // [parts per tile * num_n_blk, M]: one partial per epilogue warp column range
    int lo_col0;              // column offset of the lo plane inside the E tensor map (= N)
  };
  static constexpr int PAIR_WARPS = WARPS;
  static constexpr int PAIR_STAGES = STAGES;
  static constexpr int PARTS_PER_TILE_PAIR = WARPS / 4;
  static constexpr int STORE_COLS = 32 * GROUP;                  // columns per TMA store (tensor-map box)
  static constexpr uint32_t TILE_BYTES = 32 * STORE_COLS * 2;    // 32 rows of 64 / 128 bytes
  static constexpr uint32_t STAGE_BYTES = NBUF * TILE_BYTES;     // per warp
  static constexpr bool USES_CMAP = PLANES != 0;
  static_assert(GROUP == 1 || GROUP == 2, "one or two chunks per store");
  uint8_t* stage;
  int lane;
  float s0, s1, s2, s3;
  int row0;
  uint32_t stage_addr;
  uint32_t slot0;             // this lane's 16-byte slot 0 of a staging tile; slot c is at slot0 ^ (c << 4)
  uint32_t hold_hi[16], hold_lo[16];      // first chunk of a two-chunk group (dead unless GROUP == 2)
  const CUtensorMap* cmap;
  uint64_t keep;              // L2 policy of the E stores: the U-GEMM reads E right back, keep it out of HBM if possible
  __device__ __forceinline__ void init() {
    stage_addr = ptx::smem_u32(stage);
    keep = ptx::l2_policy_evict_last();
    // swizzle: the 16-byte chunk index is XORed with the row index (64B rows: bits 1..2 of the row; 128B rows: bits 0..2)
    slot0 = (GROUP == 1) ? stage_addr + uint32_t(lane) * 64u + (uint32_t((lane >> 1) & 3) << 4)
                         : stage_addr + uint32_t(lane) * 128u + (uint32_t(lane & 7) << 4);
  }
  __device__ __forceinline__ void finish() {
    if (PLANES != 0) {
      if (lane == 0) ptx::bulk_wait_group_read<0>();   // the copy engine still reads this CTA's shared memory
      __syncwarp();
    }
  }
  __device__ __forceinline__ void begin_tile(const Params&, const GemmShape&, int row, int, int, const CUtensorMap* c) {
    s0 = s1 = s2 = s3 = 0.f;
    row0 = row - lane;
    cmap = c;
  }
  template <int SLOT>      // 16-byte slots SLOT .. SLOT+3 of this lane's row in staging tile BUF
  __device__ __forceinline__ void stage4(const uint32_t (&w)[16], uint32_t buf_off) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((slot0 ^ uint32_t((SLOT + c) << 4)) + buf_off),
                   "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3])
                   : "memory");
  }
  // one plane of a group: a = the held first chunk (GROUP == 2 only), b = the current chunk; BUF = staging tile
  template <int BUF>
  __device__ __forceinline__ void store_group(const uint32_t (&a)[16], const uint32_t (&b)[16], int gcol) {
    // the store that last used this staging tile (NBUF stores ago) must have been read by the copy engine
    if (lane == 0) ptx::bulk_wait_group_read<NBUF - 1>();
    __syncwarp();
    if (GROUP == 2) {
      stage4<0>(a, BUF * TILE_BYTES);
      stage4<4>(b, BUF * TILE_BYTES);
    } else {
      stage4<0>(b, BUF * TILE_BYTES);
    }
    ptx::fence_proxy_async();                          // generic-proxy writes -> visible to the copy engine
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d_hint(cmap, stage_addr + BUF * TILE_BYTES, gcol, row0, keep);
      ptx::bulk_commit_group();
    }
  }
  template <int C, int NC>      // chunk C of the NC chunks this warp handles per tile
  __device__ __forceinline__ void chunk(const Params& p, const GemmShape& s, int row, int col0, float (&v)[32]) {
    // staging tiles are reused round-robin across tiles: the stores of one tile must fill whole rounds
    static_assert(PLANES == 0 || NBUF == 1 || ((PLANES * NC / GROUP) % NBUF == 0 && NC % GROUP == 0),
                  "stores per tile must be a multiple of the staging tiles");
    float e[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) e[j] = ptx::ex2_approx(v[j]);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      s0 += e[j];
      s1 += e[j + 1];
      s2 += e[j + 2];
      s3 += e[j + 3];
    }
    if constexpr (PLANES != 0) {
      uint32_t hi[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) hi[j] = ptx::pack_bf16x2(e[2 * j], e[2 * j + 1]);
      uint32_t lo[16];
      if constexpr (PLANES == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          lo[j] = ptx::pack_bf16x2(e[2 * j] - __uint_as_float(hi[j] << 16), e[2 * j + 1] - __uint_as_float(hi[j] & 0xffff0000u));
      }
      if constexpr (GROUP == 2 && (C & 1) == 0) {
        // first half of a 64-column group: keep it until the second half arrives
#pragma unroll
        for (int j = 0; j < 16; ++j) hold_hi[j] = hi[j];
        if constexpr (PLANES == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) hold_lo[j] = lo[j];
        }
      } else {
        constexpr int S = C / GROUP;                           // store index of this warp within the tile
        const int gcol = col0 - 32 * (GROUP - 1);
        store_group<(PLANES * S) % NBUF>(hold_hi, hi, gcol);
        if constexpr (PLANES == 2) store_group<(PLANES * S + 1) % NBUF>(hold_lo, lo, p.lo_col0 + gcol);
      }
    }
  }
  __device__ __forceinline__ void end_tile(const Params& p, const GemmShape& s, int row, int part, int) {
    if (row < s.M) p.rowsum_part[int64_t(part) * s.M + row] = (s0 + s1) + (s2 + s3);
  }
};

// ------------------------------------------------------------------ host side

// bf16 row-major [rows, cols] (leading dimension ld elements) -> 2-D tensor map with a
// [box_rows x 64] box and the 128-byte swizzle (operand loads and the epilogue's 32-row store tiles);
// box_cols = 32 selects the 64-byte swizzle.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols = UMMA_BK);

static inline void fill_segments(GemmShape& s, int planes, int K) {
  s.kb_per_seg = K / UMMA_BK;
  if (planes == 2) {
    s.num_seg = 3;   // hi*hi, hi*lo, lo*hi
    s.a_k0[0] = 0; s.b_k0[0] = 0;
    s.a_k0[1] = 0; s.b_k0[1] = K;
    s.a_k0[2] = K; s.b_k0[2] = 0;
  } else {
    s.num_seg = 1;
    s.a_k0[0] = s.b_k0[0] = 0;
    s.a_k0[1] = s.b_k0[1] = s.a_k0[2] = s.b_k0[2] = 0;
  }
}

// One problem of a grouped launch as the host describes it.  out = the epilogue's output tensor for
// epilogues that store through TMA (bf16 [M, out_cols], leading dimension out_ld), else unused.
template <class Epi>
struct GemmProblem {
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  int M, N, K, planes, splits;
  typename Epi::Params ep;
  void* out = nullptr; int64_t out_cols = 0; int64_t out_ld = 0;
  int kb_per_split = 0;      // > 0: k-block steps per split-K slice (the last slice may be shorter); overrides splits
};

// reserved_sms: SMs this launch leaves free (a bandwidth-bound kernel or a collective running beside the
// persistent grid needs somewhere to live); a per-call argument, the library keeps no scheduling state.
static inline int gemm_sm_budget(int reserved_sms) {
  const int n = sm_count() - (reserved_sms > 0 ? reserved_sms : 0);
  return n < 2 ? 2 : n;
}

// cost of a unit in k-block steps: its MMA steps plus the part of its epilogue / turn-around that the
// pipeline does not hide (~ 6 steps for a 256-wide tile)
constexpr int UMMA_UNIT_FIXED_COST = 6;

// Longest-processing-time-first placement of the units of a launch onto `workers` CTAs / CTA pairs.
template <class Epi>
static inline void build_schedule(GroupedArgs<Epi>& g, int workers) {
  UnitSchedule& sc = g.sched;
  const int n = g.tile_begin[g.num_problems];
  sc.num_units = n;
  sc.explicit_order = 0;
  if (n > UMMA_MAX_UNITS || workers > UMMA_MAX_WORKERS || n <= workers) return;
  std::vector<int> cost(n);
  bool uniform = true;
  for (int p = 0; p < g.num_problems; ++p) {
    const GemmShape& s = g.shape[p];
    const int tiles_mn = s.num_m_blk * s.num_n_blk;
    const int total_kb = s.num_seg * s.kb_per_seg;
    for (int t = g.tile_begin[p]; t < g.tile_begin[p + 1]; ++t) {
      const int split = (t - g.tile_begin[p]) / tiles_mn;
      const int kb0 = split * s.kb_per_split;
      const int kb1 = std::min(kb0 + s.kb_per_split, total_kb);
      cost[t] = (kb1 - kb0) + UMMA_UNIT_FIXED_COST;
      uniform = uniform && cost[t] == cost[0];
    }
  }
  if (uniform) return;                         // equal units: round-robin is already optimal
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  std::vector<long> load(workers, 0);
  std::vector<std::vector<uint16_t>> mine(workers);
  for (int i = 0; i < n; ++i) {
    int best = 0;
    for (int w = 1; w < workers; ++w)
      if (load[w] < load[best]) best = w;
    load[best] += cost[idx[i]];
    mine[best].push_back(uint16_t(idx[i]));
  }
  int pos = 0;
  for (int w = 0; w < workers; ++w) {
    sc.begin[w] = uint16_t(pos);
    for (uint16_t u : mine[w]) sc.order[pos++] = u;
  }
  for (int w = workers; w <= UMMA_MAX_WORKERS; ++w) sc.begin[w] = uint16_t(pos);
  sc.explicit_order = 1;
}

// makespan (in k-block steps incl. the fixed cost) of the same placement, for choosing split counts
static inline long lpt_makespan(std::vector<int> cost, int workers) {
  std::sort(cost.begin(), cost.end(), [](int a, int b) { return a > b; });
  std::vector<long> load(workers, 0);
  for (int c : cost) {
    int best = 0;
    for (int w = 1; w < workers; ++w)
      if (load[w] < load[best]) best = w;
    load[best] += c;
  }
  return *std::max_element(load.begin(), load.end());
}

template <class Epi>
static inline int fill_problems(const GemmProblem<Epi>* probs, int n, int m_tile, int BN, uint32_t b_box_rows,
                                TmapSet& tm, GroupedArgs<Epi>& g) {
  HMMC_REQUIRE(n >= 1 && n <= UMMA_MAX_PROBLEMS, "umma gemm: %d problems (max %d)", n, UMMA_MAX_PROBLEMS);
  g.num_problems = 0;
  g.tile_begin[0] = 0;
  for (int i = 0; i < n; ++i) {
    const GemmProblem<Epi>& pr = probs[i];
    HMMC_REQUIRE(pr.K % UMMA_BK == 0 && pr.K > 0, "umma gemm: K=%d must be a positive multiple of %d", pr.K, UMMA_BK);
    HMMC_REQUIRE(pr.planes == 1 || pr.planes == 2, "umma gemm: planes must be 1 or 2");
    HMMC_REQUIRE(pr.lda % 8 == 0 && pr.ldb % 8 == 0, "umma gemm: leading dimensions must be multiples of 8");
    HMMC_REQUIRE((reinterpret_cast<uintptr_t>(pr.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(pr.B) & 15) == 0,
                 "umma gemm: operands must be 16-byte aligned");
    if (pr.M <= 0 || pr.N <= 0) continue;
    const int k = g.num_problems;
    GemmShape& s = g.shape[k];
    s.M = pr.M;
    s.N = pr.N;
    s.num_m_blk = (pr.M + m_tile - 1) / m_tile;
    s.num_n_blk = (pr.N + BN - 1) / BN;
    fill_segments(s, pr.planes, pr.K);
    const int total_kb = s.num_seg * s.kb_per_seg;
    int splits = pr.splits < 1 ? 1 : (pr.splits > total_kb ? total_kb : pr.splits);
    s.kb_per_split = (total_kb + splits - 1) / splits;
    if (pr.kb_per_split > 0) s.kb_per_split = pr.kb_per_split < total_kb ? pr.kb_per_split : total_kb;
    s.num_splits = (total_kb + s.kb_per_split - 1) / s.kb_per_split;   // every split gets >= 1 k-block
    int rc = make_tmap_bf16(&tm.a[k], pr.A, uint64_t(pr.M), uint64_t(pr.planes) * pr.K, uint64_t(pr.lda), UMMA_BM);
    if (rc) return rc;
    rc = make_tmap_bf16(&tm.b[k], pr.B, uint64_t(pr.N), uint64_t(pr.planes) * pr.K, uint64_t(pr.ldb), b_box_rows);
    if (rc) return rc;
    if (Epi::USES_CMAP && pr.out != nullptr) {
      HMMC_REQUIRE(pr.N % 64 == 0 && pr.out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(pr.out) & 15) == 0,
                   "umma gemm: TMA-store epilogue needs N %% 64 == 0 and a 16-byte aligned output");
      rc = make_tmap_bf16(&tm.c[k], pr.out, uint64_t(pr.M), uint64_t(pr.out_cols), uint64_t(pr.out_ld), 32, Epi::STORE_COLS);
      if (rc) return rc;
    } else {
      tm.c[k] = tm.a[k];
    }
    g.ep[k] = pr.ep;
    g.tile_begin[k + 1] = g.tile_begin[k] + s.num_m_blk * s.num_n_blk * s.num_splits;
    g.num_problems = k + 1;
  }
  for (int k = g.num_problems; k < UMMA_MAX_PROBLEMS && g.num_problems > 0; ++k) {
    g.tile_begin[k + 1] = g.tile_begin[g.num_problems];
    tm.a[k] = tm.a[0];
    tm.b[k] = tm.b[0];
    tm.c[k] = tm.c[0];
    g.shape[k] = g.shape[0];
    g.ep[k] = g.ep[0];
  }
  return HMMC_OK;
}

template <class Kern, class Epi>
static inline int launch_gemm(Kern kern, int grid, int cluster, size_t smem, int threads, cudaStream_t stream,
                              const TmapSet& tm, const GroupedArgs<Epi>& g) {
  HMMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  count_launch();
  HMMC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, g));
  return HMMC_OK;
}

template <int BN, class Epi>
int launch_umma_grouped(const GemmProblem<Epi>* probs, int n, cudaStream_t stream, int reserved_sms = 0) {
  using Cfg = UmmaCfg<BN, Epi>;
  TmapSet tm;
  GroupedArgs<Epi> g;
  int rc = fill_problems(probs, n, UMMA_BM, BN, BN, tm, g);
  if (rc) return rc;
  if (g.num_problems == 0) return HMMC_OK;
  const int tiles = g.tile_begin[g.num_problems];
  const int budget = gemm_sm_budget(reserved_sms);
  const int grid = tiles < budget ? tiles : budget;
  build_schedule(g, grid);
  return launch_gemm(umma_gemm_kernel<BN, Epi>, grid, 1, Cfg::SMEM_BYTES, UMMA_THREADS, stream, tm, g);
}

// CTA-pair launch of a grouped GEMM (BN = 256).  GemmShape::num_m_blk counts 256-row super tiles;
// everything else (segments, split-K, epilogue parameters) is identical to launch_umma_grouped.
template <class Epi>
int launch_umma_grouped_pair(const GemmProblem<Epi>* probs, int n, cudaStream_t stream, int reserved_sms = 0) {
  constexpr int BN = UMMA_PAIR_BN;
  TmapSet tm;
  GroupedArgs<Epi> g;
  int rc = fill_problems(probs, n, 2 * UMMA_BM, BN, BN / 2, tm, g);
  if (rc) return rc;
  if (g.num_problems == 0) return HMMC_OK;
  const int tiles = g.tile_begin[g.num_problems];
  int pairs = gemm_sm_budget(reserved_sms) / 2;
  if (pairs > tiles) pairs = tiles;
  build_schedule(g, pairs);
  return launch_gemm(umma_gemm_pair_kernel<Epi>, 2 * pairs, 2, UmmaPairCfg<Epi>::SMEM_BYTES, UmmaPairCfg<Epi>::THREADS,
                     stream, tm, g);
}

// number of split-K partials launch_umma_grouped will produce for a request of `splits`
static inline int effective_splits(int K, int planes, int splits) {
  const int total_kb = ((planes == 2) ? 3 : 1) * (K / UMMA_BK);
  int sp = splits < 1 ? 1 : (splits > total_kb ? total_kb : splits);
  const int per = (total_kb + sp - 1) / sp;
  return (total_kb + per - 1) / per;
}

template <int BN, class Epi>
int launch_umma_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int planes,
                     int splits, const typename Epi::Params& ep, cudaStream_t stream) {
  GemmProblem<Epi> pr{A, lda, B, ldb, M, N, K, planes, splits, ep};
  return launch_umma_grouped<BN, Epi>(&pr, 1, stream);
}

// number of split-K slices that fills the machine for an (M x N) output with BN-wide tiles
static inline int pick_splits(int M, int N, int BN, int total_kb) {
  const int tiles = ((M + UMMA_BM - 1) / UMMA_BM) * ((N + BN - 1) / BN);
  int sp = sm_count() / (tiles > 0 ? tiles : 1);
  if (sp < 1) sp = 1;
  if (sp > 32) sp = 32;
  if (sp > total_kb) sp = total_kb;
  return sp;
}

}  // namespace hmmc

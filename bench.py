#!/usr/bin/env python
"""bench.py -- HM+MoCo contrastive head on B200.

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU)
    python bench.py --impl reference ...                      # the reference's CPU path (oracle port)
    python bench.py --impl reference-gpu ...                  # the same op sequence on the GPU (torch eager)

One "step" of the default workload is the pre-train head of BASELINE.json config 4 on one
rank: momentum EMA over the 172,325,632 key-encoder parameters, the FAM + VTM + FTM InfoNCE
losses against the negative queues forward AND backward (b = 128 samples per GPU, 12 frames,
dim 512, queue 1024, T 0.07), the key all-gather and the enqueue.  Ranks are data parallel
(weak scaling); the only exchange is the key all-gather.  A retrieval leg (config 2: 1000 x 1000
x 12 similarity + top-k frames + rank metrics) is timed after the main loop and reported in the
same JSON line under "retrieval".
"""
import argparse
import json
import os
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "gallery"])
    ap.add_argument("--texts", type=int, default=1000000, help="gallery workload: captions (10 per video)")
    ap.add_argument("--videos", type=int, default=100000, help="gallery workload: videos")
    ap.add_argument("--gallery-signal", default="1.0,0.7",
                    help="gallery workload: weight of a video's captions in its video / frame embeddings (lower = harder)")
    ap.add_argument("--precision", default=os.environ.get("HMMC_BENCH_PRECISION", "bf16"),
                    choices=["fp32", "bf16", "bf16x3"])
    ap.add_argument("--batch", type=int, default=128, help="samples per GPU")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--queue", type=int, default=1024)
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-overlap", action="store_true",
                    help="issue EMA and loss strictly one after the other (default: the query-side GEMMs of the loss "
                         "run beside the EMA, as the reference's forward order allows)")
    ap.add_argument("--no-graph", action="store_true", help="issue every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


FLOPS_ALGO = lambda b, F, D, K: 2 * 2.0 * D * b * (F * K * F + K * F + F * K + 2 * K)   # SURVEY.md §8d, fwd + bwd
EMA_ELEMS = 172325632


# DRAM bytes per launch of the dominant kernels, extracted from the committed `ncu --set full` captures by
# tools/ncu_traffic.py (dram__bytes_read.sum + dram__bytes_write.sum); None when a kernel is not in the file
NCU_TRAFFIC_FILE = "profiles/r2_ncu_traffic.json"


def ncu_traffic(*kernels):
    try:
        tab = json.load(open(os.path.join(ROOT, NCU_TRAFFIC_FILE)))["kernels"]
        tab = {k.split("<")[0]: v for k, v in tab.items()}          # template arguments dropped: bert_adam_kernel<0>
        return int(sum(tab[k]["dram_bytes_read"] + tab[k]["dram_bytes_write"] for k in kernels))
    except Exception:   # noqa: BLE001
        return None


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:   # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:   # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's torch port on the host cores
# ------------------------------------------------------------------------------------------
def cpu_pretrain_steps(args, max_seconds, min_steps=1, max_steps=10 ** 9, with_ema=True):
    """Times oracle/torch_port.pretrain_step (the reference's op sequence) on all host cores.
    Returns (ms per step, steps run, cores, sample description)."""
    from hmmc_b200 import synthetic as syn
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b, F, D, K = args.batch, args.frames, args.dim, args.queue
    inp_np = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    qs = {n: torch.from_numpy(x) for n, x in syn.queues(K, F=F, D=D, seed=3).items()}
    ema = None
    if with_ema:
        sizes = syn.ema_param_numels()
        big = torch.randn(sum(sizes))
        bigk = torch.randn(sum(sizes))
        ps = list(torch.split(big, sizes))
        pks = [torch.nn.Parameter(x.clone(), requires_grad=False) for x in torch.split(bigk, sizes)]
        ema = (ps, pks)
    ptr = 0
    times = []
    t_all = time.perf_counter()
    n = 0
    while True:
        inp = {k: torch.from_numpy(v).requires_grad_(k in ("v_fea", "title_fea", "frame_fea", "frame_pred"))
               for k, v in inp_np.items()}
        t0 = time.perf_counter()
        _, ptr = P.pretrain_step(inp, qs, ptr, K, 0.07, ema=ema)
        if ptr + b > K:
            ptr = 0
        times.append(time.perf_counter() - t0)
        n += 1
        if n >= max_steps or (n >= min_steps and time.perf_counter() - t_all > max_seconds):
            break
    use = times[1:] if len(times) > 2 else times          # first step pays allocator warm-up
    ms = 1e3 * float(np.median(use))
    sample = ("%d full steps of the same workload (b=%d, F=%d, D=%d, K=%d, EMA over %d params) on %d host threads, "
              "median" % (n, b, F, D, K, EMA_ELEMS if with_ema else 0, cores))
    return ms, n, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ms, n, cores, sample = cpu_pretrain_steps(args, max_seconds=120.0, min_steps=args.warmup + 1,
                                              max_steps=args.warmup + args.steps)
    value = args.batch / (ms / 1e3)
    line = {"impl": "reference", "metric": "hm_moco_head_fwd_bwd_throughput", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, max(n - 1, 0)), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is pure Python/PyTorch and absent on the GPU box: this arm times oracle/torch_port.py, "
                    "an op-for-op torch restatement pinned to golden vectors generated from the reference"}
    print(json.dumps(line))


def gpu_eager_pretrain_ms(args, dev, steps, warmup, with_ema=True):
    """The reference's op sequence (oracle/torch_port.pretrain_step) on CUDA tensors: what a user of the
    reference gets on this GPU today (torch eager: cuBLAS / ATen kernels).  Returns ms per step (CUDA events)."""
    from hmmc_b200 import synthetic as syn
    from oracle import torch_port as P
    b, F, D, K = args.batch, args.frames, args.dim, args.queue
    inp_np = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    qs = {n: torch.from_numpy(x).to(dev) for n, x in syn.queues(K, F=F, D=D, seed=3).items()}
    ema = None
    if with_ema:
        sizes = syn.ema_param_numels()
        ps = list(torch.split(torch.randn(sum(sizes), device=dev), sizes))
        pks = [torch.nn.Parameter(x.clone(), requires_grad=False) for x in torch.split(torch.randn(sum(sizes), device=dev), sizes)]
        ema = (ps, pks)
    base = {k: torch.from_numpy(v).to(dev) for k, v in inp_np.items()}
    ptr = 0

    def one():
        nonlocal ptr
        inp = {k: v.detach().requires_grad_(k in ("v_fea", "title_fea", "frame_fea", "frame_pred")) for k, v in base.items()}
        _, ptr = P.pretrain_step(inp, qs, ptr, K, 0.07, ema=ema)
        if ptr + b > K:
            ptr = 0
    for _ in range(max(warmup, 2)):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    del ema
    torch.cuda.empty_cache()
    return e0.elapsed_time(e1) / steps


def run_reference_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    steps = max(1, min(args.steps, 50))
    ms = gpu_eager_pretrain_ms(args, dev, steps, args.warmup)
    value = args.batch / (ms / 1e3)
    print(json.dumps({"impl": "reference-gpu", "metric": "hm_moco_head_fwd_bwd_throughput", "value": value,
                      "unit": "samples/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 2), "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": workload_config(args, 1),
                      "note": "the reference's own op sequence (oracle/torch_port.py, pinned to reference-generated "
                              "goldens) on CUDA tensors: torch eager, fp32, ~1100 launches for the EMA and 48 "
                              "contrastive_loss calls; the reference itself is absent on the GPU box"}))


def gpu_eager_legs(args, dev):
    """The other configs through the same port on the GPU (rank 0, N = 1): loss only at b=256, fine-tune head at
    B=256, config-2 eval.  ms each, CUDA events."""
    from hmmc_b200 import synthetic as syn
    from oracle import torch_port as P
    out = {}

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    a2 = argparse.Namespace(**vars(args))
    a2.batch = 256
    out["loss_b256_ms"] = gpu_eager_pretrain_ms(a2, dev, 10, 2, with_ema=False)
    t, v, fr = [torch.from_numpy(x).to(dev) for x in syn.finetune_inputs(256, seed=300)]

    def ft():
        P.finetune_step(t.detach().requires_grad_(True), v.detach().requires_grad_(True), fr.detach().requires_grad_(True))
    out["finetune_B256_ms"] = timed(ft, 20)
    T, V, Fr, gt, _ = syn.eval_inputs(1000, 1000, seed=4)
    T, V, Fr = [torch.from_numpy(x).to(dev) for x in (T, V, Fr)]
    out["eval_1k_ms"] = timed(lambda: P.eval_sim_and_rank(T, V, Fr, 2), 5)
    out["note"] = ("oracle/torch_port.py (the reference's op sequence) on CUDA tensors, torch eager fp32; eval includes the "
                   "reference's per-tile D2H copies and its numpy ranking on the host")
    return out


def workload_config(args, W):
    return {"workload": "pretrain head step (BASELINE config 4): EMA(172.3M params) + FAM/VTM/FTM InfoNCE fwd+bwd + "
                        "key all-gather + enqueue",
            "per_gpu_batch": args.batch, "global_batch": args.batch * W, "frames": args.frames, "dim": args.dim,
            "queue": args.queue, "temperature": 0.07, "momentum": 0.99, "parallelism": "dp%d" % W,
            "l2": "inputs > L2: every step streams the 2.07 GB EMA state through HBM (126 MB L2)"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
GRAPHS = []        # every GraphedStep of this process (released before the process group is destroyed)


def gallery_data(Nv, cap, D, F, lo, hi, dev, signal=(1.0, 0.7)):
    """Synthetic config-5 set: captions [Nv*cap, D] (replicated), videos/frames for [lo, hi).
    Generated in fixed blocks of 4096 videos so every world size sees the same bytes."""
    BLK = 4096
    T = torch.empty(Nv * cap, D, device=dev)
    V = torch.empty(hi - lo, D, device=dev)
    Fr = torch.empty(hi - lo, F, D, device=dev)
    for b0 in range(0, Nv, BLK):
        b1 = min(b0 + BLK, Nv)
        g = torch.Generator(device=dev).manual_seed(9000 + b0 // BLK)
        t = torch.randn((b1 - b0) * cap, D, device=dev, generator=g)
        T[b0 * cap:b1 * cap] = t
        a, c = max(b0, lo), min(b1, hi)
        v = torch.randn(b1 - b0, D, device=dev, generator=g)
        # caption quality differs from video to video (weight 0.15 .. 1 of the nominal signal): well-described
        # videos rank first, poorly described ones drown among the 1e5 distractors in BOTH directions
        u = (0.15 + 0.85 * torch.rand(b1 - b0, device=dev, generator=g))[:, None]
        if a < c:
            acc = t.view(b1 - b0, cap, D).sum(1) / cap ** 0.5
            fr = torch.randn(b1 - b0, F, D, device=dev, generator=g)
            # weak enough that the correct video / caption group is NOT always first among 1e5 distractors:
            # both rank directions carry non-trivial counts (t2v and the grouped v2t are really exercised)
            V[a - lo:c - lo] = (v + signal[0] * u * acc)[a - b0:c - b0]
            Fr[a - lo:c - lo] = (fr + signal[1] * (u * acc)[:, None, :])[a - b0:c - b0]
    return T, V, Fr


# A rank is an integer function of floating-point scores: two scores closer than the arithmetic's error may be
# ordered either way.  Tolerance on a score (scale 100 x [sim + mean of the top-k frame sims], unit vectors,
# D = 512): bf16 operands 2^-9 relative each -> a few 1e-4 on a dot product -> 0.08 on the score; bf16x3 (and the
# fp32 oracle itself) a few 1e-6 -> 2e-3.  The check: the fused rank must lie between the oracle's counts taken
# with thresholds gt + tol and gt - tol.
RANK_SCORE_TOL = {"bf16": 0.08, "bf16x3": 2e-3, "fp32": 2e-3}


def gallery_checks(T, V, Fr, cap, lo, hi, Nt, Nv, k, t2v, v2t, rank, dev, n_caps=128, n_vids=8, prec="bf16"):
    """Fused ranks against the numpy oracle (oracle/head_oracle.eval_scores, fp32, the checker -- nothing of it
    is timed): the t2v ranks of n_caps sampled captions over the WHOLE gallery (every rank scores its own shard
    on its host cores, counts are summed) and the grouped v2t ranks of n_vids sampled videos of rank 0's shard
    over ALL captions.  Mismatches are near-ties that the run's precision resolved the other way: every rank is
    also checked against the oracle's band for the precision's score tolerance (RANK_SCORE_TOL)."""
    from hmmc_b200 import parallel
    from oracle import head_oracle as O
    Tn = None
    idx = torch.arange(0, Nt, max(1, Nt // n_caps), device=dev)[:n_caps]
    Ts = T[idx].cpu().numpy()
    Vn, Fn = V.cpu().numpy(), Fr.cpu().numpy()
    ref = np.concatenate([O.eval_scores(Ts, Vn[j:j + 8192], Fn[j:j + 8192], k) for j in range(0, hi - lo, 8192)], axis=1)
    gt = (idx // cap).cpu().numpy()
    own = (gt >= lo) & (gt < hi)
    gts = np.zeros(idx.numel(), np.float32)
    gts[own] = ref[own, gt[own] - lo]
    gts_t = torch.from_numpy(gts).to(dev)
    parallel.all_reduce_sum_(gts_t)                       # each caption's score comes from exactly one shard
    g_all = gts_t.cpu().numpy()[:, None]
    tol = RANK_SCORE_TOL[prec]
    cnt = torch.from_numpy((ref > g_all).sum(1).astype(np.int32)).to(dev)
    cnt_lo = torch.from_numpy((ref > g_all + tol).sum(1).astype(np.int32)).to(dev)
    cnt_hi = torch.from_numpy((ref > g_all - tol).sum(1).astype(np.int32)).to(dev)
    for c in (cnt, cnt_lo, cnt_hi):
        parallel.all_reduce_sum_(c)
    got_t = t2v[idx]
    t2v_mism = int((cnt != got_t).sum())
    outside = int(((got_t < cnt_lo) | (got_t > cnt_hi)).sum())
    out = {"t2v_sampled_captions": int(idx.numel()), "t2v_mismatches_vs_fp32_oracle": t2v_mism,
           "t2v_max_abs_rank_diff": int((cnt - got_t).abs().max()), "score_tolerance": tol,
           "t2v_outside_tolerance_band": outside, "t2v_sample_rank_sum_oracle": int(cnt.long().sum())}
    if rank == 0:
        vids = np.linspace(0, hi - lo - 1, n_vids).astype(np.int64)
        Tn = T.cpu().numpy()
        best = np.full((Nv, n_vids), -np.inf, np.float32)          # M[g, j] = max over the captions of group g
        step_caps = cap * 8192                                      # whole caption groups per block
        for c0 in range(0, Nt, step_caps):
            blk = O.eval_scores(Tn[c0:c0 + step_caps], Vn[vids], Fn[vids], k)      # [captions, n_vids]
            g0 = c0 // cap
            best[g0:g0 + blk.shape[0] // cap] = blk.reshape(-1, cap, n_vids).max(1)
        own_g = best[lo + vids, np.arange(n_vids)]
        want = (best > own_g[None, :]).sum(0)
        got = v2t[torch.from_numpy(vids).to(dev)].cpu().numpy()
        out.update({"v2t_sampled_videos": int(n_vids), "v2t_mismatches_vs_fp32_oracle": int((want != got).sum()),
                    "v2t_sample_rank_sum_oracle": int(want.sum())})
    del Tn, Vn, Fn
    return out


def gallery_core(args, W, rank, local, dev, steps, warmup):
    """One timed config-5 run on an initialised process group; returns the result dict (all
    ranks must call it: the gallery is sharded over them)."""
    import torch.distributed as dist
    from hmmc_b200 import _lib, metrics as GM, modeling, ops, parallel, retrieval
    lib = _lib.load()
    Nv, D, F, k = args.videos, args.dim, args.frames, 3
    cap = max(1, args.texts // Nv)
    Nt = Nv * cap
    per = np.full(Nv, cap, dtype=np.int64)
    lo, hi = parallel.shard_range(Nv, W, rank)
    signal = tuple(float(x) for x in args.gallery_signal.split(","))
    T, V, Fr = gallery_data(Nv, cap, D, F, lo, hi, dev, signal)
    prec = args.precision if args.precision != "fp32" else "bf16x3"
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def one():
        return retrieval.fused_eval_ranks(T, V, Fr, per, 100.0, k, prec)

    for _ in range(max(1, warmup)):
        t2v, v2t = one()
    if W > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = lib.hmmc_launch_count()
    a, c = ev(), ev()
    a.record()
    for _ in range(steps):
        t2v, v2t = one()
    c.record()
    if W > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop()
    launches = lib.hmmc_launch_count() - l0
    ms = a.elapsed_time(c) / steps
    if W > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    checks = gallery_checks(T, V, Fr, cap, lo, hi, Nt, Nv, k, t2v, v2t, rank, dev, prec=prec)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops = 2.0 * Nt * Nv * D * (F + 1)
    tf = flops / W / (ms * 1e-3) / 1e12
    tv = GM.t2v_metrics_from_ranks(t2v.cpu().numpy())
    vt = GM.metrics_from_ranks(v2t.cpu().numpy())
    del T, V, Fr
    torch.cuda.empty_cache()
    return {"metric": "retrieval_sim_rank_throughput", "value": Nt / (ms / 1e3), "unit": "queries/s", "n_gpus": W,
            "steps": steps, "warmup": max(1, warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": "large-gallery retrieval (BASELINE config 5): fused sim + top-k frames + t2v/v2t ranks",
                       "texts": Nt, "videos": Nv, "frames": F, "dim": D, "top_frames": k, "captions_per_video": cap, "signal": list(signal),
                       "parallelism": "gallery sharded x%d" % W,
                       "l2": "inputs > L2: packed captions %.2f GB + gallery shard %.2f GB" %
                             (Nt * D * 2e-9 * (2 if prec == "bf16x3" else 1),
                              (hi - lo) * 13 * D * 2e-9 * (2 if prec == "bf16x3" else 1))},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": Nt / (ms / 1e3), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "embeddings are produced on the device by the encoders; ranks stay on the device"},
            "roofline": {"kernel": "eval_rank_kernel", "bound": "tensor", "achieved": tf, "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": tf / tf_peak, "traffic": None,
                         "algorithmic_flops_per_gpu": flops / W,
                         "note": "per GPU; includes packing, ground-truth pass and collectives (whole pass timed)"},
            "checks": dict(checks, t2v_rank_sum=int(t2v.long().sum()), v2t_rank_sum=int(v2t.long().sum())),
            "metrics": {"t2v_R1": tv["R1"], "t2v_MeanR": tv["MeanR"], "v2t_R1": vt["R1"], "v2t_MeanR": vt["MeanR"]}}


def run_gallery(args):
    """BASELINE config 5: Nt captions x Nv videos x 12 frames, gallery sharded over the ranks,
    fused similarity + top-k frames + t2v / v2t rank counting (no matrix)."""
    import torch.distributed as dist
    from hmmc_b200 import ops
    W = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if W > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.device_check()
    line = gallery_core(args, W, rank, local, dev, max(1, args.steps), max(1, min(args.warmup, 2)))
    if rank == 0:
        print(json.dumps(line))
    if W > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-gpu":
        return run_reference_gpu(args)
    if args.workload == "gallery":
        return run_gallery(args)

    import torch.distributed as dist
    from hmmc_b200 import _lib, modeling, ops, retrieval
    from hmmc_b200 import synthetic as syn

    W = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if W > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.device_check()
    lib = _lib.load()

    b, F, D, K = args.batch, args.frames, args.dim, args.queue
    if (b * W) > K or K % (b * W):
        # the reference requires ptr + B <= K (no wrap inside a batch): grow the queue with the world
        K = max(K, b * W)
    task = types.SimpleNamespace(local_rank=local, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                                 contrast_num_negative=K, max_frames=F, use_frame_fea=True,
                                 head_precision=args.precision)

    class Params(torch.nn.Module):
        def __init__(self, flat, sizes):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(x, requires_grad=False) for x in torch.split(flat, sizes)])

    sizes = syn.ema_param_numels()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    enc = Params(torch.randn(sum(sizes), device=dev, generator=g), sizes)
    enc_k = Params(torch.randn(sum(sizes), device=dev, generator=g), sizes)
    model = modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=D), task).to(dev)
    model.model_pairs = [[enc, enc_k]]
    with torch.no_grad():
        for n, x in syn.queues(K, F=F, D=D, seed=3).items():
            getattr(model, n).copy_(torch.from_numpy(x))

    inp_np = syn.pretrain_inputs(b, F=F, D=D, seed=100 + rank)
    q_names = ("v_fea", "title_fea", "frame_fea", "frame_pred")
    order = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k",
             "frame_proj_k"]
    host = {n: torch.from_numpy(inp_np[n]).pin_memory() for n in order}
    sizes_in = [host[n].numel() for n in order]
    packed_host = torch.cat([host[n].reshape(-1) for n in order]).pin_memory()
    static_in = packed_host.to(dev)                       # the step's inputs live here (one H2D copy per step in e2e)

    def views(buf):
        out, off = {}, 0
        for n, k in zip(order, sizes_in):
            out[n] = buf[off:off + k].view(host[n].shape).detach().requires_grad_(n in q_names)
            off += k
        return out

    devt = views(static_in)
    h2d_bytes = packed_host.numel() * 4

    ev = lambda: torch.cuda.Event(enable_timing=True)
    ema_events, head_events = [], []
    split = modeling.use_split_schedule() and not args.no_overlap

    def step(inputs, timed, sequential=False):
        for n in q_names:
            inputs[n].grad = None
        if timed:
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
        # deferred schedule (several ranks): the previous step's key all-gather + enqueue run beside the EMA
        model.start_pending_exchange()
        if split and not timed and not sequential:
            # the reference's own order (modules/modeling.py:340-377): queries first, then the momentum update and
            # the key encoders, then the losses -> the query-side GEMMs of the loss run beside the EMA
            begun = model.head_loss_begin(*[inputs[n] for n in order[:4]])
            with torch.no_grad():
                model._momentum_update()
            loss = model.head_loss_end(begun, *[inputs[n] for n in order[4:]])
            loss.backward()
            return loss
        with torch.no_grad():
            model._momentum_update()
        if timed:
            e1.record()
        loss = model.head_loss(*[inputs[n] for n in order])
        loss.backward()
        if timed:
            e2.record()
            ema_events.append((e0, e1))
            head_events.append((e1, e2))
        return loss

    def barrier():
        if W > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(devt, False)
    barrier()
    l0 = lib.hmmc_launch_count()
    step(devt, False)
    launches_per_step = int(lib.hmmc_launch_count() - l0)      # kernels of this library in one step (counted, eager)
    barrier()

    # One step = a fixed sequence of launches: capture it once, replay it (hmmc_b200/graphs.py).
    graphed = None
    graph_note = "eager (--no-graph)"
    dbg = (lambda m: print("[bench r%d] %s" % (rank, m), file=sys.stderr, flush=True)) if os.environ.get("HMMC_BENCH_DEBUG") else (lambda m: None)
    want_graph = not args.no_graph
    if want_graph:
        try:
            from hmmc_b200.graphs import GraphedStep
            dbg("capturing")
            graphed = GraphedStep(lambda: step(devt, False))
            GRAPHS.append(graphed)
            dbg("captured")
            graph_note = "cuda graph replay"
        except Exception as e:   # noqa: BLE001
            graphed = None
            graph_note = "eager (graph capture failed: %s)" % repr(e)[:200]
            torch.cuda.synchronize()
    run_step = (lambda: graphed.replay()) if graphed is not None else (lambda: step(devt, False))
    for _ in range(3):
        run_step()
    dbg("replayed 3")
    barrier()
    dbg("barrier ok")

    # ---- timed region 1: inputs resident in HBM
    clocks = ClockSampler(local)
    launches0 = lib.hmmc_launch_count()
    clocks.start()
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    cpu_t0 = time.perf_counter()
    for _ in range(args.steps):
        run_step()
    cpu_issue_ms = 1e3 * (time.perf_counter() - cpu_t0) / args.steps     # host time to ISSUE one step (no sync)
    t1.record()
    # the timed region is exactly args.steps steps; the same step keeps running (untimed) until the clock
    # sampler has covered at least one second of it, so that the record has enough samples
    t_clock = time.perf_counter()
    barrier()
    extra = 0
    while time.perf_counter() - cpu_t0 < 1.0:
        for _ in range(50):
            run_step()
        extra += 50
        torch.cuda.synchronize()
    clk = clocks.stop()
    clk["window"] = "timed region + %d more untimed steps of the same workload (%.2f s in all)" % (
        extra, time.perf_counter() - cpu_t0)
    barrier()
    launches = lib.hmmc_launch_count() - launches0
    if graphed is not None:
        launches = launches_per_step * args.steps         # a replay re-runs the captured launches
    dbg("timed region 1 done")
    # per-kernel timing for the roofline: a short eager run right after, events around the EMA launch
    for _ in range(20):
        step(devt, True)
    barrier()
    dbg("eager breakdown done")
    ms_total = t0.elapsed_time(t1)
    detail = None
    ms_ema = float(np.mean([a.elapsed_time(c) for a, c in ema_events]))
    ms_head = float(np.mean([a.elapsed_time(c) for a, c in head_events]))
    # the step with EMA and loss strictly one after the other (replayed as a graph too): what the head costs when
    # nothing hides it, and how much the concurrent schedule saves
    # the EMA kernel alone, replayed back to back (2 GB per launch: nothing stays in L2): its duration without the
    # host gaps an eager launch carries
    if graphed is not None:
        try:
            def ema_only():
                with torch.no_grad():
                    model._momentum_update()
            gema = GraphedStep(ema_only)
            GRAPHS.append(gema)
            for _ in range(3):
                gema.replay()
            barrier()
            q0, q1 = ev(), ev()
            q0.record()
            for _ in range(30):
                gema.replay()
            q1.record()
            barrier()
            ms_ema = q0.elapsed_time(q1) / 30
        except Exception:   # noqa: BLE001
            torch.cuda.synchronize()
    ms_seq = None
    if graphed is not None and split:
        try:
            gseq = GraphedStep(lambda: step(devt, False, True))
            GRAPHS.append(gseq)
            for _ in range(3):
                gseq.replay()
            barrier()
            q0, q1 = ev(), ev()
            q0.record()
            for _ in range(50):
                gseq.replay()
            q1.record()
            barrier()
            ms_seq = q0.elapsed_time(q1) / 50
        except Exception:   # noqa: BLE001
            ms_seq = None
            torch.cuda.synchronize()

    # ---- timed region 2: end to end through the public API with host buffers.
    # Every step: ONE host->device copy of the step's packed inputs from pinned memory (on a copy
    # stream, so step i+1's upload overlaps step i's kernels, like a prefetching data loader), the
    # step itself (BirdPreTrainedModel._momentum_update + head_loss + backward, replayed as a graph
    # when available), and a device->host copy of the loss into pinned memory.  The timed region
    # ends with a full synchronise.
    e2e_steps = args.steps
    loss_host = torch.empty(e2e_steps, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    dbuf = [torch.empty_like(static_in) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])          # the step that last read this buffer is done
            dbuf[i % 2].copy_(packed_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    for e in freed:
        e.record()
    barrier()
    s0, s1 = ev(), ev()
    s0.record()
    upload(0)
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            upload(i + 1)
        torch.cuda.current_stream().wait_event(ready[i % 2])
        static_in.copy_(dbuf[i % 2], non_blocking=True)   # into the step's static inputs
        freed[i % 2].record()
        loss = run_step()
        loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
    s1.record()
    barrier()
    ms_e2e_total = s0.elapsed_time(s1)
    assert bool(torch.isfinite(loss_host).all())

    def allmax(x):
        if W == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_total, ms_e2e_total = allmax(ms_total), allmax(ms_e2e_total)
    ms_step = ms_total / args.steps
    value = W * b / (ms_step / 1e3)
    e2e_value = W * b / (ms_e2e_total / e2e_steps / 1e3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    ema_bytes = 12.0 * EMA_ELEMS
    ema_gbs = ema_bytes / (ms_ema * 1e-3) / 1e9
    head_flops = FLOPS_ALGO(b, F, D, K)
    if graphed is not None:
        # replayed (sequential) step minus the EMA kernel (events cannot sit inside a replay)
        ms_head = max((ms_seq if ms_seq is not None else ms_step) - ms_ema, 1e-6)
    head_tf = head_flops / (ms_head * 1e-3) / 1e12

    line = {"metric": "hm_moco_head_fwd_bwd_throughput", "value": value, "unit": "samples/s", "n_gpus": W,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (fp32-parity split)",
                                                                "fp32": "f32"}[args.precision],
            "data": "synthetic", "config": workload_config(args, W),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e_total / e2e_steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "note": "packed inputs uploaded on a copy stream (overlaps the previous step); loss copied to pinned memory every step"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "ema_multi_kernel", "bound": "hbm", "achieved": ema_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": ema_gbs / hbm_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
                         "traffic": ncu_traffic("ema_multi_kernel"), "traffic_source": NCU_TRAFFIC_FILE,
                         "algorithmic_bytes": ema_bytes, "ms": ms_ema, "peak_source": peak_src},
            "roofline_head": {"kernels": "infonce fwd+bwd (5 query blocks: prep_rows, umma S-GEMM+exp epilogue, "
                                         "umma U-GEMM, finish, gradient scale) + enqueue",
                              "bound": "tensor", "achieved": head_tf, "peak": tf_peak, "unit": "TFLOP/s",
                              "frac": head_tf / tf_peak, "algorithmic_flops": head_flops, "ms": ms_head,
                              "peak_source": peak_src},
            "breakdown_ms": {"ema": ms_ema, "head_fwd_bwd_enqueue": ms_head, "head_detail": detail,
                             "sequential_step": ms_seq,
                             "schedule": ("query-side GEMMs of the loss beside the EMA (head_loss_begin / head_loss_end)"
                                          if split else "EMA, then loss, then enqueue"),
                             "host_issue_per_step": cpu_issue_ms,
                             "note": "ema: the kernel alone, replayed back to back from a graph; head: the sequential step (graph replay) minus ema; eager fallbacks use CUDA events around the calls"},
            "issue_mode": graph_note}

    # ---- fine-tune head leg (config 3), all ranks
    if not args.no_retrieval:
        try:
            ft = finetune_leg(args, W, rank, local, dev)
            if rank == 0:
                line["finetune_head"] = ft
        except Exception as e:   # noqa: BLE001
            line["finetune_head"] = {"error": repr(e)[:300]}
    # ---- north_star target shape: loss forward+backward alone at batch 256, rank 0
    if rank == 0 and not args.no_retrieval:
        try:
            line["loss_b256"] = loss_target_leg(args, dev, tf_peak, peak_src)
        except Exception as e:   # noqa: BLE001
            line["loss_b256"] = {"error": repr(e)[:300]}
    # ---- projector / predictor MLP right before the head, rank 0
    if rank == 0 and not args.no_retrieval:
        try:
            line["mlp"] = mlp_leg(args, dev, tf_peak, peak_src)
        except Exception as e:   # noqa: BLE001
            line["mlp"] = {"error": repr(e)[:300]}
    # ---- optimizer step that follows the head's backward, rank 0
    if rank == 0 and not args.no_retrieval:
        try:
            line["optimizer_step"] = optimizer_leg(args, dev, sizes, hbm_peak, peak_src)
        except Exception as e:   # noqa: BLE001
            line["optimizer_step"] = {"error": repr(e)[:300]}
    # ---- retrieval legs: config 2 on rank 0; config 5 (gallery sharded over all ranks)
    if rank == 0 and not args.no_retrieval:
        line["retrieval"] = retrieval_leg(args, dev)
    if not args.no_retrieval:
        del enc, enc_k, model
        torch.cuda.empty_cache()
        try:
            big = gallery_core(args, W, rank, local, dev, 2, 1)
            line["retrieval_large"] = {k: big[k] for k in ("value", "unit", "ms_per_step", "dtype", "config", "roofline",
                                                           "checks", "metrics", "clocks")}
            # the fp32-parity mode (bit-exact ranks on a tie-free matrix): same run, three tensor-core products
            a3 = argparse.Namespace(**vars(args))
            a3.precision = "bf16x3"
            big3 = gallery_core(a3, W, rank, local, dev, 1, 1)
            line["retrieval_large"]["bf16x3"] = {k: big3[k] for k in ("value", "unit", "ms_per_step", "dtype", "checks",
                                                                       "metrics")}
            line["retrieval_large"]["bf16x3"]["roofline_frac_algorithmic"] = big3["roofline"]["frac"]
        except Exception as e:   # noqa: BLE001
            line["retrieval_large"] = {"error": repr(e)[:300]}
    if W > 1:
        # multi-rank parity (tests/multi_gpu_check.py; the numpy oracle is the checker): gathered enqueue ==
        # oracle concat (eager / deferred / graph replay), fine-tune loss and gradients == W = 1 on the
        # concatenated batch, replicated backward == reduce-scatter, sharded eval == one GPU
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import multi_gpu_check as MG
            chk = MG.run_checks(W, rank, local, dev, which=MG.CHECKS[:3])
        except Exception as e:   # noqa: BLE001
            chk = {"pass": False, "failures": [repr(e)[:300]]}
        line["checks"] = chk
    if rank == 0 and W == 1 and not args.no_cpu_baseline:
        # rank 0 at N = 1 only: under torchrun the other ranks would spin in a barrier on the same cores
        ms, n, cores, sample = cpu_pretrain_steps(args, max_seconds=args.cpu_seconds)
        line["cpu_baseline"] = {"value": b / (ms / 1e3), "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": sample, "ms_per_step": ms}
    if rank == 0 and W == 1 and not args.no_retrieval:
        # what a user of the reference gets on this GPU today: its op sequence in torch eager
        try:
            gms = gpu_eager_pretrain_ms(args, dev, 10, 2)
            eager = {"pretrain_step_ms": gms, "value": b / (gms / 1e3), "unit": "samples/s", "kind": "port on cuda"}
            eager.update(gpu_eager_legs(args, dev))
            eager["ours_over_eager"] = {
                "pretrain_step": gms / ms_step,
                "loss_b256": eager["loss_b256_ms"] / line["loss_b256"]["bf16"]["ms"] if "bf16" in line.get("loss_b256", {}) else None,
                "finetune_B256": eager["finetune_B256_ms"] / line["finetune_head"]["b256_per_gpu"]["ms_per_step"]
                if "b256_per_gpu" in line.get("finetune_head", {}) else None,
                "eval_1k": eager["eval_1k_ms"] / line["retrieval"]["ms"] if "ms" in line.get("retrieval", {}) else None}
            line["gpu_eager_baseline"] = eager
        except Exception as e:   # noqa: BLE001
            line["gpu_eager_baseline"] = {"error": repr(e)[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if W > 1:
        dist.barrier()
        torch.cuda.synchronize()
        # graphs that captured NCCL work must be destroyed before the communicator
        for gobj in GRAPHS:
            try:
                gobj.release()
            except Exception:   # noqa: BLE001
                pass
        del GRAPHS[:]
        sys.stdout.flush()
        sys.stderr.flush()
        dist.destroy_process_group()


def loss_target_leg(args, dev, tf_peak, peak_src):
    """BASELINE north_star target: the fused HM + MoCo loss forward/backward at batch 256, 12 frames,
    dim 512, queue 1024 (no EMA, no enqueue), CUDA-graph replay, in the bf16 mode and in the
    fp32-parity bf16x3 mode (which executes 3x the algorithmic FLOPs on the tensor cores)."""
    from hmmc_b200 import ops
    from hmmc_b200 import synthetic as syn
    from hmmc_b200.graphs import GraphedStep
    b, F, D, K = 256, 12, 512, 1024
    inp = syn.pretrain_inputs(b, F=F, D=D, seed=2)
    qs = {n: torch.from_numpy(x).to(dev) for n, x in syn.queues(K, F=F, D=D, seed=3).items()}
    t = {n: torch.from_numpy(x).to(dev) for n, x in inp.items()}
    qn = ("v_fea", "title_fea", "frame_fea", "frame_pred")
    for n in qn:
        t[n].requires_grad_(True)
    algo = FLOPS_ALGO(b, F, D, K)
    out = {"workload": "pre-train head loss fwd+bwd only, b=256 F=12 D=512 K=1024", "algorithmic_flops": algo,
           "peak": tf_peak, "peak_source": peak_src, "unit": "TFLOP/s"}
    for prec in ("bf16", "bf16x3"):
        def run(prec=prec):
            for n in qn:
                t[n].grad = None
            total, _ = ops.pretrain_head(t["v_fea"], t["title_fea"], t["frame_fea"], t["frame_pred"], t["v_fea_k"],
                                         t["title_fea_k"], t["frame_fea_k"], t["frame_proj_k"], qs["queue_v_cross_ng"],
                                         qs["queue_title_cross_ng"], qs["queue_frame_proj_ng"],
                                         qs["queue_frame_cross_ng"], 0.07, 0.05, 0.45, 0.45, True, prec)
            total.backward()
            return total
        for _ in range(3):
            run()
        try:
            g = GraphedStep(run)
            fn, mode = g.replay, "cuda graph replay"
        except Exception as e:   # noqa: BLE001
            fn, mode = run, "eager (%s)" % repr(e)[:80]
            torch.cuda.synchronize()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        executed = algo * (3 if prec == "bf16x3" else 1)
        out[prec] = {"ms": ms, "samples_per_s": b / (ms / 1e3), "achieved": algo / (ms * 1e-3) / 1e12,
                     "frac": algo / (ms * 1e-3) / 1e12 / tf_peak, "executed_tflops": executed / (ms * 1e-3) / 1e12,
                     "executed_frac": executed / (ms * 1e-3) / 1e12 / tf_peak, "issue_mode": mode}
    return out


def mlp_leg(args, dev, tf_peak, peak_src):
    """One projector MLP (modules/modeling.py:788-807: 512 -> 4096 -> BatchNorm -> ReLU -> 512) forward +
    backward on the b*F = 1536 frame rows of the pre-train step, fp32-parity (bf16x3) and bf16 modes."""
    from hmmc_b200.graphs import GraphedStep
    from hmmc_b200.mlp import MLP
    M, Din, Dh, Dout = args.batch * args.frames, args.dim, 4096, args.dim
    algo = 6 * 2.0 * M * Din * Dh        # two forward and four backward GEMMs of equal size
    out = {"workload": "MLP fwd+bwd, %d rows, %d -> %d -> %d, training-mode BatchNorm" % (M, Din, Dh, Dout),
           "algorithmic_flops": algo, "peak": tf_peak, "unit": "TFLOP/s", "peak_source": peak_src}
    x = torch.randn(M, Din, device=dev, requires_grad=True)
    dy = torch.randn(M, Dout, device=dev) * 0.05
    for prec in ("bf16", "bf16x3"):
        m = MLP(Din, Dh, Dout, 2, precision=prec).to(dev).train()

        def step():
            x.grad = None
            for p in m.parameters():
                p.grad = None
            y = m(x)
            y.backward(dy)
            return y
        for _ in range(3):
            step()
        try:
            g = GraphedStep(step)
            fn, mode = g.replay, "cuda graph replay"
        except Exception as e:   # noqa: BLE001
            fn, mode = step, "eager (%s)" % repr(e)[:80]
            torch.cuda.synchronize()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 100
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ex = algo * (3 if prec == "bf16x3" else 1)
        out[prec] = {"ms": ms, "rows_per_s": M / (ms / 1e3), "achieved": algo / (ms * 1e-3) / 1e12,
                     "frac": algo / (ms * 1e-3) / 1e12 / tf_peak, "executed_frac": ex / (ms * 1e-3) / 1e12 / tf_peak,
                     "issue_mode": mode}
    if not args.no_cpu_baseline:
        # the reference class restated with torch modules, on the host cores
        ref = torch.nn.Sequential(torch.nn.Linear(Din, Dh), torch.nn.BatchNorm1d(Dh), torch.nn.ReLU(inplace=True),
                                  torch.nn.Linear(Dh, Dout)).train()
        xc = torch.randn(M, Din, requires_grad=True)
        dyc = torch.randn(M, Dout) * 0.05
        ref(xc).backward(dyc)
        t0 = time.perf_counter()
        k = 0
        while k < 3 or (time.perf_counter() - t0 < 3.0 and k < 50):
            ref(xc).backward(dyc)
            k += 1
        cms = (time.perf_counter() - t0) * 1e3 / k
        out["cpu_baseline"] = {"value": M / (cms / 1e3), "unit": "rows/s", "ms": cms, "cores": torch.get_num_threads(),
                               "kind": "port", "sample": "%d full fwd+bwd passes of the same shape" % k}
    return out


def optimizer_leg(args, dev, sizes, hbm_peak, peak_src):
    """clip_grad_norm_(all, 1.0) + BertAdam.step over the 362 tensors / 172 M parameters the EMA walks
    (main_pretrain.py:277-286), as prep_optimizer configures it (warmup_cosine, b2 0.98, two lr groups,
    decay / no decay).  GPU: three launches per step.  CPU: the reference's per-parameter op sequence
    (oracle/torch_port.py) on the host cores."""
    from hmmc_b200 import _lib
    from hmmc_b200.optimization import BertAdam
    n = sum(sizes)
    flat = torch.randn(n, device=dev) * 0.02
    gflat = torch.randn(n, device=dev) * 1e-4
    params = [torch.nn.Parameter(x) for x in torch.split(flat, sizes)]
    grads = list(torch.split(gflat, sizes))
    common = dict(schedule='warmup_cosine', warmup=0.1, t_total=10000, b1=0.9, b2=0.98, e=1e-6, max_grad_norm=1.0)
    decay = [p for p in params if p.numel() >= 4096]
    nodecay = [p for p in params if p.numel() < 4096]
    opt = BertAdam([dict(common, params=decay, lr=1e-7, weight_decay=0.2),
                    dict(common, params=nodecay, lr=1e-7, weight_decay=0.0)], lr=1e-7)

    def one():
        for p, g in zip(params, grads):
            p.grad = g
        opt.step(global_max_norm=1.0)
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    lib = _lib.load()
    reps = 20
    l0 = lib.hmmc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        one()
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / reps
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    launches = (lib.hmmc_launch_count() - l0) // reps
    algo = 32.0 * n          # norm pass reads g (4 B); update pass reads p, g, m, v and writes p, m, v (28 B)
    out = {"workload": "clip_grad_norm_(1.0) + BertAdam.step, %d tensors, %d fp32 parameters" % (len(sizes), n),
           "ms_per_step": ms, "host_issue_ms_per_step": host_ms, "launches_per_step": int(launches),
           "params_per_s": n / (ms / 1e3), "grad_norm": float(opt.last_grad_norm),
           "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": algo / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": algo,
                        # dram__bytes_read.sum + dram__bytes_write.sum of grad_sqnorm_kernel and bert_adam_kernel
                        "traffic": ncu_traffic("grad_sqnorm_kernel", "bert_adam_kernel"), "traffic_source": NCU_TRAFFIC_FILE,
                        "peak_source": peak_src}}
    assert bool(torch.isfinite(flat).all())
    del opt, params, grads, flat, gflat
    torch.cuda.empty_cache()
    if not args.no_cpu_baseline:
        from oracle import torch_port as P
        frac = 8                                   # every 8th tensor: ~1/8 of the parameters, same size mix
        csz = sizes[::frac]
        cp = [torch.randn(k) * 0.02 for k in csz]
        cg = [torch.randn(k) * 1e-4 for k in csz]
        st = P.bert_adam_state(cp)
        groups = [dict(common, lr=1e-7, weight_decay=0.2), dict(common, lr=1e-7, weight_decay=0.0)]
        gof = [0 if k >= 4096 else 1 for k in csz]
        P.clip_and_bert_adam_step(cp, cg, st, groups, gof, 1.0)
        t0 = time.perf_counter()
        k = 0
        while k < 3 or (time.perf_counter() - t0 < 5.0 and k < 20):
            P.clip_and_bert_adam_step(cp, cg, st, groups, gof, 1.0)
            k += 1
        cms = (time.perf_counter() - t0) * 1e3 / k
        out["cpu_baseline"] = {"value": sum(csz) / (cms / 1e3), "unit": "params/s", "cores": torch.get_num_threads(),
                               "kind": "port", "ms_per_sample_step": cms,
                               "sample": "every %dth of the %d tensors (%d parameters), %d steps" % (frac, len(sizes), sum(csz), k)}
    return out


def finetune_leg(args, W, rank, local, dev):
    """BASELINE config 3: fine-tune head, 32 samples per GPU (global batch 32*W), 12 frames: packed
    all-gather + VTM/FTM symmetric CrossEn forward and backward.  All ranks take part."""
    import torch.distributed as dist
    from hmmc_b200 import modeling
    from hmmc_b200 import synthetic as syn
    from hmmc_b200.graphs import GraphedStep
    out = {}
    for b in ([32] if W > 1 else [32, 256]):
        t, v, fr = syn.finetune_inputs(b, seed=300 + rank)
        task = types.SimpleNamespace(local_rank=local, top_frames=2, use_frame_fea=True, head_precision="bf16x3")
        m = modeling.BirdModel(modeling.default_cross_config(), task)
        ins = [torch.from_numpy(x).to(dev).requires_grad_(True) for x in (t, v, fr)]

        def step():
            for x in ins:
                x.grad = None
            loss = m.head_loss(*ins)
            loss.backward()
            return loss
        try:
            g = GraphedStep(step)
            GRAPHS.append(g)
            run, mode = g.replay, "cuda graph replay"
        except Exception as e:   # noqa: BLE001
            run, mode = step, "eager (%s)" % repr(e)[:80]
            torch.cuda.synchronize()
        for _ in range(5):
            run()
        if W > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 100
        e0.record()
        for _ in range(reps):
            loss = run()
        e1.record()
        if W > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if W > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        out["b%d_per_gpu" % b] = {"global_batch": b * W, "ms_per_step": ms, "samples_per_s": b * W / (ms / 1e3),
                                  "loss": float(loss.detach()), "issue_mode": mode}
    out["workload"] = "BASELINE config 3: fine-tune head fwd+bwd (packed all-gather + 13 symmetric CrossEn matrices), bf16x3"
    return out


def retrieval_leg(args, dev):
    """config 2: 1000 texts x 1000 videos x 12 frames, top_frames 2: similarity + both rank directions."""
    from hmmc_b200 import modeling, ops, retrieval
    from hmmc_b200 import synthetic as syn
    T, V, Fr, gt, _ = syn.eval_inputs(1000, 1000, seed=4)
    task = types.SimpleNamespace(local_rank=0, top_frames=2, use_frame_fea=True, head_precision=args.precision)
    m = modeling.BirdModel(modeling.default_cross_config(), task)
    hT, hV, hF = [torch.from_numpy(x).pin_memory() for x in (T, V, Fr)]
    dT, dV, dF = hT.to(dev), hV.to(dev), hF.to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def dev_pass():
        sim = retrieval.similarity_matrix(m, dT, dV, dF)
        return ops.rank_count(sim)

    for _ in range(3):
        dev_pass()
    torch.cuda.synchronize()
    # device time: the pass is a fixed sequence of four launches, replayed from a CUDA graph like the other legs
    run, mode = dev_pass, "eager"
    if not args.no_graph:
        try:
            from hmmc_b200.graphs import GraphedStep
            g = GraphedStep(lambda: dev_pass()[0])
            GRAPHS.append(g)
            run, mode = g.replay, "cuda graph replay"
        except Exception:   # noqa: BLE001
            torch.cuda.synchronize()
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    reps = 100
    a, c = ev(), ev()
    a.record()
    for _ in range(reps):
        run()
    c.record()
    torch.cuda.synchronize()
    ms_dev = a.elapsed_time(c) / reps
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        sim = retrieval.similarity_matrix(m, hT.to(dev, non_blocking=True), hV.to(dev, non_blocking=True),
                                          hF.to(dev, non_blocking=True))
        t2v, v2t = ops.rank_count(sim)
        r = (t2v.cpu(), v2t.cpu())
    torch.cuda.synchronize()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / reps
    from hmmc_b200 import metrics as GM
    tv = GM.metrics_from_ranks(r[0].numpy())
    out = {"workload": "BASELINE config 2: 1000 x 1000 x 12, top_frames 2, sim + top-k + t2v/v2t ranks",
           "value": 1000.0 / (ms_dev / 1e3), "unit": "queries/s", "ms": ms_dev, "issue_mode": mode,
           "e2e": {"value": 1000.0 / (ms_e2e / 1e3), "unit": "queries/s", "ms": ms_e2e,
                   "h2d_bytes": int(4 * (T.size + V.size + Fr.size)), "d2h_bytes": 8000},
           "R1": tv["R1"], "MeanR": tv["MeanR"]}
    if not args.no_cpu_baseline:
        from oracle import torch_port as P
        torch.set_num_threads(os.cpu_count() or 1)
        cT, cV, cF = torch.from_numpy(T), torch.from_numpy(V), torch.from_numpy(Fr)
        P.eval_sim_and_rank(cT, cV, cF, 2)
        t0 = time.perf_counter()
        n = 0
        while n < 3 or time.perf_counter() - t0 < 3.0:
            P.eval_sim_and_rank(cT, cV, cF, 2)
            n += 1
        ms_cpu = 1e3 * (time.perf_counter() - t0) / n
        out["cpu_baseline"] = {"value": 1000.0 / (ms_cpu / 1e3), "unit": "queries/s", "cores": os.cpu_count(),
                               "kind": "port", "sample": "%d full passes of the same 1000x1000x12 set" % n}
    return out


if __name__ == "__main__":
    main()

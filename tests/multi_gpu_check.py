"""Multi-GPU parity checks of the head's exchange steps (NCCL, one process per GPU).

Run directly under torchrun on a GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

or through pytest (tests/test_gpu_multirank.py spawns exactly that when >= 2 GPUs are visible); bench.py
calls run_checks() at N > 1 and prints the result as the "checks" block of its JSON line, so the driver's
scaling record carries them.  Each check compares with the numpy oracle evaluated on the rank-major
concatenation of all ranks' inputs (the oracle is the checker here, nothing of it is timed or shipped):

  1. fine-tune head (modules/modeling.py:698-709): every rank's loss equals the global loss; each rank's
     gradient equals W x its slice of the single-loss gradient (the reference's dist_collect contract);
     the replicated backward (no exchange) equals the SUM reduce-scatter backward (bit for bit at W = 2,
     within the rounding of an 8-term sum otherwise);
  2. pre-train head (modules/modeling.py:244-284): loss is rank-local; after the key all-gather + enqueue
     all ranks hold the same queues, equal to the oracle's enqueue of the rank-major concatenation --
     through the eager call, the deferred schedule and a CUDA-graph replay;
  3. sharded fused eval: ranks equal the single-GPU run of the same set;
  4. MLP with SyncBatchNorm: batch statistics span all ranks' rows (modules/modeling.py:127-129).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hmmc_b200 import modeling, parallel, retrieval   # noqa: E402
from hmmc_b200 import synthetic as syn                 # noqa: E402
from oracle import head_oracle as O                    # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def _cu(dev):
    return lambda x, g=False: torch.from_numpy(np.ascontiguousarray(x)).to(dev).requires_grad_(g)


def check_finetune(W, rank, local, dev, b=16, F=12, D=512):
    cu = _cu(dev)
    parts = [syn.finetune_inputs(b, F=F, D=D, seed=50 + r) for r in range(W)]
    t, v, fr = [np.concatenate([p[i] for p in parts], 0) for i in range(3)]
    ref_loss, dt, dv, dfr = O.finetune_loss_and_grads(t, v, fr)
    out = {}
    sl = slice(rank * b, (rank + 1) * b)
    for prec, tol in (("fp32", 3e-5), ("bf16x3", 1e-4)):
        task = types.SimpleNamespace(local_rank=local, top_frames=2, use_frame_fea=True, head_precision=prec)
        m = modeling.BirdModel(modeling.default_cross_config(), task)
        a = [cu(x, True) for x in parts[rank]]
        loss = m.head_loss(*a)
        loss.backward()
        lerr = abs(float(loss.detach()) - ref_loss) / ref_loss
        gerr = max(rel(got.grad.cpu().numpy(), W * ref[sl]) for got, ref in zip(a, (dt, dv, dfr)))
        assert lerr < tol and gerr < 3 * tol, (prec, lerr, gerr)
        out["finetune_%s_loss_rel" % prec] = lerr
        out["finetune_%s_grad_rel_vs_W_x_slice" % prec] = gerr
        if prec == "bf16x3":
            # the same step through the generic SUM reduce-scatter backward: bit-identical gradients
            saved = parallel.REPLICATED_GATHER_BWD
            parallel.REPLICATED_GATHER_BWD = False
            try:
                a2 = [cu(x, True) for x in parts[rank]]
                loss2 = m.head_loss(*a2)
                loss2.backward()
            finally:
                parallel.REPLICATED_GATHER_BWD = saved
            # W identical summands: W * x is exact, a ring that adds x eight times rounds at 3x, 5x, 6x, 7x --
            # bit-identical for W = 2, within summation rounding (a few ulp) beyond
            assert torch.equal(loss.detach(), loss2.detach())
            rs_rel = max(float((x.grad - y.grad).abs().max() / y.grad.abs().max().clamp_min(1e-30)) for x, y in zip(a, a2))
            same = all(torch.equal(x.grad, y.grad) for x, y in zip(a, a2))
            assert same if W == 2 else rs_rel < 1e-6, ("replicated backward differs from the reduce-scatter backward", rs_rel)
            out["replicated_bwd_equals_reduce_scatter"] = bool(same)
            out["replicated_bwd_vs_reduce_scatter_max_rel"] = rs_rel
    out["finetune_global_batch"] = W * b
    return out


def check_pretrain_enqueue(W, rank, local, dev, b=8, F=12, D=128, K=64):
    cu = _cu(dev)
    assert K % (W * b) == 0
    qs0 = syn.queues(K, F=F, D=D, seed=3)
    steps = 4
    inps = [[syn.pretrain_inputs(b, F=F, D=D, seed=70 + 10 * s + r) for r in range(W)] for s in range(steps)]
    # oracle: enqueue of the rank-major concatenation, step after step; the loss of step s reads the queues
    # as they are before that step's enqueue
    qs = {n: x.copy() for n, x in qs0.items()}
    ptr = 0
    ref_losses, ref_queues = [], []
    for s in range(steps):
        ref_losses.append(O.pretrain_loss(inps[s][rank], qs, 0.07, dtype=np.float64))
        cat = {n: np.concatenate([i[n] for i in inps[s]], 0) for n in inps[s][0]}
        ptr = O.dequeue_and_enqueue(qs, ptr, cat["v_fea_k"], cat["tag_fea_k"], cat["title_fea_k"], cat["frame_fea_k"],
                                    cat["frame_proj_k"], K)
        ref_queues.append(({n: x.copy() for n, x in qs.items()}, ptr))
    out = {}
    order = ["v_fea", "frame_fea", "title_fea", "frame_pred", "v_fea_k", "frame_fea_k", "title_fea_k", "tag_fea_k",
             "frame_proj_k"]
    # eager: the reference's immediate gather + enqueue.  deferred / graph: the staged keys travel over peer memory
    # (parallel.PeerExchange) at the start of the next step; *_nccl: the same schedule with the NCCL all-gather;
    # graph_noflush: replays back to back (both receive slots in use), queues compared after the last step only.
    for mode in ("eager", "deferred", "graph", "deferred_nccl", "graph_nccl", "graph_noflush"):
        task = types.SimpleNamespace(local_rank=local, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                                     contrast_num_negative=K, max_frames=F, use_frame_fea=True, head_precision="fp32",
                                     defer_enqueue=(mode != "eager"),
                                     peer_exchange=(False if mode.endswith("_nccl") else None))
        m = modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=D), task).to(dev)
        with torch.no_grad():
            for n, x in qs0.items():
                getattr(m, n).copy_(torch.from_numpy(x))
        static = {n: cu(inps[0][rank][n], n in order[:4]) for n in order}

        def step():
            for n in order[:4]:
                static[n].grad = None
            loss = m.head_loss(*[static[n] for n in order])
            loss.backward()
            return loss
        graphed = None
        worst = 0.0
        for s in range(steps):
            with torch.no_grad():
                for n in order:
                    static[n].copy_(torch.from_numpy(inps[s][rank][n]))
            if mode.startswith("graph") and s == 1:
                # step 0 ran eagerly (it sized the workspaces and left its keys staged); capture without further
                # warm-up steps and replay from here on
                from hmmc_b200.graphs import GraphedStep
                graphed = GraphedStep(step, warmup=0)
            loss = graphed.replay() if graphed is not None else step()
            lerr = abs(float(loss.detach()) - ref_losses[s]) / ref_losses[s]
            del loss        # a live loss keeps the step's autograd graph (and its AccumulateGrad streams): see graphs.py
            assert lerr < 1e-5, (mode, s, lerr)
            if mode == "graph_noflush" and s < steps - 1:
                continue
            m.flush_pending_enqueue()
            want, want_ptr = ref_queues[s]
            assert int(m.queue_ptr) == want_ptr, (mode, s, int(m.queue_ptr), want_ptr)
            for n in syn.QUEUE_NAMES:
                err = float(np.abs(getattr(m, n).cpu().numpy() - want[n]).max())
                worst = max(worst, err)
                assert err <= 2e-7, (mode, s, n, err)
        out["enqueue_%s_max_abs_vs_oracle_concat" % mode] = worst
        if mode != "eager":
            # which exchange ran: the peer-memory push (default between NCCL ranks of one node) or the all-gather
            peer_used = m._hmmc_xchg[4] is not None
            assert peer_used == (not mode.endswith("_nccl") and parallel.PeerExchange.usable()), (mode, peer_used)
            out["enqueue_%s_exchange" % mode] = "peer memory" if peer_used else "nccl all-gather"
        if graphed is not None:
            graphed.release()
    out["enqueue_steps"] = steps
    out["enqueue_global_batch"] = W * b
    return out


def check_sharded_eval(W, rank, local, dev, Nv=1000):
    cu = _cu(dev)
    rs = np.random.RandomState(5)
    per = rs.randint(1, 12, size=Nv)
    T, V, Fr, gt, _ = syn.eval_inputs(int(per.sum()), Nv, seed=33, per_video=per)
    lo, hi = parallel.shard_range(Nv, W, rank)
    t2v, v2t = retrieval.fused_eval_ranks(cu(T), cu(V[lo:hi]), cu(Fr[lo:hi]), per, 100.0, 3, "bf16x3")
    # reference: whole gallery on this GPU alone (explicit range = no collectives inside)
    saved = parallel.world
    parallel.world = lambda: (1, 0)
    try:
        t2v1, v2t1 = retrieval.fused_eval_ranks(cu(T), cu(V), cu(Fr), per, 100.0, 3, "bf16x3", video_range=(0, Nv))
    finally:
        parallel.world = saved
    mism = int((t2v != t2v1).sum()) + int((v2t != v2t1).sum())
    assert mism == 0, mism
    # and against the fp32 oracle's scores (near-ties may flip in bf16x3: none expected on this set)
    ref = O.eval_scores(T, V, Fr, 3)
    want_t = (ref > ref[np.arange(T.shape[0]), gt][:, None]).sum(1)
    flips = int((t2v.cpu().numpy() != want_t).sum())
    assert flips <= 2, flips
    return {"sharded_eval_vs_single_gpu_rank_mismatches": mism, "sharded_eval_t2v_flips_vs_fp32_oracle": flips,
            "sharded_eval_captions": int(per.sum()), "sharded_eval_videos": Nv}


def check_sync_batchnorm_mlp(W, rank, local, dev):
    cu = _cu(dev)
    from hmmc_b200.mlp import MLP
    from oracle import mlp_oracle as MO
    Mr = 64
    c = syn.mlp_case(M=Mr * W, Din=64, Dh=128, Dout=64, seed=61)
    m = MLP(64, 128, 64, 2, precision="bf16x3")
    torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)          # in place on the children, as the reference relies on
    m = m.to(dev).train()
    lin1, bn = m.linear_hidden[1], m.linear_hidden[2]
    assert isinstance(bn, torch.nn.SyncBatchNorm)
    with torch.no_grad():
        lin1.weight.copy_(cu(c["W1"])); lin1.bias.copy_(cu(c["b1"]))
        bn.weight.copy_(cu(c["gamma"])); bn.bias.copy_(cu(c["beta"]))
        bn.running_mean.copy_(cu(c["rm"])); bn.running_var.copy_(cu(c["rv"]))
        m.linear_out.weight.copy_(cu(c["W2"])); m.linear_out.bias.copy_(cu(c["b2"]))
    sl = slice(rank * Mr, (rank + 1) * Mr)
    x = cu(c["x"][sl], True)
    y = m(x)
    y.backward(cu(c["dy"][sl]))
    y64, cache = MO.forward(c["x"], c["W1"], c["b1"], c["gamma"], c["beta"], c["W2"], c["b2"])
    ref = MO.backward(c["dy"], cache, c["W1"], c["gamma"], c["W2"])
    yerr = rel(y.detach().cpu().numpy(), y64[sl])
    assert yerr < 1e-5
    # elements whose pre-activation lies within the GEMM's rounding of zero may take the other ReLU branch
    slack = MO.relu_flip_slack(cache, ref, c["W1"], c["gamma"], width=2e-5)
    err = np.linalg.norm(x.grad.cpu().numpy().astype(np.float64) - ref["dx"][sl])
    assert err < 5e-5 * np.linalg.norm(ref["dx"][sl]) + slack["dx"], (err, slack)
    rm, rv = MO.running_stats(cache, c["rm"], c["rv"])
    assert rel(bn.running_mean.cpu().numpy(), rm) < 1e-5 and rel(bn.running_var.cpu().numpy(), rv) < 1e-5
    # parameter gradients are per rank (DDP sums them): their sum over ranks is the global gradient
    for name, t in (("dW1", lin1.weight.grad), ("dgamma", bn.weight.grad), ("dbeta", bn.bias.grad),
                    ("dW2", m.linear_out.weight.grad), ("db2", m.linear_out.bias.grad)):
        tot = t.clone()
        dist.all_reduce(tot)
        err = np.linalg.norm(tot.cpu().numpy().astype(np.float64) - ref[name])
        assert err < 5e-5 * np.linalg.norm(ref[name]) + slack.get(name, 0.0), (name, err, slack)
    return {"sync_batchnorm_mlp_out_rel": yerr}


CHECKS = (check_finetune, check_pretrain_enqueue, check_sharded_eval, check_sync_batchnorm_mlp)


def run_checks(W, rank, local, dev, which=CHECKS):
    """Runs the checks on an initialised NCCL group; returns {"pass": bool over ALL ranks, numbers of this rank,
    "failures": [...]} -- never raises, so bench.py can always print its line."""
    out, failures = {}, []
    for fn in which:
        try:
            out.update(fn(W, rank, local, dev))
        except Exception as e:   # noqa: BLE001
            failures.append("%s: %s" % (fn.__name__, repr(e)[:300]))
        torch.cuda.synchronize()
        dist.barrier()
    ok = torch.tensor([0 if failures else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["pass"] = bool(int(ok.item()))
    out["failures"] = failures
    out["world_size"] = W
    out["checker"] = "oracle/head_oracle.py on the rank-major concatenation of all ranks' inputs"
    return out


def main():
    W = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = run_checks(W, rank, local, dev)
    if rank == 0 or res["failures"]:
        print("multi_gpu_check rank %d W=%d: %s" % (rank, W, res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not res["pass"]:
        sys.exit(1)
    if rank == 0:
        print("multi_gpu_check W=%d ok" % W, flush=True)


if __name__ == "__main__":
    main()

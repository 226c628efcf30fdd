"""Multi-GPU parity check, run under torchrun on a GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Checks, against the numpy oracle evaluated on the gathered inputs:
  1. fine-tune head: every rank's loss equals the global loss; each rank's gradient equals
     W x its slice of the single-loss gradient (the reference's dist_collect contract);
  2. pre-train head: loss is rank-local; after the key all-gather + enqueue all ranks hold the
     same queues, equal to the oracle's enqueue of the rank-major concatenation;
  3. sharded fused eval: ranks equal the single-GPU run of the same set.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hmmc_b200 import modeling, ops, parallel, retrieval   # noqa: E402
from hmmc_b200 import synthetic as syn                     # noqa: E402
from oracle import head_oracle as O                        # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def main():
    W = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cu = lambda x, g=False: torch.from_numpy(np.ascontiguousarray(x)).to(dev).requires_grad_(g)

    # 1. fine-tune head
    b, F, D = 16, 12, 512
    parts = [syn.finetune_inputs(b, F=F, D=D, seed=50 + r) for r in range(W)]
    t, v, fr = [np.concatenate([p[i] for p in parts], 0) for i in range(3)]
    ref_loss, dt, dv, dfr = O.finetune_loss_and_grads(t, v, fr)
    for prec, tol in (("fp32", 3e-5), ("bf16x3", 1e-4)):
        task = types.SimpleNamespace(local_rank=local, top_frames=2, use_frame_fea=True, head_precision=prec)
        m = modeling.BirdModel(modeling.default_cross_config(), task)
        a = [cu(x, True) for x in parts[rank]]
        loss = m.head_loss(*a)
        loss.backward()
        assert abs(float(loss) - ref_loss) / ref_loss < tol, (prec, float(loss), ref_loss)
        sl = slice(rank * b, (rank + 1) * b)
        for got, ref in zip(a, (dt, dv, dfr)):
            assert rel(got.grad.cpu().numpy(), W * ref[sl]) < tol * 3, prec
    if rank == 0:
        print("fine-tune head W=%d ok: loss %.6f" % (W, ref_loss))

    # 2. pre-train head + enqueue
    b, F, D, K = 8, 12, 128, 64
    qs = syn.queues(K, F=F, D=D, seed=3)
    inps = [syn.pretrain_inputs(b, F=F, D=D, seed=70 + r) for r in range(W)]
    task = types.SimpleNamespace(local_rank=local, top_frames=3, contrast_momentum=0.99, contrast_temperature=0.07,
                                 contrast_num_negative=K, max_frames=F, use_frame_fea=True, head_precision="fp32")
    m = modeling.BirdPreTrainedModel(modeling.default_cross_config(temporal_hidden_size=D), task).to(dev)
    with torch.no_grad():
        for n, x in qs.items():
            getattr(m, n).copy_(torch.from_numpy(x))
    mine = inps[rank]
    tt = {n: cu(x, n in ("v_fea", "title_fea", "frame_fea", "frame_pred")) for n, x in mine.items()}
    loss = m.head_loss(tt["v_fea"], tt["frame_fea"], tt["title_fea"], tt["frame_pred"], tt["v_fea_k"],
                       tt["frame_fea_k"], tt["title_fea_k"], tt["tag_fea_k"], tt["frame_proj_k"])
    ref = O.pretrain_loss(mine, qs, 0.07, dtype=np.float64)
    assert abs(float(loss) - ref) / ref < 1e-5
    cat = {n: np.concatenate([i[n] for i in inps], 0) for n in inps[0]}
    ptr = O.dequeue_and_enqueue(qs, 0, cat["v_fea_k"], cat["tag_fea_k"], cat["title_fea_k"], cat["frame_fea_k"],
                                cat["frame_proj_k"], K)
    assert int(m.queue_ptr) == ptr
    for n in syn.QUEUE_NAMES:
        np.testing.assert_allclose(getattr(m, n).cpu().numpy(), qs[n], rtol=0, atol=2e-7)
    if rank == 0:
        print("pre-train head + gathered enqueue W=%d ok: ptr %d" % (W, ptr))

    # 3. sharded fused eval vs the single-GPU run
    rs = np.random.RandomState(5)
    Nv = 1000
    per = rs.randint(1, 12, size=Nv)
    T, V, Fr, gt, _ = syn.eval_inputs(int(per.sum()), Nv, seed=33, per_video=per)
    lo, hi = parallel.shard_range(Nv, W, rank)
    t2v, v2t = retrieval.fused_eval_ranks(cu(T), cu(V[lo:hi]), cu(Fr[lo:hi]), per, 100.0, 3, "bf16x3")
    # reference: whole gallery on this GPU alone (explicit range = no collectives inside)
    group = dist.group.WORLD
    saved = parallel.world
    parallel.world = lambda: (1, 0)
    t2v1, v2t1 = retrieval.fused_eval_ranks(cu(T), cu(V), cu(Fr), per, 100.0, 3, "bf16x3", video_range=(0, Nv))
    parallel.world = saved
    assert torch.equal(t2v, t2v1) and torch.equal(v2t, v2t1), (int((t2v != t2v1).sum()), int((v2t != v2t1).sum()))
    if rank == 0:
        print("sharded fused eval W=%d ok: %d captions x %d videos, mean t2v rank %.3f" %
              (W, int(per.sum()), Nv, float(t2v.float().mean()) + 1))

    # 4. MLP with SyncBatchNorm: the batch statistics span all ranks' rows (modules/modeling.py:127-129)
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
    if os.environ.get("HMMC_CHECK_DEBUG"):
        print("rank", rank, "y", rel(y.detach().cpu().numpy(), y64[sl]), "dx", rel(x.grad.cpu().numpy(), ref["dx"][sl]),
              "dgamma_local_sum?", float(bn.weight.grad.abs().sum()), flush=True)
    assert rel(y.detach().cpu().numpy(), y64[sl]) < 1e-5
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
    if rank == 0:
        print("MLP + SyncBatchNorm W=%d ok: %d rows per rank" % (W, Mr))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

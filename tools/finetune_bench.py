"""Launch list of the fine-tune head (sym_ce fwd+bwd) at one batch size: run under
ncu --metrics gpu__time_duration.sum to see the per-kernel device times."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops, synthetic as syn

B = int(os.environ.get("B", "256"))
prec = ops.resolve_precision(os.environ.get("PREC", "bf16x3"))
t, v, fr = [torch.from_numpy(x).cuda() for x in syn.finetune_inputs(B, seed=1)]
for _ in range(int(os.environ.get("REPS", "3"))):
    out = ops.sym_ce_raw(t, v, fr, 100.0, 0.85, 0.15, prec, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.sym_ce_raw(t, v, fr, 100.0, 0.85, 0.15, prec, True)
e1.record()
torch.cuda.synchronize()
print("B=%d eager %.1f us per call, loss %.5f" % (B, e0.elapsed_time(e1) / 20 * 1e3, float(out[0])))

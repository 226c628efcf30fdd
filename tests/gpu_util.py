import numpy as np
import torch


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cu(x, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    if grad:
        t.requires_grad_(True)
    return t


# tolerances: (loss relative, gradient relative-L2) per precision mode.
# north_star: 1e-5 in fp32, 1e-3 in bf16; the fp32 reference itself carries ~2e-5 gradient noise at
# logit scale 100 (tests/test_oracle_golden.py), so gradients are checked against the fp64 oracle.
TOL = {"fp32": (1e-5, 2e-5), "bf16x3": (1e-5, 5e-5), "bf16": (1e-3, 1e-3)}

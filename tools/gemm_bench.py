"""Throughput of the tcgen05 GEMM engine on a few shapes, next to torch.matmul (cuBLAS) bf16."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops

import os
if os.environ.get("SHAPES"):
    shapes = [tuple(int(x) for x in t.split("x")) for t in os.environ["SHAPES"].split(",")]
else:
    shapes = None
_default = [(1536, 12288, 512), (1536, 512, 12288), (4096, 4096, 4096), (8192, 8192, 2048), (3072, 12288, 512)]
shapes = shapes or _default
for M, N, K in shapes:
    A = (torch.randn(M, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    TILING = int(os.environ.get("TILING", "0"))      # 0 auto, 128 / 256 single-CTA tile width, 512 CTA-pair kernel
    for name, fn in (("hmmc umma", lambda: ops.umma_gemm_nt(A, B, K, 1, 1.0, TILING)), ("torch bf16", lambda: A @ B.t())):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%-10s BN=%s M=%5d N=%5d K=%5d  %8.1f us  %7.1f TFLOP/s" % (name, os.environ.get("TILING", "auto"), M, N, K, ms * 1e3, 2.0 * M * N * K / ms / 1e9))

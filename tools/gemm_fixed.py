import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmmc_b200 import ops
for M, N, K in [(256, 256, 64), (256, 256, 512), (3072, 1024, 512), (3072, 2048, 512), (3072, 4096, 512), (3072, 12288, 512), (3072, 12288, 1024)]:
    A = (torch.randn(M, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    for _ in range(3):
        ops.umma_gemm_nt(A, B, K, 1, 1.0)
torch.cuda.synchronize()

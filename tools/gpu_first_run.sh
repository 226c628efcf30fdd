#!/bin/bash
# staged GPU validation: safe kernels first, tcgen05 kernels under their own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== stage 1: fp32 / memory-bound kernels" 
timeout 600 python -m pytest tests -m gpu -q -k "not umma and not bf16 and not tensor_core" 2>&1 | tail -25 | tee gpurun_out/stage1.log
echo "== stage 2: tcgen05 GEMM primitive"
timeout 300 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "umma" 2>&1 | tail -40 | tee gpurun_out/stage2.log
echo "== stage 3: tensor-core paths of the heads"
timeout 600 python -m pytest tests -m gpu -q -k "bf16 or tensor_core" 2>&1 | tail -40 | tee gpurun_out/stage3.log
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -15 | tee gpurun_out/smoke.log

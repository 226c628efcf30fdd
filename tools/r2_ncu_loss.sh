#!/bin/bash
# ncu --set full of the shipped loss kernels at the north-star shape (b=256, bf16): third iteration of tools/loss_kernels.py
mkdir -p gpurun_out/r2d
B=256 PREC=bf16 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none \
  -k regex:"umma_gemm_pair_kernel|infonce_finish_kernel|prep_rows_kernel" --launch-skip 8 --launch-count 4 \
  -o gpurun_out/r2d/loss_b256 -f python tools/loss_kernels.py > gpurun_out/r2d/ncu.log 2>&1
tail -5 gpurun_out/r2d/ncu.log
ls -la gpurun_out/r2d

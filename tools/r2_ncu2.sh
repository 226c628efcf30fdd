#!/bin/bash
mkdir -p gpurun_out/r2f
NCU="ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv"
NOGRAD=1 B=256 PREC=bf16 timeout 300 $NCU --log-file gpurun_out/r2f/loss_b256_nograd.csv python tools/loss_kernels.py > gpurun_out/r2f/nograd.out 2>&1
B=256 PREC=bf16 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none \
  -k regex:"umma_gemm_pair_kernel|infonce_finish" --launch-skip 6 --launch-count 3 \
  -o gpurun_out/r2f/loss_b256 -f python tools/loss_kernels.py > gpurun_out/r2f/ncu.log 2>&1
tail -3 gpurun_out/r2f/ncu.log
grep -h "umma_gemm_pair\|finish" gpurun_out/r2f/loss_b256_nograd.csv | cut -d, -f5,15 | tail -4

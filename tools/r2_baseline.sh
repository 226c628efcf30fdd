#!/bin/bash
# round-2 starting point: timings of the three latency targets + warm-cache ncu launch lists
mkdir -p gpurun_out/r2a
timeout 600 python tools/head_bench.py > gpurun_out/r2a/head_bench.log 2>&1
NCU="ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv"
B=256 PREC=bf16 timeout 300 $NCU --log-file gpurun_out/r2a/loss_b256.csv python tools/loss_kernels.py > gpurun_out/r2a/loss_b256.out 2>&1
B=128 PREC=bf16 timeout 300 $NCU --log-file gpurun_out/r2a/loss_b128.csv python tools/loss_kernels.py > gpurun_out/r2a/loss_b128.out 2>&1
timeout 300 $NCU --log-file gpurun_out/r2a/finetune.csv python tools/finetune_bench.py > gpurun_out/r2a/finetune.out 2>&1
tail -20 gpurun_out/r2a/head_bench.log

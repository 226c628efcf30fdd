#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pretrain.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/r2c/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
tail -15 gpurun_out/r2c/pytest.log
timeout 600 python tools/head_bench.py > gpurun_out/r2c/head_bench.log 2>&1
tail -14 gpurun_out/r2c/head_bench.log
NCU="ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv"
B=256 PREC=bf16 timeout 300 $NCU --log-file gpurun_out/r2c/loss_b256.csv python tools/loss_kernels.py > gpurun_out/r2c/loss_b256.out 2>&1
B=128 PREC=bf16 timeout 300 $NCU --log-file gpurun_out/r2c/loss_b128.csv python tools/loss_kernels.py > gpurun_out/r2c/loss_b128.out 2>&1
tail -3 gpurun_out/r2c/loss_b256.out

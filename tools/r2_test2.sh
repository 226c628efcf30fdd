#!/bin/bash
mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_pretrain.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/r2e/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e/pytest.log
tail -5 gpurun_out/r2e/pytest.log
timeout 600 python tools/head_bench.py > gpurun_out/r2e/head_bench.log 2>&1
tail -4 gpurun_out/r2e/head_bench.log
NCU="ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv"
B=256 PREC=bf16 timeout 300 $NCU --log-file gpurun_out/r2e/loss_b256.csv python tools/loss_kernels.py > gpurun_out/r2e/loss_b256.out 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2e/bench.err

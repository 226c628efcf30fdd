#!/bin/bash
D=gpurun_out/$1; mkdir -p $D
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $D/check.log 2>&1
echo "check rc=$?"
grep "multi_gpu_check" $D/check.log | cut -c1-1800
timeout 900 python -m pytest tests/test_gpu_finetune.py tests/test_gpu_pretrain.py tests/test_gpu_edge_cases.py -x -q -m gpu > $D/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $D/pytest.log
timeout 300 python tools/head_bench.py 2>&1 | grep -v Warn | head -9 | tee $D/head_bench.log
timeout 200 python tools/variant_sweep.py 2>&1 | grep variant | tee $D/sweep.log

#!/bin/bash
# usage: r2_ft.sh <outdir>  -- fine-tune / eval GPU tests, then their graph-replay timings
D=gpurun_out/$1; mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_finetune.py tests/test_gpu_eval.py -x -q -m gpu > $D/pytest.log 2>&1
echo "pytest rc=$?" >> $D/pytest.log
tail -3 $D/pytest.log
timeout 300 python tools/ft_eval_bench.py 2>&1 | grep -v Warning | tee $D/ft_eval.log

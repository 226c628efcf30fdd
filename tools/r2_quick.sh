#!/bin/bash
# usage: r2_quick.sh <outdir> [variant ...]   -- pretrain/edge GPU tests, loss timing per library variant, ncu launch list
D=gpurun_out/$1; mkdir -p $D; shift
timeout 900 python -m pytest tests/test_gpu_pretrain.py tests/test_gpu_edge_cases.py -x -q -m gpu > $D/pytest.log 2>&1
echo "pytest rc=$?" >> $D/pytest.log
tail -3 $D/pytest.log
timeout 200 python tools/variant_sweep.py 2>&1 | grep variant | tee $D/sweep.log
for v in "$@"; do VARIANT=$v timeout 200 python tools/variant_sweep.py 2>&1 | grep variant | tee -a $D/sweep.log; done
NCU="ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv"
B=256 PREC=bf16 timeout 300 $NCU --log-file $D/loss_b256.csv python tools/loss_kernels.py > $D/loss_b256.out 2>&1
grep "umma_gemm\|finish\|prep_rows\|scale_tensors" $D/loss_b256.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for r in rows[-5:]: print(r[4][:70], r[-1])"

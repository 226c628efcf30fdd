#!/bin/bash
D=gpurun_out/$1; mkdir -p $D
timeout 1500 python -m pytest tests -x -q -m gpu > $D/pytest.log 2>&1
echo "pytest rc=$?" >> $D/pytest.log
tail -6 $D/pytest.log
timeout 600 python tools/head_bench.py > $D/head_bench.log 2>&1
tail -4 $D/head_bench.log
NCU="ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv"
B=256 PREC=bf16 timeout 300 $NCU --log-file $D/loss_b256.csv python tools/loss_kernels.py > $D/loss_b256.out 2>&1
python - <<'P'
import csv,collections,sys,os
d=collections.OrderedDict()
fn=os.path.join(sys.argv[1] if len(sys.argv)>1 else "", "loss_b256.csv")
P
grep "umma_gemm\|finish\|prep_rows\|scale_tensors" $D/loss_b256.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for r in rows[-5:]: print(r[4][:70], r[-1])"

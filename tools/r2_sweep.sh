#!/bin/bash
D=gpurun_out/$1; mkdir -p $D
for v in "" A B C D E F G H; do
  VARIANT=$v timeout 200 python tools/variant_sweep.py >> $D/sweep.log 2>&1
done
cat $D/sweep.log | grep variant
NCU="ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv"
B=256 PREC=bf16 timeout 300 $NCU --log-file $D/loss_b256.csv python tools/loss_kernels.py > $D/loss_b256.out 2>&1
grep "umma_gemm\|finish\|prep_rows\|scale_tensors" $D/loss_b256.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for r in rows[-5:]: print(r[4][:70], r[-1])"

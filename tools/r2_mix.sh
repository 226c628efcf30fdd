#!/bin/bash
D=gpurun_out/$1; mkdir -p $D
for f in 1 0; do for p in fp32 bf16; do echo "== FLUSH=$f PREC=$p"; FLUSH=$f PREC=$p timeout 120 python tools/debug_capture.py 2>&1 | grep -v Warning | tail -14; done; done > $D/debug.log 2>&1
cat $D/debug.log | cut -c1-400
timeout 1500 python -m pytest tests -x -q -m gpu > $D/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $D/pytest.log
timeout 300 python tools/head_bench.py 2>&1 | grep -v Warn | head -9 | tee $D/head_bench.log
timeout 200 python tools/variant_sweep.py 2>&1 | grep variant | tee $D/sweep.log

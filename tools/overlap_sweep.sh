#!/bin/bash
# N=1 step time for several SM budgets of the loss GEMMs running beside the EMA
for r in ${SWEEP:-100 120 132}; do
  HMMC_LOSS_GEMM_RESERVED=$r timeout 200 python bench.py --no-cpu-baseline --no-retrieval --steps 300 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown_ms']; print('reserved', $r, 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'seq', b['sequential_step'], 'ema', round(b['ema'],4), 'head', round(b['head_fwd_bwd_enqueue'],4))"
done

#!/bin/bash
# Sweep the SMs left to the key all-gather while it overlaps the loss GEMMs (N = $1 ranks).
N=${1:-2}
for r in ${SWEEP:-0 8 16 24 32}; do
  HMMC_GATHER_RESERVED_SMS=$r timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port $((29600 + r)) bench.py --gpus $N --steps 300 --no-cpu-baseline --no-retrieval 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reserved', $r, 'ms', round(d['ms_per_step'],4), 'value', round(d['value']))"
done

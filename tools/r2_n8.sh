#!/bin/bash
# usage: r2_n8.sh <outdir> <ngpus>   -- the bench line (with its checks block) at N GPUs, driver defaults
D=gpurun_out/$1; N=${2:-8}; mkdir -p $D
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > $D/bench_n$N.json 2> $D/bench_n$N.err
echo "bench rc=$?"
tail -c 400 $D/bench_n$N.err
python - <<P
import json
L=json.loads(open("$D/bench_n$N.json").read().strip().splitlines()[-1])
print({k:L.get(k) for k in ("value","ms_per_step","n_gpus","gpu_launches","e2e")}); print(L.get("checks")); print(L.get("breakdown_ms"))
P

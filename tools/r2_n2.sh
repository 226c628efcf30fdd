#!/bin/bash
D=gpurun_out/$1; mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pretrain.py -x -q -m gpu > $D/pytest.log 2>&1
echo "pytest rc=$?" >> $D/pytest.log
tail -25 $D/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $D/bench_n2.json 2> $D/bench_n2.err
echo "bench rc=$?"
tail -c 600 $D/bench_n2.err
python - <<P
import json
L=json.loads(open("$D/bench_n2.json").read().strip().splitlines()[-1])
print({k:L[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}, L.get("checks"), L["breakdown_ms"])
P

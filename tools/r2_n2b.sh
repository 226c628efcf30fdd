#!/bin/bash
D=gpurun_out/$1; mkdir -p $D
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $D/check.log 2>&1
echo "check rc=$?"
grep -v "Warning\|warn" $D/check.log | tail -25

#!/bin/bash
# usage: r2_sweep.sh <outdir> [variant ...]   -- loss timing (graph replay) for the in-tree library and each variant
D=gpurun_out/$1; mkdir -p $D; shift
timeout 200 python tools/variant_sweep.py 2>&1 | grep variant | tee $D/sweep.log
for v in "$@"; do VARIANT=$v timeout 200 python tools/variant_sweep.py 2>&1 | grep variant | tee -a $D/sweep.log; done

#!/bin/bash
# usage: r2_launches.sh <outdir>  -- ncu launch list (gpu__time_duration.sum) of the default bench command, short
D=gpurun_out/$1; mkdir -p $D
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $D/plain.json 2> $D/plain.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1500 --csv --log-file $D/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $D/ncu.json 2> $D/ncu.err; echo "ncu rc=$?"
wc -l $D/launches.csv

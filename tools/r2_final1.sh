#!/bin/bash
# usage: r2_final1.sh <outdir>  -- the whole GPU suite, smoke(), then the N=1 bench line (ours + both reference arms)
D=gpurun_out/$1; mkdir -p $D
timeout 1500 python -m pytest tests/ -x -q -m gpu > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log; tail -3 $D/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $D/smoke.log 2>&1; tail -1 $D/smoke.log
timeout 900 python bench.py > $D/bench_n1.json 2> $D/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_ref.json 2> $D/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 2 > $D/bench_refgpu.json 2> $D/bench_refgpu.err; echo "refgpu rc=$?"
python - <<P
import json
for f in ("bench_n1","bench_ref","bench_refgpu"):
    try:
        L=json.loads(open("$D/%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:L.get(k) for k in ("impl","value","ms_per_step","gpu_launches")}, (L.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
P

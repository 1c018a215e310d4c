#!/bin/bash
# bench.py under several env configurations (one line each)
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'])
"; }
for v in "$@"; do run "$v"; done

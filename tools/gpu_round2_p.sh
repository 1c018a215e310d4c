#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_ops.py -q -x -k "basic_block" > gpurun_out/t_b64.log 2>&1; echo "basic_block tests rc=$?"; tail -n 12 gpurun_out/t_b64.log
timeout 120 python tools/bblock_bench.py 256 64 28 2>&1 | tail -n 5
POCO_B200_BLOCK64_RESIDENT=1 timeout 120 python tools/bblock_bench.py 256 64 28 2>&1 | tail -n 2
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'], d['parity_err'])
"; tail -n 2 gpurun_out/bench_var.err | cut -c1-200; }
run X=1
run POCO_B200_FUSE_BLOCK64=1
run X=2
run POCO_B200_FUSE_BLOCK64=1
run POCO_B200_FUSE_BLOCK64=1 POCO_B200_COST=105,70,170,260

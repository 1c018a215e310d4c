#!/bin/bash
# round 2, session 5: per-op table of the shipped plan + lane cost / share / fused-64 variants, one box, one call
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'])
"; }
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-mode --dump-ops gpurun_out/r02g_ops_b256_fp16.csv > gpurun_out/r02g_bench_ops.log 2> gpurun_out/r02g_bench_ops.err
tail -n 1 gpurun_out/r02g_bench_ops.log | cut -c1-300
run X=1
run POCO_B200_COST=80,105,170,260
run POCO_B200_COST=60,105,170,260
run POCO_B200_COST=80,105,200,300
run POCO_B200_FUSE_BLOCK64=1
run POCO_B200_FUSE_BLOCK64=1 POCO_B200_COST=80,90,170,260
run POCO_B200_SHARE_SCALE=1.5
run POCO_B200_SHARE_SCALE=3
run POCO_B200_FUSE_BLOCK=0
run X=2

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_ops.py -q -x -k "branch" > gpurun_out/t_branch.log 2>&1; echo "branch tests rc=$?"; tail -n 5 gpurun_out/t_branch.log
BR_PROF=1 timeout 120 python tools/branch_bench.py 256 4 2>&1 | head -n 6
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'])
"; }
run X=1
run POCO_B200_BRANCH_COST=1.0
run POCO_B200_BRANCH_COST=0.35
run POCO_B200_BRANCH_COST=0.5 POCO_B200_COST=85,105,170,260
run POCO_B200_BRANCH_COST=0.5 POCO_B200_COST=85,115,170,290
run POCO_B200_BRANCH_COST=0.35 POCO_B200_COST=85,115,170,290
run POCO_B200_FUSE_BRANCH=0
run X=2

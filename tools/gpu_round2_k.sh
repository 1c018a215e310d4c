#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'])
"; }
run X=1
run POCO_B200_LANE_PRIO=-1,-2,-3
run POCO_B200_LANE_PRIO=-3,-2,-1
run POCO_B200_LANE_PRIO=-2,0,-2
run POCO_B200_LANE_PRIO=-1,-2,-3 POCO_B200_SHARE_SCALE=3
run POCO_B200_LANE_PRIO=-1,-2,-3 POCO_B200_SHARE_SCALE=0
run POCO_B200_LANE_PRIO=-1,-1,-3 POCO_B200_BRANCH_COST=0.35
run X=2

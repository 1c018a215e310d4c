#!/bin/bash
# fuse(m) + branches(m + 1) in one fork / join region (POCO_B200_MERGE_PHASES=1): e2e tests in that mode, then A/B on one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
POCO_B200_MERGE_PHASES=1 timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x > gpurun_out/t_merge.log 2>&1; echo "e2e tests (merged) rc=$?"; tail -n 2 gpurun_out/t_merge.log
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'], d['parity_err']['pred_pose'])
"; }
run X=1
run POCO_B200_MERGE_PHASES=1
run X=2
run POCO_B200_MERGE_PHASES=1

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'])
"; }
timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -k "conv_tc or basic_block or branch or bottleneck" > gpurun_out/t_q.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/t_q.log
run POCO_B200_QUANT=0
run X=1
run POCO_B200_SHARE_SCALE=2.5
run POCO_B200_SHARE_SCALE=3
run POCO_B200_QUANT=0
run X=2
run POCO_B200_SHARE_SCALE=2.5
run POCO_B200_SHARE_SCALE=3

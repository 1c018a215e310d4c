#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_ops.py -q -x -k "basic_block" > gpurun_out/t_bb.log 2>&1; echo "basic_block tests rc=$?"; tail -n 3 gpurun_out/t_bb.log
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'], d['parity_err']['pred_pose'])
"; tail -n 1 gpurun_out/bench_var.err | cut -c1-160; }
for i in 1 2; do
run X=1
run POCO_B200_FUSE_BLOCK_S2D=0
done

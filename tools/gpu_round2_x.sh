#!/bin/bash
# Bottleneck tail with four epilogue warp sets (default build) against two (tools/libpoco_b200_tailsets2.so), one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_ops.py -q -x -k "bottleneck_tail" > gpurun_out/t_tail.log 2>&1; echo "tail tests rc=$?"; tail -n 2 gpurun_out/t_tail.log
echo "-- 4 sets"; timeout 120 python tools/btail_bench.py 256 2>&1 | tail -n 4
echo "-- 2 sets"; POCO_B200_LIB=$PWD/tools/libpoco_b200_tailsets2.so timeout 120 python tools/btail_bench.py 256 2>&1 | tail -n 4
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'], d['lib']['sha256_16'])
"; }
for i in 1 2; do
run X=1
run POCO_B200_LIB=$PWD/tools/libpoco_b200_tailsets2.so
done

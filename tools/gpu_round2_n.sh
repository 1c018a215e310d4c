#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_split.py -q -x -k "upsample or fuse" > gpurun_out/t_up.log 2>&1; echo "upsample tests rc=$?"; tail -n 4 gpurun_out/t_up.log
run() { echo "== $*"; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-mode $EXTRA 2> gpurun_out/bench_var.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_forward'], d['parity_err'])
"; }
EXTRA="--dump-ops gpurun_out/r02h_ops_b256_fp16.csv" run X=1
grep upsample gpurun_out/r02h_ops_b256_fp16.csv
run POCO_B200_UPSAMPLE_PER_PIXEL=1
run POCO_B200_OUT_LANES=1
run X=2
run POCO_B200_UPSAMPLE_PER_PIXEL=1
run POCO_B200_OUT_LANES=1

#!/bin/bash
# quick evaluation of a kernel change: conv / split tests, then both precision modes with per-op tables
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-eval}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/t_conv.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t_conv.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline --dump-ops gpurun_out/ops_${tag}_fp16.csv > gpurun_out/bench_${tag}.log 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${tag}.log').read().strip().splitlines()[-1])
print('fp16', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
o=d['other_precision_mode']; print('split', o['value'], o['ms_per_step'], o['parity_err'])
print(d['kernel_time_share'])
PY

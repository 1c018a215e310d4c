#!/bin/bash
# last pass of round 2 on the shipped library (sha unchanged) with the final plan defaults: gpu suite, smoke, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -n 2 gpurun_out/r02h_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02h_smoke.log
timeout 400 python bench.py --dump-ops gpurun_out/r02h_ops_b256_fp16.csv > gpurun_out/r02h_bench_fp16.json 2> gpurun_out/r02h_bench_fp16.err; echo "bench fp16 rc=$?"
timeout 120 python tools/region_times.py 256 > gpurun_out/r02h_region_times_b256.txt 2>&1
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02h_bench_fp16.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'], d['roofline']['traffic_lib_matches_this_build'],
      'all', d['roofline']['tensor_kernels']['all'], 'launches', d['launches_per_forward'], 'other', d['other_precision_mode']['value'], 'cpu', d['cpu_baseline']['value'], d['lib'])
PY

#!/bin/bash
# end-of-round evidence, round 2 session 5 (one box, one call): full gpu suite, smoke, both bench lines, per-op tables,
# region times, ncu launch list + DRAM traffic per tensor-kernel launch stamped with the library hash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > gpurun_out/r02h_device.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 gpurun_out/r02h_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02h_smoke.log
timeout 400 python bench.py --dump-ops gpurun_out/r02h_ops_b256_fp16.csv > gpurun_out/r02h_bench_fp16.json 2> gpurun_out/r02h_bench_fp16.err; echo "bench fp16 rc=$?"
timeout 400 python bench.py --precision split --no-other-mode --no-cpu-baseline --steps 10 --dump-ops gpurun_out/r02h_ops_b256_split.csv > gpurun_out/r02h_bench_split.json 2> gpurun_out/r02h_bench_split.err; echo "bench split rc=$?"
timeout 120 python tools/region_times.py 256 > gpurun_out/r02h_region_times_b256.txt 2>&1
timeout 120 python tools/branch_bench.py 256 4 > gpurun_out/r02h_branch_bench.txt 2>&1
timeout 120 python tools/bblock_bench.py 256 64 28 > gpurun_out/r02h_bblock64_bench.txt 2>&1
timeout 120 python tools/bblock_bench.py 256 32 56 >> gpurun_out/r02h_bblock64_bench.txt 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r02h_bench_fp16.json', 'gpurun_out/r02h_bench_split.json'):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'],
              'all', d['roofline'].get('tensor_kernels', {}).get('all'), 'launches', d['launches_per_forward'], 'parity', d['parity_err'], 'cpu', d['cpu_baseline'])
    except Exception as e:
        print(f, 'parse failed', e)
PY
TAG=r02h timeout 1500 bash tools/ncu_traffic.sh > gpurun_out/r02h_ncu_traffic.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/r02h_ncu_traffic.log | cut -c1-600

#!/bin/bash
# end-of-round evidence (session 3): full gpu suite, smoke, default bench, SMPL stage throughput
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_verbose.log 2>&1; echo "pytest gpu rc=$?"
grep -E "passed|failed" gpurun_out/t_gpu_verbose.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/t_gpu_verbose.log | head
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench_cliff_w32.log 2> gpurun_out/bench_cliff_w32.err; echo "bench rc=$?"
timeout 60 python tools/smpl_bench.py 256 30 > gpurun_out/smpl_bench2.log 2>&1; echo "smpl bench rc=$?"; tail -n 1 gpurun_out/smpl_bench2.log
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_cliff_w32.log').read().splitlines() if l.startswith('{')][-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],'launches',d['gpu_launches'],'cpu',d['cpu_baseline'])
except Exception as e: print('parse failed', e)
PY

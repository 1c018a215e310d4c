#!/bin/bash
# end-of-round evidence: full gpu suite with parity numbers, smoke, default bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_verbose.log 2>&1; echo "pytest gpu rc=$?"
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_cliff_w32.log 2> gpurun_out/bench_cliff_w32.err; echo "bench rc=$?"
grep -E "passed|failed" gpurun_out/t_gpu_verbose.log | tail -2
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_cliff_w32.log').read().splitlines() if l.startswith('{')][-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],'launches',d['gpu_launches'])
except Exception as e: print('parse failed', e)
PY

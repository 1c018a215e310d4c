#!/bin/bash
# quick GPU check: conv diagnostics, conv tests, e2e goldens, bench with per-op dump
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/diag_conv.py all > gpurun_out/diag.log 2>&1; echo "diag rc=$?"
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops_b256.csv ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -c OK gpurun_out/diag.log; grep FAIL gpurun_out/diag.log; tail -n 3 gpurun_out/t_gpu.log; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],d['kernel_time_share'])
except Exception as e: print('bench parse failed',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
head -n 25 gpurun_out/ops_b256.csv

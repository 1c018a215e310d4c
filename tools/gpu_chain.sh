#!/bin/bash
# chain bring-up: chain parity tests first (bounded), then the full GPU suite and the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "chain" > gpurun_out/t_chain.log 2>&1; rc=$?; echo "chain tests rc=$rc"
tail -n 15 gpurun_out/t_chain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -n 5 gpurun_out/t_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops_b256.csv ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],d['kernel_time_share'])
except Exception as e: print('bench parse failed',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
head -n 30 gpurun_out/ops_b256.csv

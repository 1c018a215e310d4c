#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -k "dx_in_n" > gpurun_out/t_dxn.log 2>&1; echo "dxn rc=$?"; tail -n 6 gpurun_out/t_dxn.log; timeout 600 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/t_ops.log 2>&1; rc=$?; echo "ops tests rc=$rc"; tail -n 12 gpurun_out/t_ops.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_ops.py > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rest rc=$?"; tail -n 4 gpurun_out/t_gpu.log
timeout 300 python tools/region_times.py 256 > gpurun_out/region_times.txt 2>&1; head -3 gpurun_out/region_times.txt; tail -n 3 gpurun_out/region_times.txt
timeout 400 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops_b256.csv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],d['kernel_time_share'])
except Exception as e: print('bench parse failed',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY

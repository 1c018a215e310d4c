#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/t_ops.log 2>&1; rc=$?; echo "ops tests rc=$rc"; tail -n 8 gpurun_out/t_ops.log
if [ $rc -ne 0 ]; then exit 1; fi
for g in 1 2 4; do echo "== MGROUP $g"; POCO_B200_MGROUP=$g timeout 200 python tools/conv_bench.py 256 0 2>&1 | grep -v "s2" ; done
echo "== MGROUP 4 roles"; timeout 200 python tools/conv_bench.py 256 1,2,3 2>&1 | grep -E "32->32|64->64 k3 s1 h28"
timeout 400 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops_b256.csv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'],d['kernel_time_share'])
except Exception as e: print('bench parse failed',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY

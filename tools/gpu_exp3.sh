#!/bin/bash
# session-3 evidence: ncu of the SMPL stage kernels, stock-PyTorch-eager on the same B200, e2e with the mesh stage
mkdir -p gpurun_out
timeout 90 ncu --set full --clock-control none --import-source on -k regex:smpl_ -c 3 -f -o gpurun_out/r01d_smpl_kernels python tools/smpl_bench.py 256 1 > gpurun_out/ncu_smpl.log 2>&1; echo "ncu rc=$?"
timeout 90 python tools/torch_eager_b200.py cliff_w32 256 3 > gpurun_out/torch_eager.log 2> gpurun_out/torch_eager.err; echo "eager rc=$?"; tail -n 1 gpurun_out/torch_eager.log
timeout 120 python bench.py --with-smpl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_with_smpl.log 2> gpurun_out/bench_with_smpl.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_with_smpl.log').read().splitlines() if l.startswith('{')][-1])
    print('value',d['value'],'e2e',d['e2e'])
except Exception as e: print('parse failed', e)
PY

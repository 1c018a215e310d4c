#!/bin/bash
# end-of-round evidence (session 2): full gpu suite with parity numbers, smoke, default bench, other presets
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_verbose.log 2>&1; echo "pytest gpu rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_cliff_w32.log 2> gpurun_out/bench_cliff_w32.err; echo "bench rc=$?"
for P in "pare_w32 128" "cliff_w48cls 256" "pare_r50 256"; do set -- $P; timeout 300 python bench.py --preset $1 --batch $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.log 2> gpurun_out/bench_$1.err; echo "bench $1 rc=$?"; done
grep -E "passed|failed" gpurun_out/t_gpu_verbose.log | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'value',d.get('value'),'ms/step',d.get('ms_per_step'),'e2e',(d.get('e2e') or {}).get('value'), 'roofline', (d.get('roofline') or {}).get('achieved'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f,'parse failed',e)
PY

#!/bin/bash
# N-GPU validation: sharded == single, then the scaling bench at N=1 and N=2 (weak scaling, 256 crops/GPU)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?"
grep dist_check gpurun_out/dist_check.log | tail -5
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "benchN rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n*.log')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, 'n_gpus',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])
    except Exception as e: print(f,'parse failed',e)
PY
tail -n 5 gpurun_out/bench_n$N.err

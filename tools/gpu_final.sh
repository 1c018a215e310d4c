#!/bin/bash
# end-of-round evidence: full gpu test-suite (verbose parity numbers), smoke, benches for every preset, ncu launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_verbose.log 2>&1; echo "pytest gpu rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --dump-ops gpurun_out/ops_b256.csv > gpurun_out/bench_cliff_w32.log 2> gpurun_out/bench_cliff_w32.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"
for P in "pare_w32 128" "cliff_w48cls 256" "pare_r50 256"; do set -- $P; timeout 600 python bench.py --preset $1 --batch $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.log 2> gpurun_out/bench_$1.err; echo "bench $1 rc=$?"; done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:poco -c 1500 --csv --log-file gpurun_out/launches_poco_b256.csv python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
grep -E "passed|failed" gpurun_out/t_gpu_verbose.log | tail -2

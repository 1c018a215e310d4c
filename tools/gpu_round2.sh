#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops_b256.csv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 6 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -n 3 gpurun_out/t_gpu.log; cat gpurun_out/bench.log

#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/mma_bench > gpurun_out/mma_bench.csv 2>&1; echo "mma_bench rc=$?"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 4 -o gpurun_out/prof_conv2 python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/mma_bench.csv

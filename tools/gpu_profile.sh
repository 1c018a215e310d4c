#!/bin/bash
# ncu evidence for profiles/: launch list of one forward + full captures of representative conv launches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_b256.csv python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 10 -o gpurun_out/prof_conv_stage2 python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_full_a.log 2>&1; echo "ncu full a rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 300 -c 25 -o gpurun_out/prof_conv_tail python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1; echo "ncu full b rc=$?"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
POCO_B200_LANES=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:^(conv_tc|pack_image|fuse_sum|upsample2x|avgpool|linear|copy2d|rot6d)_kernel" -s 352 -c 352 --csv --log-file gpurun_out/launches_poco_b256.csv python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
tail -c 600 gpurun_out/bench.log

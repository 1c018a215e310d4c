#!/bin/bash
# ncu launch list of the library's own kernels over the first forwards of bench.py (per-launch time + DRAM bytes)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc_kernel|fuse_sum_kernel|upsample2x_kernel|pack_|avgpool_kernel|linear_kernel|copy2d_kernel|rot6d_kernel" -c 720 --csv --log-file gpurun_out/r01c_launches_b256.csv python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/r01c_launches_b256.csv

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/t_ops.log 2>&1; rc=$?; echo "ops tests rc=$rc"; tail -n 4 gpurun_out/t_ops.log
if [ $rc -ne 0 ]; then exit 1; fi
bash tools/gpu_bench_variants.sh POCO_B200_HALF=0 POCO_B200_HALF=1 POCO_B200_HALF=2 POCO_B200_HALF=0 POCO_B200_HALF=1

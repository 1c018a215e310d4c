#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -q -x -k "linear or golden or e2e or preset or fresh" > gpurun_out/t_lin.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_lin.log
bash tools/gpu_bench_variants.sh "POCO_B200_LIB=$PWD/tools/bin/lib_ref.so" POCO_B200_HALF=1 "POCO_B200_LIB=$PWD/tools/bin/lib_ref.so" POCO_B200_HALF=1

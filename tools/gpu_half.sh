#!/bin/bash
mkdir -p gpurun_out
POCO_B200_HALF=3 timeout 200 python -m pytest tests/test_gpu_e2e.py -q -x -k "golden or preset or e2e" > gpurun_out/t_e2e_half3.log 2>&1; echo "e2e tests (HALF=3) rc=$?"; tail -n 3 gpurun_out/t_e2e_half3.log
bash tools/gpu_bench_variants.sh POCO_B200_HALF=1 POCO_B200_HALF=3 POCO_B200_HALF=1 POCO_B200_HALF=3

#!/bin/bash
# quick check after a host-side change: e2e parity tests + default bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?"; tail -n 2 gpurun_out/t_e2e.log
bash tools/gpu_bench_variants.sh POCO_B200_HALF=1 POCO_B200_HALF=1

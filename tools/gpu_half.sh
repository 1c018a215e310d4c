#!/bin/bash
mkdir -p gpurun_out
POCO_B200_TWOPHASE=1 timeout 200 python -m pytest tests/test_gpu_e2e.py -q -x > gpurun_out/t_e2e_2p.log 2>&1; echo "e2e tests (TWOPHASE=1) rc=$?"; tail -n 3 gpurun_out/t_e2e_2p.log
bash tools/gpu_bench_variants.sh POCO_B200_TWOPHASE=0 POCO_B200_TWOPHASE=1 POCO_B200_TWOPHASE=0 POCO_B200_TWOPHASE=1

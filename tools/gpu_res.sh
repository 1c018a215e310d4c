#!/bin/bash
mkdir -p gpurun_out
POCO_B200_RES_RING=0 timeout 600 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/t_ops_ldg.log 2>&1; rc=$?; echo "ops tests (register residual) rc=$rc"; tail -n 6 gpurun_out/t_ops_ldg.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/t_ops.log 2>&1; echo "ops tests (ring) rc=$?"; tail -n 3 gpurun_out/t_ops.log
for r in 4 0; do echo "== RES_RING $r"; POCO_B200_RES_RING=$r timeout 200 python tools/conv_bench.py 256 0 2>&1 | grep "res1" ; done
bash tools/gpu_bench_variants.sh POCO_B200_RES_RING=4 POCO_B200_RES_RING=0 POCO_B200_RES_RING=4 POCO_B200_RES_RING=0

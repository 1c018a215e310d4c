#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/t_ops.log 2>&1; rc=$?; echo "ops tests rc=$rc"; tail -n 8 gpurun_out/t_ops.log
if [ $rc -ne 0 ]; then exit 1; fi
for g in 1 2 4; do echo "== MGROUP $g"; POCO_B200_MGROUP=$g timeout 200 python tools/conv_bench.py 256 0,1 2>&1 | grep -E "32->32|64->64 k3|128->128" ; done

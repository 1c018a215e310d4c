#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -k "crop" > gpurun_out/t_crop.log 2>&1; echo "crop test rc=$?"; tail -n 12 gpurun_out/t_crop.log
timeout 100 python tools/crop_bench.py 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -q -x -k "crop or uncert" > gpurun_out/t_next.log 2>&1; echo "next-row tests rc=$?"; tail -n 12 gpurun_out/t_next.log

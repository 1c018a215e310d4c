#!/bin/bash
# f4 validation: SMPL stage GPU tests + stage throughput
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_smpl.py -q -m gpu > gpurun_out/t_smpl.log 2>&1; echo "smpl tests rc=$?"; tail -n 25 gpurun_out/t_smpl.log
timeout 60 python tools/smpl_bench.py 256 30 > gpurun_out/smpl_bench.log 2>&1; echo "smpl bench rc=$?"; tail -n 2 gpurun_out/smpl_bench.log

#!/bin/bash
# first GPU pass of round 2: split-precision tests, then the old suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py -x -q -m gpu -s > gpurun_out/t_split.log 2>&1
echo "split rc=$?" >> gpurun_out/t_split.log
tail -30 gpurun_out/t_split.log

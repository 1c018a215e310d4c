#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BR_PROF=1 timeout 120 python tools/branch_bench.py 256 4 2>&1 | tail -n 14
timeout 120 python tools/region_times.py 256 > gpurun_out/r02h_region_times_b256.txt 2>&1; tail -n 45 gpurun_out/r02h_region_times_b256.txt | cut -c1-260
POCO_B200_FUSE_BRANCH=0 timeout 120 python tools/region_times.py 256 > gpurun_out/r02h_region_times_b256_nobranch.txt 2>&1; grep -E "total|lanes  *(8|12|16) ops" gpurun_out/r02h_region_times_b256_nobranch.txt | cut -c1-100,200-300

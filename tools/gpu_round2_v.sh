#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for w in 1 2 1 2 4; do timeout 120 python tools/split_batch_bench.py base $w 256 2>/dev/null | tail -n 1; done
POCO_B200_SHARE_SCALE=1 timeout 120 python tools/split_batch_bench.py scale1 2 256 2>/dev/null | tail -n 1
POCO_B200_SHARE_SCALE=1.5 timeout 120 python tools/split_batch_bench.py scale1.5 2 256 2>/dev/null | tail -n 1
POCO_B200_LANES=0 timeout 120 python tools/split_batch_bench.py nolanes 2 256 2>/dev/null | tail -n 1
POCO_B200_LANES=0 timeout 120 python tools/split_batch_bench.py nolanes 4 256 2>/dev/null | tail -n 1

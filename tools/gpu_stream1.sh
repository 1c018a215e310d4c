#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu -k "stream" 2>&1 | tail -3
for D in 1 8 32; do
  timeout 200 python tools/stream_bench.py $D 200 --no-graph 2>/dev/null | tail -1 | tee -a gpurun_out/stream_1gpu.jsonl
  timeout 200 python tools/stream_bench.py $D 200 2>/dev/null | tail -1 | tee -a gpurun_out/stream_1gpu.jsonl
done

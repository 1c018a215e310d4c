#!/bin/bash
# BASELINE configs[4]: 8 concurrent 1080p streams, one per GPU (run with gpurun --gpus 8)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/stream8_gpus.txt
for D in 8 32; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/stream_bench.py $D 300 2>gpurun_out/stream8_err.log | grep -E '^\{|^stream' | tee -a gpurun_out/stream8.log
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/stream_bench.py 8 300 --no-graph --no-chains 2>>gpurun_out/stream8_err.log | grep -E '^\{|^stream' | tee -a gpurun_out/stream8.log
timeout 200 python tools/stream_bench.py 8 300 2>/dev/null | grep -E '^\{|^stream' | tee -a gpurun_out/stream8.log
timeout 200 python tools/stream_bench.py 1 300 2>/dev/null | grep -E '^\{|^stream' | tee -a gpurun_out/stream8.log
tail -3 gpurun_out/stream8_err.log

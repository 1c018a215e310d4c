#!/bin/bash
# 2-GPU sanity of the shipped build: the NCCL test the 1-GPU boxes skip + the bench line under torchrun as the driver launches it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_n2_devices.txt
timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -k "gpus or nccl or shard or dist" > gpurun_out/r02h_n2_tests.log 2>&1; echo "2-gpu tests rc=$?"; tail -n 3 gpurun_out/r02h_n2_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r02h_bench_n2.json | cut -c1-400
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err; echo "reference arm rc=$?"; cut -c1-500 gpurun_out/r02h_bench_reference.json

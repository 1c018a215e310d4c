#!/bin/bash
# BASELINE configs[4]: one 1080p stream per GPU (run under `gpurun --gpus N -- 'NGPU=N bash tools/gpu_streams.sh'`)
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 300 python tools/stream_bench.py 8 200 > gpurun_out/stream_n1.log 2>&1; echo "stream n1 rc=$?"; tail -n 1 gpurun_out/stream_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/stream_bench.py 8 200 > gpurun_out/stream_n$N.log 2>&1; echo "stream n$N rc=$?"; tail -n 1 gpurun_out/stream_n$N.log
timeout 300 python tools/stream_bench.py 8 200 --smpl > gpurun_out/stream_n1_smpl.log 2>&1; echo "stream smpl rc=$?"; tail -n 1 gpurun_out/stream_n1_smpl.log

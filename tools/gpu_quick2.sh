#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_e2e.py -q -x -k "stream" > gpurun_out/t_stream.log 2>&1; echo "stream test rc=$?"; tail -n 12 gpurun_out/t_stream.log
timeout 200 python tools/stream_bench.py 8 200 2>&1 | tail -1
timeout 200 python tools/stream_bench.py 32 100 2>&1 | tail -1

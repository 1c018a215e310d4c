#!/bin/bash
# session-2 first check: MMA issue microbench with rotating accumulators, GPU tests, bench
mkdir -p gpurun_out
timeout 120 tools/bin/mma_bench2 > gpurun_out/mma_bench2.csv 2>&1; echo "mma_bench2 rc=$?"
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops_b256.csv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/t_gpu.log; tail -n 1 gpurun_out/bench.log; cat gpurun_out/mma_bench2.csv

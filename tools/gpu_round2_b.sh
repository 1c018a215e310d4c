#!/bin/bash
# round 2, pass b: whole GPU suite + smoke + bench (both precision modes) + per-op table
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_gpu.log
tail -5 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -12 gpurun_out/smoke.log
timeout 600 python bench.py --dump-ops gpurun_out/ops_b256.csv > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log
timeout 600 python bench.py --precision split --no-other-mode --no-cpu-baseline --dump-ops gpurun_out/ops_b256_split.csv > gpurun_out/bench_split.log 2> gpurun_out/bench_split.err; echo "bench split rc=$?"
cat gpurun_out/bench_split.log

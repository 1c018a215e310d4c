#!/bin/bash
# conv correctness (ops + chains + split) then issuer cycle accounting and slopes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/t_conv.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_conv.log
export POCO_B200_HALF=0
timeout 100 python tools/conv_prof.py 64 64 3 1 28 0 1024 0,5,7 2>/dev/null
timeout 100 python tools/conv_prof.py 32 32 3 1 56 0 512 0,5,7 2>/dev/null
unset POCO_B200_HALF
out=gpurun_out/slope2.csv
echo "tag,case,B,debug,best_us" > $out
run() { tag=$1; shift; envs=$1; shift; env $envs timeout 200 python tools/conv_slope.py $tag "$@" >> $out 2>>gpurun_out/slope_err.log || echo "$tag FAILED" >> $out; }
run def X=1 32 32 3 1 56 0 0,5 256,512
run def X=1 32 32 3 1 56 1 0 256,512
run def X=1 64 64 3 1 28 0 0,5 256,512
run def X=1 64 64 3 1 28 1 0 256,512
run def X=1 128 128 3 1 14 1 0,5 256,512
run def X=1 256 256 3 1 7 1 0 256,512
run def X=1 32 64 3 2 56 1 0 256,512
cat $out
timeout 300 python tools/split_batch_bench.py e2e 1 2>/dev/null

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
out=gpurun_out/slope.csv
echo "tag,case,B,debug,best_us" > $out
run() { tag=$1; shift; envs=$1; shift; env $envs timeout 200 python tools/conv_slope.py $tag "$@" >> $out 2>>gpurun_out/slope_err.log || echo "$tag FAILED" >> $out; }
run half1 X=1 32 32 3 1 56 0 0,1,2,3,5,7 64,128,256,512
run half0 POCO_B200_HALF=0 32 32 3 1 56 0 0,1,2,3,5,7 64,128,256,512
run half1 X=1 64 64 3 1 28 0 0,1,2,3,5,7 128,256,512,1024
run half0 POCO_B200_HALF=0 64 64 3 1 28 0 0,1,2,3,5,7 128,256,512,1024
run def X=1 128 128 3 1 14 1 0,1,2,3,5,7 256,512,1024,2048
run def X=1 256 256 3 1 7 1 0,1,2,3,5,7 256,512,1024,2048
cat $out

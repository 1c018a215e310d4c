#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
out=gpurun_out/batch_ab.csv
echo "tag,case,B,debug,best_us" > $out
run() { tag=$1; shift; envs=$1; shift; env $envs timeout 200 python tools/conv_slope.py $tag "$@" >> $out 2>>gpurun_out/batch_ab_err.log || echo "$tag FAILED" >> $out; }
for b in 0 12 36; do
  L=POCO_B200_LIB=$PWD/tools/bin/lib_batch$b.so
  run b$b "$L" 32 32 3 1 56 0 0,5 256,512
  run b$b "$L" 32 32 3 1 56 1 0 256,512
  run b$b "$L POCO_B200_HALF=0" 64 64 3 1 28 0 0,5 256,512
  run b${b}h "$L" 64 64 3 1 28 1 0,5 256,512
  run b$b "$L" 128 128 3 1 14 1 0,5 256,512
  run b$b "$L" 256 256 3 1 7 1 0 256,512
  run b$b "$L" 256 256 3 1 56 0 0 256
  run b$b "$L" 64 256 1 1 56 1 0 256
  run b$b "$L" 32 64 3 2 56 1 0 256
done
cat $out
for b in 0 12 36; do
  POCO_B200_LIB=$PWD/tools/bin/lib_batch$b.so timeout 300 python tools/split_batch_bench.py lib_batch$b 1 2>>gpurun_out/batch_ab_err.log | tee -a gpurun_out/batch_ab_e2e.csv
done
tail -3 gpurun_out/batch_ab_err.log

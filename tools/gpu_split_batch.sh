#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
out=gpurun_out/split_batch.csv
echo "tag,preset,B,ways,ms_per_step,crops_per_s" > $out
run() { tag=$1; shift; envs=$1; shift; env $envs timeout 200 python tools/split_batch_bench.py $tag "$@" >> $out 2>>gpurun_out/split_batch_err.log || echo "$tag FAILED" >> $out; }
run base X=1 1
run base X=1 2
run base X=1 4
run half2 POCO_B200_HALF=2 1
run half2 POCO_B200_HALF=2 2
run half2 POCO_B200_HALF=2 4
run nolanes POCO_B200_LANES=0 1
run nolanes POCO_B200_LANES=0 2
run nolanes POCO_B200_LANES=0 4
run nolanes_half2 "POCO_B200_LANES=0 POCO_B200_HALF=2" 2
run nolanes_half2 "POCO_B200_LANES=0 POCO_B200_HALF=2" 4
run pdl POCO_B200_PDL=1 1
run pdl POCO_B200_PDL=1 2
cat $out
tail -5 gpurun_out/split_batch_err.log

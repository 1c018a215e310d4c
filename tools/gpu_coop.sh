#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
out=gpurun_out/coop.csv
echo "tag,preset,B,ways,ms_per_step,crops_per_s" > $out
run() { tag=$1; shift; envs=$1; shift; env $envs timeout 200 python tools/split_batch_bench.py $tag "$@" >> $out 2>>gpurun_out/coop_err.log || echo "$tag FAILED" >> $out; }
run base X=1 1
run coop POCO_B200_COOP=1 1
run coop_nolanes "POCO_B200_COOP=1 POCO_B200_LANES=0" 1
run nolanes "POCO_B200_LANES=0" 1
run coop_2way POCO_B200_COOP=1 2
run pdl_only POCO_B200_PDL=1 1
cat $out
tail -3 gpurun_out/coop_err.log
POCO_B200_COOP=1 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_tcgen05 or chain" 2>&1 | tail -2
POCO_B200_COOP=1 timeout 100 python tools/conv_timeline.py 64 64 3 1 28 1 256 2>/dev/null | tail -6

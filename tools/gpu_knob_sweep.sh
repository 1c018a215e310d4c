#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
out=gpurun_out/knob_sweep.csv
echo "tag,case,debug,us" > $out
run() { tag=$1; shift; env "$@" timeout 120 python tools/conv_knobs.py $tag ${DBGS:-0} >> $out 2>gpurun_out/knob_err.log || echo "$tag FAILED" >> $out; }
DBGS=0,1,2,3,5 run old POCO_B200_RING_ALWAYS=1 POCO_B200_MAX_STAGES=8
DBGS=0,1,2,3,5 run new X=1
run new_g1 POCO_B200_MGROUP=1
run new_g4 POCO_B200_MGROUP=4
run new_rr2 POCO_B200_RES_RING=2
run new_rr8 POCO_B200_RES_RING=8
run new_half0 POCO_B200_HALF=0
run new_st4 POCO_B200_MAX_STAGES=4
run new_st8 POCO_B200_MAX_STAGES=8
run new_half0_g1 POCO_B200_HALF=0 POCO_B200_MGROUP=1
cat $out

#!/bin/bash
mkdir -p gpurun_out
for v in _ref _cg; do export POCO_B200_LIB=$PWD/tools/bin/lib$v.so; echo "== $v"; timeout 100 python tools/s2_bench.py 2>&1 | grep -v case; done
unset POCO_B200_LIB
bash tools/gpu_bench_variants.sh POCO_B200_RES_RING=4 "POCO_B200_LIB=$PWD/tools/bin/lib_cg.so" POCO_B200_RES_RING=2 POCO_B200_RES_RING=8 POCO_B200_RES_RING=4 "POCO_B200_LIB=$PWD/tools/bin/lib_cg.so"

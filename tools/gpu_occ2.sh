#!/bin/bash
mkdir -p gpurun_out
for v in _ref _occ2; do
  export POCO_B200_LIB=$PWD/tools/bin/lib$v.so
  echo "== variant $v"
  timeout 200 python tools/conv_bench.py 256 0 2>&1 | grep -v case
done
export POCO_B200_LIB=$PWD/tools/bin/lib_occ2.so
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "tcgen05 or chain" 2>&1 | tail -2
unset POCO_B200_LIB
bash tools/gpu_bench_variants.sh POCO_B200_PDL=0 "POCO_B200_PDL=0 POCO_B200_LIB=$PWD/tools/bin/lib_occ2.so" "POCO_B200_PDL=1 POCO_B200_LIB=$PWD/tools/bin/lib_occ2.so" POCO_B200_PDL=0

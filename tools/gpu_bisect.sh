#!/bin/bash
mkdir -p gpurun_out
for v in _fdc2ccc HEAD HEADRING HEADG1; do
  unset POCO_B200_LIB POCO_B200_RES_RING POCO_B200_MGROUP
  case $v in
    HEAD) ;;
    HEADRING) export POCO_B200_RES_RING=4 ;;
    HEADG1) export POCO_B200_RES_RING=4 POCO_B200_MGROUP=1 ;;
    *) export POCO_B200_LIB=$PWD/tools/bin/lib$v.so ;;
  esac
  echo "== variant $v"
  timeout 200 python tools/conv_bench.py 256 0 2>&1 | grep -E "res1|res0"
done
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
for v in _fdc2ccc HEADRING; do
  unset POCO_B200_LIB POCO_B200_RES_RING POCO_B200_MGROUP
  case $v in
    HEADRING) export POCO_B200_RES_RING=4 ;;
    *) export POCO_B200_LIB=$PWD/tools/bin/lib$v.so ;;
  esac
  echo "== variant $v"
  timeout 200 python tools/conv_bench.py 256 0 2>&1 | grep -E "res1"
done

#!/bin/bash
mkdir -p gpurun_out
for v in A B A B; do
  if [ $v = B ]; then unset POCO_B200_LIB; else export POCO_B200_LIB=$PWD/tools/bin/lib$v.so; fi
  echo "== variant $v"
  timeout 200 python tools/conv_bench.py 256 0 2>&1 | tee gpurun_out/ab_$v.csv
done

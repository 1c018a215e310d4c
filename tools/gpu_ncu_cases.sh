#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -o gpurun_out/r01b_conv_cases python tools/prof_cases.py 256 > gpurun_out/ncu_cases.log 2>&1; echo "ncu rc=$?"
tail -n 5 gpurun_out/ncu_cases.log; ls -la gpurun_out/*.ncu-rep

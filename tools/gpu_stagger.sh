#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for st in 0 600 1200 2000 4000; do
  echo "== stagger $st"
  POCO_B200_STAGGER=$st POCO_B200_HALF=0 timeout 100 python tools/conv_prof.py 64 64 3 1 28 0 1024 0,5 2>/dev/null | tail -4
  POCO_B200_STAGGER=$st POCO_B200_HALF=0 timeout 100 python tools/conv_prof.py 32 32 3 1 56 0 512 0 2>/dev/null | tail -2
done

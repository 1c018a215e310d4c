#!/bin/bash
# one GPU-box session: op tests, conv diagnostics, end-to-end parity, smoke, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 240 python tools/diag_conv.py linear > gpurun_out/diag_linear.log 2>&1; echo "diag linear rc=$?"
timeout 240 python tools/diag_conv.py gather > gpurun_out/diag_gather.log 2>&1; echo "diag gather rc=$?"
timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "not tcgen05 and not identity and not invariance" > gpurun_out/t_ops_simple.log 2>&1; echo "ops simple rc=$?"
timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tcgen05 or identity or invariance" > gpurun_out/t_conv_tc.log 2>&1; echo "conv tc rc=$?"
POCO_B200_CONV_IMPL=1 POCO_B200_GRAPH=0 timeout 400 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s -k "goldens" > gpurun_out/t_e2e_debugconv.log 2>&1; echo "e2e debugconv rc=$?"
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -s > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/diag_linear.log gpurun_out/diag_gather.log gpurun_out/t_ops_simple.log gpurun_out/t_conv_tc.log gpurun_out/t_e2e_debugconv.log gpurun_out/t_e2e.log gpurun_out/smoke.log gpurun_out/bench.log

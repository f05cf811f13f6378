#!/bin/bash
# Runs on the B200 box under gpurun: kernel unit tests first (short timeouts so a hung
# tcgen05 pipeline cannot wedge the box), then the parity suite.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== gemm"; timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/t_gemm.log
echo "== update"; timeout 1200 python -m pytest tests/test_gpu_update.py -q -m gpu 2>&1 | tail -60 | tee gpurun_out/t_update.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 500 --warmup 20 2>&1 | tail -5 | tee gpurun_out/bench.log

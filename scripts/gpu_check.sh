#!/bin/bash
# Runs on the B200 box under gpurun: kernel unit tests first (short timeouts so a hung
# tcgen05 pipeline cannot wedge the box), then the parity suite.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== simt gemm" ; timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "simt" 2>&1 | tail -15 | tee gpurun_out/t_gemm_simt.log
echo "== tcgen05 gemm"; timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "tcgen05 or exact" 2>&1 | tail -40 | tee gpurun_out/t_gemm_tc.log
echo "== update simt"; timeout 900 python -m pytest tests/test_gpu_update.py -q -m gpu -k "simt or ring or sampler" 2>&1 | tail -40 | tee gpurun_out/t_update_simt.log
echo "== update all"; timeout 1200 python -m pytest tests/test_gpu_update.py -q -m gpu -k "not simt and not ring and not sampler" 2>&1 | tail -40 | tee gpurun_out/t_update_tc.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_update.py tests/test_gpu_configs.py tests/test_golden.py -x -q -m gpu > gpurun_out/r02k_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02k_tests.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_SIDE_DELAY": 0, "DQNB_BN_SIDE_L1": 64}' '{}' > gpurun_out/r02k_sweep.txt 2>&1
cat gpurun_out/r02k_sweep.txt
python scripts/trace_update.py 1024 > gpurun_out/r02k_trace.txt 2>&1

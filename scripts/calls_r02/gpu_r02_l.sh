#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_update.py tests/test_gpu_configs.py tests/test_golden.py -x -q -m gpu > gpurun_out/r02l_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02l_tests.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_BN32_MAX_TILES": 0}' '{"DQNB_BN32_MAX_TILES": 16}' '{"DQNB_BN32_MAX_TILES": 64}' '{"DQNB_BN32_CLUSTER": 1}' '{"DQNB_BN32_MAX_TILES": 64, "DQNB_BN32_CLUSTER": 1}' '{}' > gpurun_out/r02l_sweep.txt 2>&1
cat gpurun_out/r02l_sweep.txt
python scripts/trace_update.py 1024 > gpurun_out/r02l_trace.txt 2>&1

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/r02n_tests_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -5 gpurun_out/r02n_tests_gemm.log
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_configs.py tests/test_golden.py -x -q -m gpu > gpurun_out/r02n_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02n_tests.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_BN32_MAX_TILES": 0}' '{"DQNB_BN_FWD": 128}' '{"DQNB_BN_DX": 128}' '{"DQNB_BN_FWD_SIDE": 128}' '{}' > gpurun_out/r02n_sweep.txt 2>&1
cat gpurun_out/r02n_sweep.txt
python scripts/trace_update.py 1024 > gpurun_out/r02n_trace.txt 2>&1
python scripts/perf_gemm_ts.py > gpurun_out/r02n_probe.txt 2>&1; cat gpurun_out/r02n_probe.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_configs.py tests/test_golden.py -x -q -m gpu > gpurun_out/r02h_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02h_tests.log
DQNB_SIDE_DELAY=1 timeout 600 python -m pytest tests/test_gpu_update.py -x -q -m gpu > gpurun_out/r02h_tests_sd.log 2>&1
echo "tests(side_delay) rc=$?"; tail -3 gpurun_out/r02h_tests_sd.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_STORE_WAIT_FULL": 1}' '{"DQNB_SIDE_DELAY": 1}' '{}' '{"DQNB_STORE_WAIT_FULL": 1}' '{"DQNB_SIDE_DELAY": 1}' '{"DQNB_SIDE_DELAY": 1, "DQNB_CLUSTER_B": 1}' > gpurun_out/r02h_sweep.txt 2>&1
cat gpurun_out/r02h_sweep.txt
python scripts/trace_update.py 1024 '{"DQNB_SIDE_DELAY": 1}' > gpurun_out/r02h_trace_side_delay.txt 2>&1
python scripts/trace_update.py 1024 > gpurun_out/r02h_trace.txt 2>&1

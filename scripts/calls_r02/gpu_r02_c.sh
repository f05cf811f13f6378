#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_async.py tests/test_golden.py -x -q -m gpu > gpurun_out/r02c_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02c_tests.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_GATHER_AHEAD": 0}' '{"DQNB_BN_BIG": 128}' '{"DQNB_BN_DX": 128}' '{"DQNB_BN_DW": 128}' '{"DQNB_BN_DX": 128, "DQNB_BN_DW": 128}' '{"DQNB_ST_DX": 2}' '{"DQNB_ST_DX": 3}' '{"DQNB_ST_DW": 3}' '{"DQNB_CLUSTER_B": 1}' '{"DQNB_BN_FWD": 128}' '{"DQNB_FUSE_COLSUM": 0}' '{}' > gpurun_out/r02c_sweep.txt 2>&1
cat gpurun_out/r02c_sweep.txt
python scripts/trace_update.py 1024 '{"DQNB_BN_BIG": 128}' > gpurun_out/r02c_trace_bnbig.txt 2>&1
python scripts/trace_update.py 1024 > gpurun_out/r02c_trace.txt 2>&1

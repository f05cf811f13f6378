#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_multi.py tests/test_golden.py -x -q -m gpu > gpurun_out/r02d_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02d_tests.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_BN_DW": 64}' '{"DQNB_FUSE_COLSUM": 0}' '{"DQNB_FUSE_TL": 0}' '{"DQNB_ACTOR_LATE": 1}' '{"DQNB_ACTOR_LATE": 1, "DQNB_FUSE_COLSUM": 0}' '{"DQNB_FUSE_COLSUM": 0, "DQNB_ST_DW": 3}' '{"DQNB_ST_DW": 3}' '{}' > gpurun_out/r02d_sweep.txt 2>&1
cat gpurun_out/r02d_sweep.txt
python scripts/trace_update.py 1024 '{"DQNB_ACTOR_LATE": 1}' > gpurun_out/r02d_trace_actor_late.txt 2>&1

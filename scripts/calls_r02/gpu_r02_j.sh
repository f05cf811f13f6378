#!/bin/bash
mkdir -p gpurun_out
export DQNB_SIDE_DELAY=1 DQNB_BN_SIDE_L1=128
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_DW_AFTER_DX": 2}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_DW_AFTER_DX": 6}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_DW_AFTER_DX": 4}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_DW_AFTER_DX": 6, "DQNB_ST_DW": 3}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_DW_AFTER_DX": 2, "DQNB_BN_DW": 64}' > gpurun_out/r02j_sweep.txt 2>&1
cat gpurun_out/r02j_sweep.txt
python scripts/trace_update.py 1024 '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_DW_AFTER_DX": 6}' > gpurun_out/r02j_trace_dw6.txt 2>&1

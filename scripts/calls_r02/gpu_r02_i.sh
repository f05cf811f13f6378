#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_SIDE_DELAY": 1}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128}' '{"DQNB_BN_SIDE_L1": 128}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_CLUSTER_B": 1}' '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_BN_FWD_SIDE": 128}'  '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_BN_FWD_SIDE": 128, "DQNB_CLUSTER_B": 1}' '{}' > gpurun_out/r02i_sweep.txt 2>&1
cat gpurun_out/r02i_sweep.txt
python scripts/trace_update.py 1024 '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128}' > gpurun_out/r02i_trace_sd_l1.txt 2>&1
python scripts/trace_update.py 1024 '{"DQNB_SIDE_DELAY": 1, "DQNB_BN_SIDE_L1": 128, "DQNB_BN_FWD_SIDE": 128, "DQNB_CLUSTER_B": 1}' > gpurun_out/r02i_trace_sd_all.txt 2>&1

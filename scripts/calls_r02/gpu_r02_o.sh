#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sweep_sched.py 1024 '{"DQNB_BN_DW": 64}' '{"DQNB_BN_DW": 64, "DQNB_ST_DW": 3}' '{"DQNB_ST_DW": 3}' '{"DQNB_DW_AFTER_DX": 2}' '{"DQNB_DW_AFTER_DX": 4}' '{"DQNB_BN_DW": 64, "DQNB_DW_AFTER_DX": 2}' '{"DQNB_CLUSTER_B": 1}' '{"DQNB_FUSE_COLSUM": 0}' '{"DQNB_GATHER_AHEAD": 0}' '{"DQNB_ST_FWD": 3}' '{"DQNB_ST_DX": 3}' '{}' > gpurun_out/r02o_sweep.txt 2>&1
cat gpurun_out/r02o_sweep.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_act.py -x -q -m gpu > gpurun_out/r02f_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r02f_tests.log
timeout 300 python scripts/act_probe.py > gpurun_out/r02f_probe.log 2>&1; cat gpurun_out/r02f_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'act_' -c 12 --csv --log-file gpurun_out/r02f_act_ncu.csv python scripts/act_probe.py 3 > gpurun_out/r02f_probe_ncu.log 2>&1
grep -E "act_" gpurun_out/r02f_act_ncu.csv | head -12 | awk -F'","' '{print $5, $9, $NF}'

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02f3_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02f3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f3_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python scripts/sweep_sched.py 1024 > gpurun_out/r02f3_sweep.txt 2>&1; cat gpurun_out/r02f3_sweep.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02f2_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r02f2_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02f2_smoke.log

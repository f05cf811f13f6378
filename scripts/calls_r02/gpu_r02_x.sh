#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02x_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r02x_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02x_smoke.log

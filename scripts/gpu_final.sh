#!/bin/bash
# Round-end validation on one B200: full GPU test suite, smoke, bench (both arms), ncu passes.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/t_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 100 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-220
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-220
bash scripts/gpu_profile.sh

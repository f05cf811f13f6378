#!/bin/bash
# final 1-GPU pass of the round: full GPU suite, smoke, both bench arms with the driver's default flags
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02f1_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r02f1_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f1_smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/r02f1_bench_n1.json 2> gpurun_out/r02f1_bench_n1.err
echo "bench rc=$?"; tail -4 gpurun_out/r02f1_bench_n1.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02f1_bench_reference.json 2> gpurun_out/r02f1_bench_reference.err
echo "reference rc=$?"; tail -4 gpurun_out/r02f1_bench_reference.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f1_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "windows_ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["over_device_value"])
print({k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_us", "achieved_over_step", "frac_of_3xtf32_ceiling", "frac_of_3xtf32_ceiling_over_step", "traffic")})
r = json.loads(open("gpurun_out/r02f1_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["cpu_baseline"]["cores"])
PY

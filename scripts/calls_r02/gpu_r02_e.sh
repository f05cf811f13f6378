#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_act.py tests/test_gpu_update.py tests/test_gpu_rollout.py tests/test_gpu_async.py -x -q -m gpu > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r02e_tests.log
python - <<'PY' > gpurun_out/r02e_act_latency.txt 2>&1
import time, numpy as np, sys
sys.path.insert(0, '.')
from __graft_entry__ import load_package
from bench import synth_replay
P = load_package()
d = P.DQNB(state_size=58, batch=1024, hidden=(1024, 512, 256, 128), replay_capacity=70000, max_act_batch=128)
d.init_params(2, 0.01)
s, a, r, mc, term, sn = synth_replay(65536, 58, 1)
d.add_transitions(s, a, r, mc, sn, term)
d.update(5)
for busy in (0, 1):
    for n in (1, 8, 64, 128):
        x = np.ascontiguousarray(s[:n])
        for _ in range(20): d.select_actions(x)
        if busy: last = d.update_async(400)
        ts = []
        for _ in range(300):
            t0 = time.perf_counter(); d.select_actions(x); ts.append(time.perf_counter() - t0)
        if busy: d.results(last, 1)
        ts = np.array(ts) * 1e6
        print(f"busy={busy} n={n:4d}: median {np.median(ts):7.1f} us  p10 {np.percentile(ts,10):7.1f}  p90 {np.percentile(ts,90):7.1f}")
ms = d.benchmark(500)
print("update us:", ms / 500 * 1e3)
d.close()
PY
cat gpurun_out/r02e_act_latency.txt

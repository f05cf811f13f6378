#!/usr/bin/env python
"""bench.py — transition-updates/sec of the replay-minibatch actor-critic update (BASELINE.json metric).

One "step" = one complete UpdateActorCritic (reference src/dqn.cpp:828-972: sample, gather, 5 forward
and 3 backward passes, two clipped Adam steps, soft target update) over a minibatch of B transitions
per GPU drawn from an HBM-resident replay ring of synthetic 58-dim transitions.

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one process per GPU)
  python bench.py --impl reference [...]                     the CPU oracle port on the host cores

`value`   : device-timed (CUDA events on the handle's stream), replay already resident in HBM.
`e2e`     : the same metric through the C-ABI with HOST buffers every step: dqnb_add_transitions of B
            fresh rows (pinned staging -> H2D on the copy stream) + dqnb_update_async (1 update) + the
            previous step's (critic_loss, avg_q) read from the host-mapped result ring (dqnb_results);
            `blocking_value` is the same loop with the blocking dqnb_update.
`roofline`: the dominant kernel (gemm_tc_kernel, tcgen05 3xTF32) timed live with CUDA events.
`cpu_baseline`: the oracle port (Caffe operation order, OpenBLAS sgemm when loadable) on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_TRANSITION = {  # algorithmic minimum, BASELINE.md §2 / SURVEY §8a
    (58, (1024, 512, 256, 128)): 14_708_224,
    (77, (1024, 512, 256, 128)): 14_980_608,
    (58, (1024, 1024, 1024, 1024)): 63_942_656,
}


def flop_per_transition(S, hidden):
    key = (S, tuple(hidden))
    if key in FLOP_PER_TRANSITION:
        return FLOP_PER_TRANSITION[key]
    # 5 forward passes, 2 weight-gradient backward passes (no layer-1 dX), 1 dX-only pass
    dims_a = [S] + list(hidden)
    dims_c = [S + 10] + list(hidden)
    mac = lambda d, head: sum(d[i] * d[i + 1] for i in range(len(d) - 1)) + d[-1] * head
    fa, fc = mac(dims_a, 10), mac(dims_c, 1)
    fwd = 2 * (2 * fa + 3 * fc)
    bwd_w = lambda d, head: 2 * (2 * mac(d, head) - d[0] * d[1])
    dx_only = 2 * (mac(dims_c, 1) - dims_c[0] * dims_c[1] + 10 * dims_c[1])
    return fwd + bwd_w(dims_a, 10) + bwd_w(dims_c, 1) + dx_only


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synth_replay(n, S, seed):
    """SURVEY §8(d) generator: states U(-1,1), actions per GetRandomActorOutput ranges
    (dqn.cpp:664-682), ~100-step episodes, rewards N(0,0.1) with +5 on 1% of terminals,
    Monte-Carlo returns by the LabelTransitions rule (dqn.cpp:783-797)."""
    rng = np.random.default_rng(seed)
    s = rng.uniform(-1, 1, (n, S)).astype(np.float32)
    sn = np.empty_like(s)
    sn[:-1] = s[1:]
    sn[-1] = rng.uniform(-1, 1, S)
    a = np.empty((n, 10), np.float32)
    a[:, 0:4] = rng.uniform(-1, 1, (n, 4))
    a[:, 4] = rng.uniform(-100, 100, n)
    a[:, 5:8] = rng.uniform(-180, 180, (n, 3))
    a[:, 8] = rng.uniform(0, 100, n)
    a[:, 9] = rng.uniform(-180, 180, n)
    term = (rng.uniform(size=n) < 0.01).astype(np.uint8)
    term[-1] = 1
    r = rng.normal(0, 0.1, n).astype(np.float32)
    r[(term == 1) & (rng.uniform(size=n) < 0.01)] += 5.0
    mc = np.empty(n, np.float32)
    g = 0.0
    for i in range(n - 1, -1, -1):   # vectorising this is not worth it: runs once, outside timing
        g = float(r[i]) if term[i] else float(r[i]) + 0.99 * g
        mc[i] = g
    return s, a, r, mc, term, sn


def cpu_oracle_time(S, B, hidden, n_updates, seed=2, use_blas=True, threads=None):
    """Times n_updates oracle updates (the reference's Benchmark protocol, dqn.cpp:487-498)."""
    from oracle import oracle as O
    O.build()
    blas = bool(use_blas and O.load_blas())
    # torchrun exports OMP_NUM_THREADS=1, which OpenBLAS honours at load time: ask for every host core explicitly
    O.lib().dqo_set_threads(int(threads or os.cpu_count() or 1))
    ocfg = O.make_config(state_size=S, batch=B, hidden=hidden, caffe_wasted_work=1, use_blas=1 if blas else 0)
    rng = np.random.default_rng(seed)
    a0, c0 = O.init_params(ocfg, False, rng, "caffe"), O.init_params(ocfg, True, rng, "caffe")
    st = O.OracleState(ocfg, a0, c0, a0, c0)
    batch = O.synth_batch(ocfg, rng, p_term=0.01)
    st.update(*batch)  # warm-up (page-in, BLAS thread pool)
    t0 = time.perf_counter()
    for _ in range(n_updates):
        st.update(*batch)
    dt = time.perf_counter() - t0
    cores = O.lib().dqo_get_threads() if hasattr(O.lib(), "dqo_get_threads") else os.cpu_count()
    return dt / n_updates, blas, int(cores)


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port: the reference needs Caffe,
    which is not installable here) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, hidden = args.state_size, tuple(args.hidden)
    B = args.batch
    probe, blas, cores = cpu_oracle_time(S, B, hidden, 1)
    from oracle import oracle as O
    O.lib().dqo_set_threads(int(os.cpu_count() or 1))
    # bound the whole run to ~2 minutes: shrink the per-step sample if needed
    budget = 120.0
    total_steps = args.steps + args.warmup
    Bs = B
    if probe * total_steps > budget:
        Bs = max(32, int(B * budget / (probe * total_steps)) // 32 * 32)
    from oracle import oracle as O
    ocfg = O.make_config(state_size=S, batch=Bs, hidden=hidden, caffe_wasted_work=1, use_blas=1 if blas else 0)
    rng = np.random.default_rng(2)
    a0, c0 = O.init_params(ocfg, False, rng, "caffe"), O.init_params(ocfg, True, rng, "caffe")
    st = O.OracleState(ocfg, a0, c0, a0, c0)
    batch = O.synth_batch(ocfg, rng, p_term=0.01)
    for _ in range(args.warmup):
        st.update(*batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.update(*batch)
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    sample = (f"{args.steps} timed UpdateActorCritic steps of {Bs} transitions each (workload batch {B}); "
              f"oracle port in Caffe operation order incl. force_backward work, "
              f"{'OpenBLAS cblas_sgemm' if blas else 'portable C sgemm'}, {cores} threads")
    line = {
        "impl": "reference", "metric": "transition-updates/sec", "value": val, "unit": "transitions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same config block as the native arm prints for these flags (the driver compares the two); how this arm runs
        # the workload - fp32 on the host cores, no GPUs - is said in `implementation`
        "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "implementation": {"precision": "fp32 on the host cores (CPU port of the reference's Caffe operation order)",
                           "parallelism": f"{cores} host threads, rank 0 only"},
        "cpu_baseline": {"value": val, "unit": "transitions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    """The workload only (identical on both arms for the same flags; the driver compares the two blocks)."""
    return {
        "workload": f"cfg2: synthetic {args.state_size}-dim replay, {args.replay} stored transitions, "
                    f"batch {args.batch} UpdateActorCritic per GPU",
        "state_size": args.state_size, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
        "hidden": list(args.hidden), "replay_transitions": args.replay, "ranks": world,
        "l2": "replay ring (0.6 GB) exceeds L2; weights/activations are L2-resident by design of the workload",
    }


def implementation_note(args, world):
    """How THIS arm runs the workload (not part of `config`)."""
    return {
        "parallelism": (f"dp{world} (replay sharded; gradient exchange: " +
                        ("fused reduce-scatter/all-gather kernel over NVLink peer memory)" if getattr(args, "comm", "p2p") == "p2p"
                         else "ncclAllReduce)")) if world > 1 else "single GPU",
        "precision": "3xTF32 split-fp32 operands on tcgen05 (fp32-faithful: parity mode == benchmarked mode)",
    }


def measure_tensor_peaks(torch):
    """Dense TF32 and bf16 matmul throughput of this GPU through torch (cuBLAS), 8192^3, a short sustained loop:
    the denominators of `frac_of_3xtf32_ceiling` (BASELINE.md section 3 asks for the TF32 peak to be measured)."""
    out = {}
    torch.backends.cuda.matmul.allow_tf32 = True
    for name, dt in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
        a = torch.randn(8192, 8192, device="cuda", dtype=dt)
        b = torch.randn(8192, 8192, device="cuda", dtype=dt)
        for _ in range(3):
            a @ b
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 60 if name == "tf32" else 120
        e0.record()
        for _ in range(n):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        out[name + "_tflops_sustained"] = 2 * 8192 ** 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b
    return out


def act_latency(d, states, busy_updates=0):
    """Wall-clock microseconds of one SelectActions call through the C-ABI with host buffers (median of 200 calls);
    busy_updates > 0: while that many asynchronous updates are in flight on the learner's stream."""
    out = {}
    for n in (1, 8, 64):
        x = np.ascontiguousarray(states[:n])
        for _ in range(20):
            d.select_actions(x)
        last = d.update_async(busy_updates) if busy_updates else 0
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            d.select_actions(x)
            ts.append(time.perf_counter() - t0)
        if busy_updates:
            d.results(last, 1)
        out[str(n)] = round(float(np.median(ts)) * 1e6, 2)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--state-size", dest="state_size", type=int, default=58)
    ap.add_argument("--hidden", type=int, nargs="+", default=[1024, 512, 256, 128])
    ap.add_argument("--replay", type=int, default=1_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / wide-MLP / act-path / peak legs (profiling runs)")
    ap.add_argument("--windows", type=int, default=3, help="timed windows of --steps updates each; the median is reported")
    ap.add_argument("--gemm-mode", type=int, default=0)
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="gradient exchange for --gpus > 1")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch  # device plumbing + rendezvous only
    from __graft_entry__ import load_package
    pkg = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")      # host-side waits that must not spin on the GPU
    # NVML is initialised here, long before the timed region: nvmlInit() inside it skewed the ranks' start (round 1)
    sampler = ClockSampler(local)
    S, B, hidden = args.state_size, args.batch, tuple(args.hidden)
    from scripts import dp_parity

    def make_learner(hid, rows, seed_off=0):
        d = pkg.DQNB(device=local, state_size=S, batch=B, hidden=hid, replay_capacity=rows + B + 8,
                     seed=3 + rank + seed_off, world_size=world, rank=rank, gemm_mode=args.gemm_mode, max_act_batch=64)
        if world > 1:
            dp_parity.connect(pkg, d, args.comm, rank, world, dist, torch)
        d.init_params(seed=2, std=0.01)      # identical replicas: same seed on every rank
        return d

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_window(d, steps):
        """One timed window: ranks leave the barrier, run one untimed update (with N > 1 its gradient exchange
        re-aligns the ranks on the device), then exactly `steps` updates between two CUDA events; max over ranks."""
        barrier()
        d.update(1)
        torch.cuda.synchronize()
        ms = d.benchmark(steps)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity of this very configuration before anything is timed (N > 1: the driver's scaling runs carry it) ----
    parity = None
    if not args.no_extras and (world > 1 or os.environ.get("DQNB_BENCH_PARITY")):
        parity = dp_parity.check(pkg, dist, torch, rank, world, local, S, B, hidden, comm=args.comm, n_updates=2)

    rows = max(args.replay // world, 4 * B)
    d = make_learner(hidden, rows)
    s, a, r, mc, term, sn = synth_replay(rows, S, seed=1 + rank)
    for i in range(0, rows, 65536):
        j = min(rows, i + 65536)
        d.add_transitions(s[i:j], a[i:j], r[i:j], mc[i:j], sn[i:j], term[i:j])
    d.sync()

    # ---- device-timed value: median of `windows` windows of exactly K steps ----------------------------
    d.update(args.warmup)
    barrier()
    sampler.start()
    l0 = d.kernel_launches()
    windows = [timed_window(d, args.steps) for _ in range(max(1, args.windows))]
    launches = round((d.kernel_launches() - l0) / (len(windows) * (args.steps + 1)) * args.steps)   # kernels inside one timed window
    sampler.stop_flag = True
    sampler.join()
    barrier()
    ms_max = float(np.median(windows))
    value = B * world * args.steps / (ms_max * 1e-3)
    # roofline leg, taken right here under the same clocks as the timed windows (after the cuBLAS peak measurement
    # further down the GPU sits at its power cap for a while: the same replay then reads 13 instead of 8.7 us per launch)
    gemm_ms = n_gemm = gemm_err = None
    if rank == 0:
        try:
            gemm_ms, n_gemm = gemm_only_time(pkg, d, reps=50)
        except Exception as e:
            gemm_err = str(e)
    barrier()

    # ---- e2e through the C-ABI with host buffers ----------------------------------------------
    # Every step: B fresh host rows -> pinned staging -> H2D into the ring (dqnb_add_transitions), one
    # update (dqnb_update_async), and the (critic_loss, avg_q) of the previous step read back from the
    # host-mapped result ring (dqnb_results) while the current one runs - the learner loop of a caller
    # that logs its loss one step late.  All results are in host memory when the clock stops.  The
    # strictly blocking variant (dqnb_update returns the loss of the same step) is reported beside it.
    e2e_steps = args.steps
    fresh = synth_replay(B * 8, S, seed=100 + rank)

    def add_rows(k):
        o = (k % 8) * B
        d.add_transitions(fresh[0][o:o + B], fresh[1][o:o + B], fresh[2][o:o + B], fresh[3][o:o + B],
                          fresh[5][o:o + B], fresh[4][o:o + B])

    d.update(3)
    barrier()
    t0 = time.perf_counter()
    for k in range(min(e2e_steps, 500)):
        add_rows(k)
        loss, avgq = d.update(1)
    d.sync()
    dt_sync = (time.perf_counter() - t0) / min(e2e_steps, 500)
    barrier()
    d.update(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = 0
    for k in range(e2e_steps):
        add_rows(k)
        step = d.update_async(1)
        if k > 0:
            loss, avgq = d.results(step - 1, 1)
    loss, avgq = d.results(step, 1)
    d.sync()
    dt = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([dt, dt_sync], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = B * world / float(t[0].item())
    e2e_sync_val = B * world / float(t[1].item())
    Sp = (S + 63) // 64 * 64
    h2d = B * (2 * Sp + 16) * 4 + 8 + 4
    d2h = 8

    line = {
        "metric": "transition-updates/sec", "value": value, "unit": "transitions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 tensor-core products, fp32 accumulate)",
        "data": "synthetic", "config": workload_config(args, world), "implementation": implementation_note(args, world),
        "clocks": sampler.result(),
        "windows_ms_per_step": [w / args.steps for w in windows],
        "e2e": {"value": e2e_val, "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "blocking_value": e2e_sync_val,
                # an end-to-end loop cannot be faster than its device-timed core: anything above ~1.02 flags a skewed window
                "over_device_value": e2e_val / value,
                "what": "per step: dqnb_add_transitions(B host rows -> pinned -> H2D on the copy stream) + dqnb_update_async(1) + "
                        "dqnb_results of the previous step (host-mapped result ring); blocking_value = same loop with dqnb_update(1)"},
        "gpu_launches": int(launches),
        "final": {"critic_loss": float(loss[-1]), "avg_q": float(avgq[-1])},
    }
    if parity is not None:
        line["parity"] = parity

    # ---- extras: wide-MLP (BASELINE cfg5 per-GPU shape), act path, measured tensor peaks ----------------
    peaks = None
    if not args.no_extras:
        wide_hidden = (1024, 1024, 1024, 1024)
        wrows = 65536
        dw = make_learner(wide_hidden, wrows, seed_off=100)
        ws = synth_replay(wrows, S, seed=50 + rank)
        dw.add_transitions(ws[0], ws[1], ws[2], ws[3], ws[5], ws[4])
        dw.update(20)
        wsteps = min(args.steps, 300)
        wms = timed_window(dw, wsteps)
        wfpt = flop_per_transition(S, wide_hidden)
        line["wide_mlp"] = {
            "workload": f"cfg5 per-GPU shape: 1024x4 towers, batch {B} per GPU, global batch {B * world}",
            "value": B * world * wsteps / (wms * 1e-3), "unit": "transitions/s", "ms_per_step": wms / wsteps, "steps": wsteps,
            "achieved_tflops_over_step_per_gpu": wfpt * B / (wms / wsteps * 1e-3) / 1e12,
        }
        if rank == 0:
            try:
                gms, ng = gemm_only_time(pkg, dw, reps=20)
                line["wide_mlp"]["gemm_launches_per_step"] = ng
                line["wide_mlp"]["gemm_achieved_tflops"] = wfpt * B / (gms * 1e-3) / 1e12
            except Exception as e:
                line["wide_mlp"]["gemm_error"] = str(e)
        barrier()
        dw.close()
        # every rank runs the act-path leg: with N > 1 the updates in flight exchange gradients, so all ranks must
        # enqueue them (a lone rank would wait for its peers until the exchange times out)
        barrier()
        act_idle = act_latency(d, s)
        barrier()
        act_busy = act_latency(d, s, busy_updates=300)
        barrier()
        if rank == 0:
            line["act_path"] = {
                "unit": "us per dqnb_select_actions call (host rows in, host rows out; median of 200)",
                "rows": [1, 8, 64], "idle": act_idle, "while_updating": act_busy,
                "what": "skinny-M kernels on the act stream reading the actor snapshot; never waits for the learner's stream",
            }
            try:
                peaks = measure_tensor_peaks(torch)
            except Exception as e:
                peaks = {"error": str(e)}

    if rank == 0:
        hbm, bf16_burst, bf16_sus, src = measured_peaks()
        fpt = flop_per_transition(S, hidden)
        # dominant kernel: the layer GEMMs.  Timed live: all GEMM launches of one update, back to back.
        try:
            if gemm_ms is None:
                raise RuntimeError(gemm_err or "GEMM-only replay not measured")
            traffic, traffic_src = ncu_traffic_per_launch()
            ach = fpt * B / (gemm_ms * 1e-3) / 1e12
            line["roofline"] = {
                "bound": "tensor", "achieved": ach, "peak": bf16_sus, "unit": "TFLOP/s", "frac": ach / bf16_sus,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
                "kernel": "dqnb::gemm_tc_kernel", "launches_per_step": n_gemm, "avg_launch_us": 1e3 * gemm_ms / n_gemm,
                "algorithmic_flop_per_step": fpt * B,
                "note": "3xTF32 issues 3 TF32 MMA passes per algorithmic product: ceiling = measured TF32 peak / 3 "
                        "(live torch/cuBLAS 8192^3 measurement below; bf16 peak / 6 when it is unavailable)",
                # the same FLOP over the whole update (the GEMMs of independent branches overlap inside the step)
                "achieved_over_step": fpt * B / (ms_max / args.steps * 1e-3) / 1e12,
                "share_of_step": gemm_ms / (ms_max / args.steps),
            }
            tf32 = (peaks or {}).get("tf32_tflops_sustained")
            ceil3 = tf32 / 3.0 if tf32 else bf16_sus / 6.0
            line["roofline"]["tf32_peak_measured"] = tf32
            line["roofline"]["ceiling_3xtf32"] = ceil3
            line["roofline"]["frac_of_3xtf32_ceiling"] = ach / ceil3
            line["roofline"]["frac_of_3xtf32_ceiling_over_step"] = line["roofline"]["achieved_over_step"] / ceil3
            if "wide_mlp" in line and "gemm_achieved_tflops" in line["wide_mlp"]:
                w = line["wide_mlp"]
                w["frac_of_bf16_sustained"] = w["gemm_achieved_tflops"] / bf16_sus
                w["frac_of_3xtf32_ceiling"] = w["gemm_achieved_tflops"] / ceil3
                w["frac_of_3xtf32_ceiling_over_step"] = w["achieved_tflops_over_step_per_gpu"] / ceil3
        except Exception as e:  # keep the bench line even if the auxiliary measurement fails
            line["roofline"] = {"bound": "tensor", "achieved": None, "peak": bf16_sus, "unit": "TFLOP/s",
                                "frac": None, "traffic": None, "error": str(e)}
        if peaks:
            line["measured_peaks_live"] = peaks
        if not args.no_cpu_baseline and world >= 1:
            per, blas, cores = cpu_oracle_time(S, B, hidden, 1)
            n = int(max(1, min(50, args.cpu_seconds / max(per, 1e-3))))
            per, blas, cores = cpu_oracle_time(S, B, hidden, n)
            line["cpu_baseline"] = {
                "value": B / per, "unit": "transitions/s", "cores": cores, "kind": "port",
                "sample": f"{n} oracle UpdateActorCritic steps at batch {B} (Caffe operation order incl. "
                          f"force_backward work; {'OpenBLAS cblas_sgemm' if blas else 'portable C sgemm'})",
                "ms_per_update": per * 1e3,
            }
        print(json.dumps(line))
    if world > 1 and d.comm_status() != 0:
        raise SystemExit("gradient exchange timed out on a peer: the run is invalid")
    if parity is not None and not parity["ok"]:
        raise SystemExit("data-parallel parity check failed: " + json.dumps(parity))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(group=cpu_group)       # the other ranks wait on the host while rank 0 runs the CPU leg
    d.close()
    if dist is not None:
        dist.destroy_process_group()


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per gemm_tc_kernel launch, averaged over the launches of
    the newest committed `ncu --set full` summary (profiles/*_ncu_gemm_full.txt).  None if no capture."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_gemm_full.txt")), key=os.path.basename)   # r01 < r01b < r01d < r02 ..
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for line in open(files[-1]):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m:
            tot += float(m.group(2)) * unit.get(m.group(3), 1.0)
            n += m.group(1) == "read"
    return (tot / n if n else None), os.path.basename(files[-1])


def gemm_only_time(pkg, d, reps=50):
    """Average device time of the GEMM launches of one update (CUDA events on the handle's stream)."""
    lib = pkg.lib()
    if not hasattr(lib, "dqnb_benchmark_gemms"):
        raise RuntimeError("dqnb_benchmark_gemms not exported")
    lib.dqnb_benchmark_gemms.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    ms, n = C.c_float(), C.c_int32()
    rc = lib.dqnb_benchmark_gemms(d._h, reps, C.byref(ms), C.byref(n))
    if rc != 0:
        raise RuntimeError(lib.dqnb_last_error().decode())
    return ms.value, n.value


if __name__ == "__main__":
    main()

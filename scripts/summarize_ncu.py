"""Turns the ncu output of scripts/gpu_r02_profile.sh (gpurun_out/<tag>_launches.csv, <tag>_gemm_raw.csv,
<tag>_small_raw.csv) into the text summaries kept under profiles/.

usage: python scripts/summarize_ncu.py r02z r02          (input tag, output tag)
"""
import csv
import os
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "r02z"
dst = sys.argv[2] if len(sys.argv) > 2 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")
HBM_PEAK = 6542.1   # GB/s, MEASURED_PEAKS.json hbm_gbs


def short(name):
    name = re.sub(r"\(.*", "", name).replace("dqnb::", "").replace("void ", "")
    m = re.match(r"gemm_tc_kernel<(\d+), (\d+), (\d+), (\d+)>", name)
    if m:
        epi = {"0": "fwd", "1": "dx", "2": "plain"}[m.group(4)]
        return f"gemm_tc<{'MN' if m.group(1) == '1' else 'K'}/{'MN' if m.group(2) == '1' else 'K'},bn{m.group(3)},{epi}>"
    return name


def to_ns(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


# ---- launch list: the launches of the LAST update of the run ---------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(G, f"{src}_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
names = [short(r[4]) for r in rows]
gathers = [i for i, n in enumerate(names) if n.startswith("gather_kernel")]
# one update = from one gather to the next (the gather runs ahead on its own stream, so it is listed first)
lo, hi = gathers[-2], gathers[-1]
upd = rows[lo:hi]
out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none (scripts/gpu_r02_profile.sh), bench.py --steps 2 --warmup 3:",
       f"# the {len(upd)} kernel launches of one update (graph-replayed; ncu serialises the graph's kernel nodes and runs them",
       "# cold-cache, so compare SHARES, not absolute times), cfg2 shape S=58 B=1024 1024-512-256-128"]
tot, agg = 0.0, {}
for r in upd:
    ns = to_ns(r[-1], r[-2])
    n = short(r[4])
    tot += ns
    agg.setdefault(n, [0.0, 0])
    agg[n][0] += ns; agg[n][1] += 1
    out.append(f"{r[0]:>5} {n:44s} grid={r[8]:16s} {ns / 1000:8.2f} us")
out.append(f"TOTAL {tot / 1000:.1f} us serialised")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append(f"{k:50s} n={v[1]:3d} {v[0] / 1000:8.1f} us  {100 * v[0] / tot:5.1f}%")
gemm_share = sum(v[0] for k, v in agg.items() if k.startswith("gemm_tc")) / tot
out.append(f"all gemm_tc_kernel instances: {100 * gemm_share:.1f}% of the serialised kernel time")
open(os.path.join(PR, f"{dst}_ncu_launches.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-16:]))


def raw_summary(path, want, title):
    rr = list(csv.reader(open(path)))
    hdr, units = rr[0], rr[1]
    idx = {h: i for i, h in enumerate(hdr)}
    o = [title]
    for r in rr[2:]:
        o.append("---")
        o.append(f"  {'kernel':75s} {short(r[idx['Kernel Name']])}")
        for w in want:
            if w in idx:
                o.append(f"  {w:75s} {r[idx[w]]} {units[idx[w]]}")
        if "dram__bytes_read.sum" in idx and "gpu__time_duration.sum" in idx:
            b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            ns = to_ns(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])
            o.append(f"  {'derived: DRAM GB/s (read+write bytes / duration)':75s} {b / ns:.1f} GB/s = {100 * b / ns / HBM_PEAK:.1f}% of the measured {HBM_PEAK} GB/s")
    return o, len(rr) - 2


WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
p = os.path.join(G, f"{src}_gemm_raw.csv")
if os.path.exists(p):
    o, n = raw_summary(p, WANT, "# ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel (scripts/gpu_r02_profile.sh): the 38 GEMM "
                       "launches of one update, cfg2 shape; cold-cache serialised replays")
    open(os.path.join(PR, f"{dst}_ncu_gemm_full.txt"), "w").write("\n".join(o) + "\n")
    print(n, "gemm captures summarised")
p = os.path.join(G, f"{src}_small_raw.csv")
if os.path.exists(p):
    o, n = raw_summary(p, WANT, "# ncu --set full --clock-control none: the memory-/latency-bound kernels of one update (adam, reduce, gather, colsum, head kernels), "
                       "cfg2 shape; cold-cache serialised replays (adam's operands come from HBM here; inside the real step part of them is L2-resident)")
    open(os.path.join(PR, f"{dst}_ncu_small_kernels.txt"), "w").write("\n".join(o) + "\n")
    print(n, "small-kernel captures summarised")

"""Turns gpurun_out/launches.csv + prof_gemm.ncu-rep into the text summaries kept under profiles/."""
import csv, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
per_update = int(sys.argv[2]) if len(sys.argv) > 2 else 53
rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 10 and r[0].isdigit()]
rows = rows[-per_update:]
out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none (scripts/gpu_profile.sh); last update of the run ({per_update} launches),",
       "# cfg2 shape S=58 B=1024 1024-512-256-128, eager launches; per-launch times are cold-cache and serialised: compare SHARES."]
tot = 0; agg = {}
for r in rows:
    val = float(r[-1].replace(',', '')); unit = r[-2]
    ns = val if unit in ('ns', 'nsecond') else val * 1000
    short = re.sub(r'\(.*', '', r[4]).replace('dqnb::', '').replace('void ', '')
    tot += ns; agg.setdefault(short, [0, 0]); agg[short][0] += ns; agg[short][1] += 1
    out.append(f"{r[0]:>4} {short[:40]:40s} grid={r[8]:16s} {ns/1000:8.2f} us")
out.append(f"TOTAL {tot/1000:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append(f"{k[:50]:50s} n={v[1]:3d} {v[0]/1000:8.1f} us  {100*v[0]/tot:5.1f}%")
open(f'profiles/{tag}_ncu_launches.txt', 'w').write("\n".join(out) + "\n")
print("\n".join(out[-14:]))
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_gemm.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
o2 = ["# ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel (scripts/gpu_profile.sh), cfg2 shape; cold-cache serialised replays"]
for r in rr[2:]:
    o2.append('---')
    for w in want:
        if w in idx:
            o2.append(f"  {w:75s} {r[idx[w]]} {units[idx[w]]}")
open(f'profiles/{tag}_ncu_gemm_full.txt', 'w').write("\n".join(o2) + "\n")
print(len(rr) - 2, "gemm captures summarised")

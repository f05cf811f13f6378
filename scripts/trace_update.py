"""GPU: timeline of every kernel of one update (DQNB_TRACE=1), from in-kernel globaltimer stamps.

usage: python scripts/trace_update.py [batch]
GEMM rows : CTA (0,0,0): entry, pdl_wait passed, first operands, accumulators done, epilogue stores issued;
            grid-wide: latest CTA to pass pdl_wait, latest CTA exit.
other rows: earliest / latest CTA past pdl_wait, latest exit of thread 0 of any CTA.
"""
import os, sys, ctypes as C
os.environ["DQNB_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import synth_replay
P = load_package()
L = P.lib()
L.dqnb_debug_trace.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int64]
L.dqnb_debug_trace.restype = C.c_int64
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
if len(sys.argv) > 2:          # knob overrides, e.g. '{"DQNB_SCHED": 1}'
    import json
    for k, v in json.loads(sys.argv[2]).items():
        os.environ[k] = str(v)
    print("knobs:", sys.argv[2])
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:          # under torchrun: the data-parallel step (exchange kernels in the timeline); rank 0 prints
    import torch
    import torch.distributed as dist
    from scripts import dp_parity
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
d = P.DQNB(device=local, state_size=58, batch=B, hidden=(1024, 512, 256, 128), replay_capacity=70000, use_graph=1, world_size=world, rank=rank,
           seed=3 + rank)
if world > 1:
    dp_parity.connect(P, d, "p2p", rank, world, dist, torch)
d.init_params(2, 0.01)
s, a, r, mc, term, sn = synth_replay(65536, 58, 1)
d.add_transitions(s, a, r, mc, sn, term)
d.update(20)
ms = d.benchmark(200)
print(f"graph replay (tracing on): {ms/200*1e3:.1f} us per update")
d.update(1)
if os.environ.get("TRACE_GEMMS_ONLY"):      # timeline of bench.py's roofline leg: the update's GEMM launches back to back on one stream
    import bench
    ms, ng = bench.gemm_only_time(P, d, reps=3)
    print(f"GEMM-only replay: {ng} launches, {ms*1e3:.1f} us per replay")
buf = (C.c_longlong * 16384)()
n = L.dqnb_debug_trace(d._h, buf, 16384)
t = np.array(list(buf[:n]), dtype=np.int64).reshape(-1, 16)
kinds = ["GEMM", "GEMM_GROUP", "GATHER", "SAMPLE", "HEAD_FWD", "CRITIC_HEAD", "ACTOR_HEAD_BWD", "HEAD_BWD_W", "COLSUM",
         "REDUCE", "ALLREDUCE", "P2P_ALLREDUCE", "ADAM", "PREP", "FINALIZE", "FORK", "JOIN"]
valid = t[:, 0] > 0
t0 = t[valid, 0].min()
us = lambda v: (v - t0) / 1e3 if v > 0 else float("nan")
print(" op kind           br grid(x,y,z)  kb |  entry  waited(all)  operands(all)  acc_done(all)  epi_done(all) exit(all) | wait->exit")
for i in range(len(t)):
    meta = int(t[i, 7])
    k, br = (meta & 0xffff) // 1000, (meta & 0xffff) % 1000
    name = kinds[k] if k < len(kinds) else str(k)
    if name in ("FORK", "JOIN"):
        print(f"{i:3d} {name:16s}")
        continue
    if t[i, 0] <= 0:
        print(f"{i:3d} {name:16s} {br}   (not launched)")
        continue
    if name == "GEMM":
        gx, gy, gz, kb = (meta >> 16) & 0xfff, (meta >> 28) & 0xfff, (meta >> 40) & 0xff, (meta >> 48) & 0xfff
        print(f"{i:3d} {name:16s} {br} ({gx:2d},{gy:2d},{gz:2d}) {kb:3d} | {us(t[i,0]):7.1f} {us(t[i,2]):6.1f}({us(t[i,1]):6.1f}) "
              f"{us(t[i,3]):6.1f}({us(t[i,8]):6.1f}) {us(t[i,4]):6.1f}({us(t[i,9]):6.1f}) {us(t[i,5]):6.1f}({us(t[i,10]):6.1f}) {us(t[i,6]):8.1f} | {us(t[i,6]) - us(t[i,2]):6.1f}"
              f" | epi: acc+{us(t[i,11]) - us(t[i,4]):4.2f} staged+{us(t[i,12]) - us(t[i,11]):4.2f} stored+{us(t[i,5]) - us(t[i,12]):4.2f} synced+{us(t[i,13]) - us(t[i,5]):4.2f}")
    else:
        extra = ""
        if name == "P2P_ALLREDUCE":
            extra = (f" | flagA +{us(t[i,3]) - us(t[i,0]):4.1f}  summed +{us(t[i,4]) - us(t[i,3]):4.1f}  fenced +{us(t[i,5]) - us(t[i,4]):4.1f}"
                     f"  last block +{us(t[i,8]) - us(t[i,5]):4.1f}  flagB sent +{us(t[i,9]) - us(t[i,8]):4.1f}  flagB seen +{us(t[i,10]) - us(t[i,9]):4.1f}")
        print(f"{i:3d} {name:16s} {br}                  | {'':7s} {us(t[i,0]):6.1f}({us(t[i,1]):6.1f}) {'':14s} {'':14s} {'':14s} {us(t[i,2]):8.1f} | "
              f"{us(t[i,2]) - us(t[i,0]):6.1f}{extra}")
d.close()

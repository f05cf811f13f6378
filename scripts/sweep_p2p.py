"""Multi-GPU A/B of exchange knobs in ONE set of processes (same box, same clocks):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
      scripts/sweep_p2p.py '{"DQNB_P2P_PUSH": 0}' '{}'
"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import synth_replay
from scripts import dp_parity

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = load_package()
s, a, r, mc, term, sn = synth_replay(65536, 58, 1 + rank)
cfgs = [json.loads(x) for x in sys.argv[1:]] or [{}]
KEYS = set(k for c in cfgs for k in c)
for rep in range(2):
    for cfg in cfgs:
        for k in KEYS:
            os.environ.pop(k, None)
        for k, v in cfg.items():
            os.environ[k] = str(v)
        for hidden in ((1024, 512, 256, 128),):
            d = P.DQNB(device=local, state_size=58, batch=1024, hidden=hidden, replay_capacity=70000, seed=3 + rank,
                       world_size=world, rank=rank)
            dp_parity.connect(P, d, "p2p", rank, world, dist, torch)
            d.init_params(2, 0.01)
            d.add_transitions(s, a, r, mc, sn, term)
            d.update(30)
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize(); dist.barrier(); d.update(1); torch.cuda.synchronize()
                t = torch.tensor([d.benchmark(300) / 300 * 1e3], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = min(best, float(t.item()))
            st = d.comm_status()
            d.close()
            if rank == 0:
                tag = " ".join(f"{k[5:]}={v}" for k, v in cfg.items()) or "(defaults)"
                print(f"{best:7.1f} us/update  x{world}  {1024 * world / best:6.3f}e6 tr/s  comm_status={st}  {tag}", flush=True)
dist.barrier()
dist.destroy_process_group()

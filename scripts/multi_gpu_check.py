"""Run under torchrun (one process per GPU): data-parallel update == single learner at the global
batch (oracle), and replicas stay bit-identical.  Used by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 scripts/multi_gpu_check.py [small,cfg2,cfg3,cfg5]
Cases (per-GPU batch; the global batch is batch * world):
  small : S=58, B=128, 256-128-64-64, both exchange paths (P2P kernel and ncclAllReduce), 5 updates
  cfg2  : BASELINE cfg2 shape, S=58, B=1024
  cfg3  : BASELINE cfg3, S=77 (2v1), B=2048 per GPU -> global 4096 on 2 GPUs
  cfg5  : BASELINE cfg5, 1024x4 towers, B=1024 per GPU -> global 8192 on 8 GPUs
The BASELINE shapes run twice: frozen weights (1e-4 on every tap incl. both all-reduced gradients) and the
reference learning rates (bookkeeping, parameters within Adam's step ambiguity, bit-identical replicas).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from scripts import dp_parity

SHAPES = {
    "small": (58, 128, (256, 128, 64, 64)),
    "cfg2": (58, 1024, (1024, 512, 256, 128)),
    "cfg3": (77, 2048, (1024, 512, 256, 128)),
    "cfg5": (58, 1024, (1024, 1024, 1024, 1024)),
}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    cases = (sys.argv[1] if len(sys.argv) > 1 else "small").split(",")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = load_package()
    ok_all = True
    for case in cases:
        S, B, hidden = SHAPES[case]
        runs = [("p2p", False, 5), ("nccl", False, 5)] if case == "small" else [("p2p", True, 1), ("p2p", False, 2)]
        for comm, frozen, n_updates in runs:
            res = dp_parity.check(P, dist, torch, rank, world, local, S, B, hidden, comm=comm, n_updates=n_updates, frozen=frozen)
            ok_all &= res["ok"]
            if rank == 0:
                print(f"[{case} x{world} {comm} {'frozen' if frozen else 'lr'}] " + json.dumps(res), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok_all else "MULTI_GPU_FAIL")
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()

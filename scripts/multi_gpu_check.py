"""Run under torchrun (one process per GPU): data-parallel update == single learner at the global
batch (oracle), and replicas stay bit-identical.  Used by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 scripts/multi_gpu_check.py
"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from oracle import oracle as O

def connect(P, d, comm, rank, world):
    """Wires the replicas together: 'p2p' = the IPC/NVLink exchange kernel (product path), 'nccl' = ncclAllReduce."""
    if comm == "nccl":
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(P.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        d.comm_init(bytes(idt.cpu().numpy().tobytes()))
    else:
        mine = torch.frombuffer(bytearray(d.comm_p2p_handle()), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        d.comm_p2p_init(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = load_package()
    ok_all = True
    for comm in ("p2p", "nccl"):
        ok_all &= run(P, comm, rank, world, local)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok_all else "MULTI_GPU_FAIL")
    sys.exit(0 if ok_all else 1)


def run(P, comm, rank, world, local):
    S, B, hidden = 58, 128, (256, 128, 64, 64)
    n_updates = 5
    rng = np.random.default_rng(0)                     # same stream on every rank
    ocfg = O.make_config(state_size=S, batch=B * world, hidden=hidden)
    a0, c0 = O.init_params(ocfg, False, rng, "warm"), O.init_params(ocfg, True, rng, "warm")
    st = O.OracleState(ocfg, a0, c0, a0, c0)
    n = 1000
    shards = [O.synth_batch(O.make_config(state_size=S, batch=n, hidden=hidden), rng) for _ in range(world)]
    idx = rng.integers(0, n, (n_updates, world, B)).astype(np.int32)
    d = P.DQNB(device=local, state_size=S, batch=B, hidden=hidden, replay_capacity=2048, world_size=world, rank=rank)
    connect(P, d, comm, rank, world)
    d.set_params(P.ACTOR, a0); d.set_params(P.CRITIC, c0); d.clone_targets()
    s, a, r, mc, term, sn = shards[rank]
    d.add_transitions(s, a, r, mc, sn, term)
    ok = True
    for u in range(n_updates):
        loss, avgq = d.update_with_indices(idx[u, rank])
        cat = lambda k: np.concatenate([shards[w][k][idx[u, w]] for w in range(world)])
        oloss, oavgq = st.update(cat(0), cat(1), cat(2), cat(3), cat(4), cat(5))
        e1 = abs(loss - oloss) / abs(oloss); e2 = abs(avgq - oavgq) / (abs(oavgq) + 1e-6)
        if rank == 0:
            print(f"[{comm}] update {u}: loss {loss:.6f} oracle {oloss:.6f} rel {e1:.2e} | avg_q {avgq:.6f} oracle {oavgq:.6f} rel {e2:.2e}")
        ok &= e1 < 5e-4 and e2 < 5e-4
    for net, ref, lr in ((P.CRITIC, st.critic, 1e-3), (P.ACTOR, st.actor, 1e-5), (P.CRITIC_TARGET, st.critic_target, 1e-6)):
        got = d.get_params(net)
        err = np.abs(got - ref).max()
        t = torch.from_numpy(got).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        if rank == 0:
            print(f"net {net}: max|param - oracle| = {err:.3e}; replicas bit-identical: {same}")
        ok &= bool(same) and err < 0.5 * lr * n_updates + 1e-7
    ok &= d.comm_status() == 0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    dist.barrier()
    d.close()
    return flag.item() == 1

if __name__ == "__main__":
    main()

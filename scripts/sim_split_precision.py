"""CPU study: which split-operand tensor-core product keeps UpdateActorCritic within 1e-4 of fp32/float64?

Emulates the GEMM numerics of the candidate tcgen05 modes inside the float64 autograd model
(tests/refmodel.py) and reports the per-tensor relative error (max|a-b| / max|b|, the parity metric
of tests/util.py) against the exact float64 evaluation with frozen weights (lr = 0):

  tf32x3 : operands hi = tf32(x), lo = x - hi (exact), product hi*hi + lo*hi + hi*lo   (round-1 mode)
  bf16x2 : operands hi = bf16(x), lo = bf16(x - hi), product hi*hi + lo*hi + hi*lo      (3 kind::f16 MMAs,
           half the tensor time and half the operand bytes)
  bf16x3 : hi, mid, lo bf16 (24 bits), 6 products                                         (same cost as tf32x3)

Stored activations are what the planes can represent (hi + lo), i.e. bf16x2 activations carry 16-17
mantissa bits.  Products are accumulated in float64 here, so the numbers are the *representation* error
of each mode; the tensor core's fp32 accumulation adds ~1e-6 on top (measured in round 1).

    python scripts/sim_split_precision.py [cfg2|cfg3|cfg5]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
import refmodel as R  # noqa: E402

MODE = "exact"


def rnd_bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def rnd_tf32(x):
    u = x.to(torch.float32).view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF
    return u.view(torch.float32).to(torch.float64)


def planes(x):
    """The planes a mode stores for x (list of float64 tensors whose sum is the stored value)."""
    x = x.to(torch.float32).to(torch.float64)
    if MODE == "tf32x3":
        hi = rnd_tf32(x)
        return [hi, rnd_tf32(x - hi)]          # the tensor core truncates lo to tf32 as well
    if MODE == "bf16x2":
        hi = rnd_bf16(x)
        return [hi, rnd_bf16(x - hi)]
    if MODE == "bf16x3":
        hi = rnd_bf16(x)
        mid = rnd_bf16(x - hi)
        return [hi, mid, rnd_bf16(x - hi - mid)]
    return [x]


def stored(x):
    if MODE in ("exact", "tf32x3"):
        return x
    return sum(planes(x))


def prod(a, b):
    """a @ b with the mode's retained cross terms (plane i of a times plane j of b for i + j < n)."""
    if MODE == "exact":
        return a @ b
    pa, pb = planes(a), planes(b)
    n = len(pa)
    out = 0
    for i in range(n):
        for j in range(n):
            if i + j < n:
                out = out + pa[i] @ pb[j]
    return out


class SplitLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W):
        ctx.save_for_backward(x, W)
        return prod(x, W.T)

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        return prod(g, W), prod(g.T, x)


class Quant(torch.autograd.Function):
    """activation / gradient as the planes store it (straight-through)"""
    @staticmethod
    def forward(ctx, x):
        return stored(x)

    @staticmethod
    def backward(ctx, g):
        return stored(g)


def tower(cfg, v, x):
    x = Quant.apply(x)
    for l in range(cfg.n_hidden):
        x = SplitLinear.apply(x, v[f"ip{l+1}.W"]) + v[f"ip{l+1}.b"]
        x = Quant.apply(torch.nn.functional.leaky_relu(x, 0.01))
    return x


def run(cfg, st64, batch, mode):
    global MODE
    MODE = mode
    R.tower = tower
    _, dia = R.update(cfg, st64, *batch)
    return dia


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    S, B, hidden = {"cfg2": (58, 1024, (1024, 512, 256, 128)), "cfg3": (77, 4096, (1024, 512, 256, 128)),
                    "cfg5": (58, 1024, (1024, 1024, 1024, 1024)), "small": (58, 128, (256, 128, 64, 32))}[name]
    cfg = O.make_config(state_size=S, batch=B, hidden=hidden, actor_lr=0.0, critic_lr=0.0)
    rng = np.random.default_rng(3)
    a = O.init_params(cfg, False, rng, "warm")
    c = O.init_params(cfg, True, rng, "warm")
    at = (a + rng.normal(0, 1e-3, a.size)).astype(np.float32)
    ct = (c + rng.normal(0, 1e-3, c.size)).astype(np.float32)
    z = lambda n: np.zeros(n, np.float64)
    st = dict(actor=a.astype(np.float64), critic=c.astype(np.float64), actor_target=at.astype(np.float64),
              critic_target=ct.astype(np.float64), actor_m=z(a.size), actor_v=z(a.size), critic_m=z(c.size),
              critic_v=z(c.size), actor_iter=0, critic_iter=0)
    batch = O.synth_batch(cfg, rng, p_term=0.2)
    ref = run(cfg, st, batch, "exact")
    keys = ["y", "q", "critic_grad", "a_pi", "q_pi", "d_raw", "d_inv", "actor_grad"]
    print(f"{name}: S={S} B={B} hidden={hidden}   relerr = max|a-b|/max|b| vs exact float64")
    print("mode      " + " ".join(f"{k:>12s}" for k in keys + ["critic_loss", "avg_q"]))
    for mode in ("tf32x3", "bf16x3", "bf16x2"):
        d = run(cfg, st, batch, mode)
        errs = []
        for k in keys:
            x, y = np.asarray(d[k], np.float64).ravel(), np.asarray(ref[k], np.float64).ravel()
            errs.append(np.abs(x - y).max() / (np.abs(y).max() + 1e-300))
        errs.append(abs(d["critic_loss"] - ref["critic_loss"]) / abs(ref["critic_loss"]))
        errs.append(abs(d["avg_q"] - ref["avg_q"]) / abs(ref["avg_q"]))
        print(f"{mode:9s} " + " ".join(f"{e:12.2e}" for e in errs))


if __name__ == "__main__":
    main()

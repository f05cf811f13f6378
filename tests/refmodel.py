"""Independent float64 autograd derivation of UpdateActorCritic (dqn.cpp:828-972).

Used ONLY to validate the C oracle (the second of the two derivations SURVEY §8c asks for):
losses are written down (L_c = sum (q-y)^2 / (2B), L_a = -sum Q(s, pi(s)) with the inverting-
gradients rescale applied to dL/da) and torch.autograd produces every gradient; nothing here
shares code with oracle/dqn_oracle.c.
"""
import numpy as np
import torch

from oracle import oracle as O


def unpack(cfg, flat, critic):
    t = torch.tensor(np.asarray(flat, dtype=np.float64), requires_grad=True)
    views = {}
    for name, off, shape in O.net_blobs(cfg, critic):
        views[name] = t[off:off + int(np.prod(shape))].reshape(shape)
    return t, views


def tower(cfg, v, x):
    for l in range(cfg.n_hidden):
        x = x @ v[f"ip{l+1}.W"].T + v[f"ip{l+1}.b"]
        x = torch.nn.functional.leaky_relu(x, 0.01)
    return x


def actor_fwd(cfg, v, s):
    h = tower(cfg, v, s)
    return torch.cat([h @ v["action_layer.W"].T + v["action_layer.b"],
                      h @ v["actionpara_layer.W"].T + v["actionpara_layer.b"]], dim=1)


def critic_fwd(cfg, v, s, a10):
    h = tower(cfg, v, torch.cat([s, a10], dim=1))
    return (h @ v["q_values_layer.W"].T + v["q_values_layer.b"]).reshape(-1)


def invert(a, d):
    lo = torch.tensor([-1.0] * 4 + [0, -180, -180, -180, 0, -180], dtype=torch.float64)
    hi = torch.tensor([1.0] * 4 + [100, 180, 180, 180, 100, 180], dtype=torch.float64)
    up = (hi - a) / (hi - lo)
    dn = (a - lo) / (hi - lo)
    return torch.where(d < 0, d * up, torch.where(d > 0, d * dn, d))


def adam(cfg, lr, p, g, m, v, it):
    g = g.clone()
    norm = float(torch.sqrt((g * g).sum()))
    if cfg.clip_gradients >= 0 and norm > cfg.clip_gradients:
        g = g * (cfg.clip_gradients / norm)
    b1, b2 = float(np.float32(cfg.momentum)), float(np.float32(cfg.momentum2))
    t = it + 1
    corr = np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - float(np.float32(lr)) * corr * m / (torch.sqrt(v) + float(np.float32(cfg.delta)))
    return p, m, v, norm


def update(cfg, st, s, a10, r, mc, term, sn):
    """st: dict of float64 numpy arrays (actor, critic, actor_target, critic_target, *_m, *_v) and
    iters.  Returns new state dict + diagnostics.  Pure float64."""
    B = cfg.batch
    T = lambda x: torch.tensor(np.asarray(x, dtype=np.float64))
    s, a10, r, mc, sn = T(s), T(a10), T(r), T(mc), T(sn)
    term = torch.tensor(np.asarray(term).astype(bool))
    out = {}
    with torch.no_grad():
        _, vat = unpack(cfg, st["actor_target"], False)
        _, vct = unpack(cfg, st["critic_target"], True)
        qn = critic_fwd(cfg, vct, sn, actor_fwd(cfg, vat, sn))
        off = torch.where(term, r, r + cfg.gamma * qn)
        y = cfg.beta * mc + (1 - cfg.beta) * off
    out["y"] = y.numpy()
    tc, vc = unpack(cfg, st["critic"], True)
    q = critic_fwd(cfg, vc, s, a10)
    loss = ((q - y) ** 2).sum() / (2 * B)
    (gc,) = torch.autograd.grad(loss, tc)
    out["critic_grad"], out["critic_loss"], out["q"] = gc.numpy(), float(loss.detach()), q.detach().numpy()
    newc, cm, cv, out["critic_gnorm"] = adam(cfg, cfg.critic_lr, tc.detach(), gc, T(st["critic_m"]),
                                             T(st["critic_v"]), st["critic_iter"])
    ta, va = unpack(cfg, st["actor"], False)
    _, vc2 = unpack(cfg, newc.numpy(), True)
    a_pi = actor_fwd(cfg, va, s)
    a_leaf = a_pi.detach().clone().requires_grad_(True)
    q_pi = critic_fwd(cfg, vc2, s, a_leaf)
    (d_raw,) = torch.autograd.grad(-q_pi.sum(), a_leaf)
    d_inv = invert(a_pi.detach(), d_raw)
    (ga,) = torch.autograd.grad(a_pi, ta, grad_outputs=d_inv)
    out.update(a_pi=a_pi.detach().numpy(), q_pi=q_pi.detach().numpy(), d_raw=d_raw.numpy(),
               d_inv=d_inv.numpy(), actor_grad=ga.numpy(), avg_q=float(q_pi.detach().mean()))
    newa, am, av, out["actor_gnorm"] = adam(cfg, cfg.actor_lr, ta.detach(), ga, T(st["actor_m"]),
                                            T(st["actor_v"]), st["actor_iter"])
    ns = dict(actor=newa.numpy(), critic=newc.numpy(), actor_m=am.numpy(), actor_v=av.numpy(),
              critic_m=cm.numpy(), critic_v=cv.numpy(), actor_iter=st["actor_iter"] + 1,
              critic_iter=st["critic_iter"] + 1, actor_target=np.asarray(st["actor_target"], np.float64),
              critic_target=np.asarray(st["critic_target"], np.float64))
    if cfg.soft_update_freq > 0 and max(ns["actor_iter"], ns["critic_iter"]) % cfg.soft_update_freq == 0:
        tau = float(np.float32(cfg.tau))
        ns["critic_target"] = (1 - tau) * ns["critic_target"] + tau * ns["critic"]
        ns["actor_target"] = (1 - tau) * ns["actor_target"] + tau * ns["actor"]
    return ns, out

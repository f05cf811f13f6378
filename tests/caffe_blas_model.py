"""Second CPU restatement of UpdateActorCritic (dqn.cpp:828-972), written as the sequence of BLAS calls BVLC Caffe
(pinned by the reference's README at 2ef584785c8ade90260eb117f189146364494183) issues for it - independent of
oracle/dqn_oracle.c (different language, different GEMM, no shared code) and of tests/refmodel.py (no autograd).

TEST INFRASTRUCTURE ONLY.  Every array is float32; `sgemm` is numpy's matmul on float32 operands (OpenBLAS
cblas_sgemm), the level-1 calls are single float32 numpy operations with the same number of roundings as the BLAS
call they stand for.  Call sequence per Caffe layer / solver function:

  InnerProductLayer::Forward_cpu   gemm(N, T, M, N, K, 1, bottom, W, 0, top); gemm(N, N, M, N, 1, 1, ones, b, 1, top)
  ReLULayer::Forward_cpu           top = max(x, 0) + slope * min(x, 0)                     (in place)
  ReLULayer::Backward_cpu          bottom_diff = top_diff * ((bottom_data > 0) + slope * (bottom_data <= 0))
  InnerProductLayer::Backward_cpu  gemm(T, N, N, K, M, 1, top_diff, bottom, 1, W_diff); gemv(T, M, N, 1, top_diff, ones, 1, b_diff);
                                   gemm(N, N, M, K, N, 1, top_diff, W, 0, bottom_diff)
  SplitLayer::Backward_cpu         caffe_add(diff0, diff1)                                   (the actor's two heads)
  EuclideanLossLayer               sub; dot; loss = dot / num / 2;  axpby(1 / num, diff, 0, bottom_diff)
  SGDSolver::ClipGradients         sum of blob.sumsq_diff() (sdot per blob, float accumulation); scal(clip / l2)
  AdamSolver::ComputeUpdateValue   axpby(1-b1, g, b1, m); mul(g, g, t); axpby(1-b2, t, b2, v); powx(v, .5, t);
                                   add_scalar(eps, t); div(m, t, t); scale(lr * correction, t, g)
  Net::Update                      axpy(-1, diff, data)
  DQN::SoftUpdateNet               axpby(tau, from, 1 - tau, to)                             (dqn.cpp:1085-1096)
"""
import numpy as np

F = np.float32
SLOPE = F(0.01)   # dqn.cpp:300


def axpby(alpha, x, beta, y):
    """caffe_cpu_axpby = cblas_sscal(beta, y); cblas_saxpy(alpha, x, y): two roundings per element."""
    y *= F(beta)
    y += F(alpha) * x
    return y


class Net:
    """One Caffe net of the reference (Tower + heads, dqn.cpp:400-454) over a flat learnable_params vector."""

    def __init__(self, blobs, params, critic):
        self.critic = critic
        self.p = params                       # float32 view, updated in place
        self.g = np.zeros_like(params)        # param diffs
        self.v = {name: (off, shape) for name, off, shape in blobs}
        self.n_hidden = sum(1 for name, _, _ in blobs if name.startswith("ip") and name.endswith(".W"))
        self.heads = ["q_values_layer"] if critic else ["action_layer", "actionpara_layer"]

    def W(self, name, diff=False):
        off, shape = self.v[name]
        return (self.g if diff else self.p)[off:off + int(np.prod(shape))].reshape(shape)

    def forward(self, x):
        """Returns head outputs; keeps bottoms / in-place activations for the backward pass."""
        self.bottoms, ones = [], np.ones((x.shape[0], 1), F)
        for l in range(1, self.n_hidden + 1):
            self.bottoms.append(x)
            top = x @ self.W(f"ip{l}.W").T                      # gemm(NoTrans, Trans)
            top += ones @ self.W(f"ip{l}.b")[None, :]           # rank-1 bias gemm
            top = np.maximum(top, F(0)) + SLOPE * np.minimum(top, F(0))
            x = top
        self.top = x
        outs = []
        for h in self.heads:
            o = x @ self.W(h + ".W").T
            o += ones @ self.W(h + ".b")[None, :]
            outs.append(o)
        return outs

    def backward(self, head_diffs, param_grads=True):
        """BackwardFrom the last head layer: accumulates param diffs, returns the diff at the net input."""
        x, ones = self.top, np.ones(self.top.shape[0], F)
        d = None
        for h, td in zip(self.heads, head_diffs):
            if param_grads:
                self.W(h + ".W", True)[...] += td.T @ x          # gemm(Trans, NoTrans), beta = 1
                self.W(h + ".b", True)[...] += td.T @ ones       # gemv(Trans)
            bd = td @ self.W(h + ".W")                           # gemm(NoTrans, NoTrans)
            d = bd if d is None else d + bd                      # SplitLayer: caffe_add
        for l in range(self.n_hidden, 0, -1):
            y = x
            d = d * ((y > 0).astype(F) + SLOPE * (y <= 0).astype(F))     # ReLU backward on the in-place blob
            x = self.bottoms[l - 1]
            if param_grads:
                self.W(f"ip{l}.W", True)[...] += d.T @ x
                self.W(f"ip{l}.b", True)[...] += d.T @ ones
            d = d @ self.W(f"ip{l}.W")
        return d


def solver_apply_update(net, m, v, it, lr, cfg):
    """SGDSolver::ApplyUpdate with AdamSolver: ClipGradients, ComputeUpdateValue per blob, Net::Update.
    Returns the L2 norm ClipGradients saw."""
    sumsq = F(0)
    for name, (off, shape) in net.v.items():                     # learnable_params order
        blob = net.g[off:off + int(np.prod(shape))]
        sumsq = F(sumsq + F(np.dot(blob, blob)))                 # Blob::sumsq_diff = caffe_cpu_dot
    l2 = F(np.sqrt(sumsq))
    clip = F(cfg.clip_gradients)
    if clip >= 0 and l2 > clip:
        net.g *= F(clip / l2)                                    # Blob::scale_diff = cblas_sscal
    b1, b2 = F(cfg.momentum), F(cfg.momentum2)
    t = it + 1
    corr = F(np.sqrt(1.0 - np.power(float(b2), t)) / (1.0 - np.power(float(b1), t)))   # std::pow promotes to double
    step = F(F(lr) * corr)
    g = net.g
    axpby(F(1) - b1, g, b1, m)
    tmp = g * g
    axpby(F(1) - b2, tmp, b2, v)
    tmp = np.sqrt(v)                                             # caffe_powx(v, 0.5)
    tmp = tmp + F(cfg.delta)
    tmp = m / tmp
    g[...] = step * tmp
    net.p -= g                                                   # Blob::Update: axpy(-1, diff, data)
    return float(l2)


def invert_gradients(a_pi, d):
    """dqn.cpp:927-957, float arithmetic."""
    lo = np.array([-1] * 4 + [0, -180, -180, -180, 0, -180], F)
    hi = np.array([1] * 4 + [100, 180, 180, 180, 100, 180], F)
    out = d.copy()
    neg, pos = d < 0, d > 0
    up = (hi - a_pi) / (hi - lo)
    dn = (a_pi - lo) / (hi - lo)
    out[neg] = (d * up)[neg]
    out[pos] = (d * dn)[pos]
    return out


def update(cfg, blobs_a, blobs_c, st, s, a10, r, mc, term, sn):
    """st: dict of float32 arrays actor, critic, actor_target, critic_target, *_m, *_v + iters (modified in place).
    Returns (critic_loss, avg_q, taps)."""
    B = s.shape[0]
    s, a10, r, mc, sn = (np.ascontiguousarray(x, F) for x in (s, a10, r, mc, sn))
    term = np.asarray(term).astype(bool)
    actor, critic = Net(blobs_a, st["actor"], False), Net(blobs_c, st["critic"], True)
    actor_t, critic_t = Net(blobs_a, st["actor_target"], False), Net(blobs_c, st["critic_target"], True)
    taps = {}
    # dqn.cpp:889-900: targets from the target nets on the next states of the non-terminal transitions
    nt = ~term
    q_next = np.zeros(B, F)
    if nt.any():
        a_t = np.concatenate(actor_t.forward(sn[nt]), axis=1)
        q_next[nt] = critic_t.forward(np.concatenate([sn[nt], a_t], axis=1))[0][:, 0]
    y = np.empty(B, F)
    for n in range(B):
        off = r[n] if term[n] else F(float(r[n]) + cfg.gamma * float(q_next[n]))          # double expression, float store
        y[n] = F(cfg.beta * float(mc[n]) + (1 - cfg.beta) * float(off))
    taps["y"], taps["q_next"] = y, q_next
    # dqn.cpp:904 critic_solver_->Step(1): ClearParamDiffs, forward, loss, backward, ApplyUpdate
    critic.g[...] = 0
    q = critic.forward(np.concatenate([s, a10], axis=1))[0][:, 0]
    diff = q - y
    loss = F(F(np.dot(diff, diff)) / F(B) / F(2))
    critic.backward([(F(1) / F(B) * diff)[:, None]])
    taps["q"], taps["critic_grad"] = q.copy(), critic.g.copy()
    taps["critic_gnorm"] = solver_apply_update(critic, st["critic_m"], st["critic_v"], st["critic_iter"], cfg.critic_lr, cfg)
    st["critic_iter"] += 1
    # dqn.cpp:908-965 actor update through the (updated) critic
    critic.g[...] = 0
    actor.g[...] = 0
    a_pi = np.concatenate(actor.forward(s), axis=1)
    q_pi = critic.forward(np.concatenate([s, a_pi], axis=1))[0][:, 0]
    avg_q = float(np.sum(q_pi.astype(np.float64)) / float(F(B)))                            # accumulate(.., 0.0) / float(n)
    d_in = critic.backward([np.full((B, 1), -1, F)], param_grads=False)
    d_raw = d_in[:, -10:]
    d_inv = invert_gradients(a_pi, d_raw)
    actor.backward([d_inv[:, :4], d_inv[:, 4:]])
    taps.update(a_pi=a_pi, q_pi=q_pi, d_raw=d_raw, d_inv=d_inv, actor_grad=actor.g.copy())
    taps["actor_gnorm"] = solver_apply_update(actor, st["actor_m"], st["actor_v"], st["actor_iter"], cfg.actor_lr, cfg)
    st["actor_iter"] += 1
    if cfg.soft_update_freq > 0 and max(st["actor_iter"], st["critic_iter"]) % cfg.soft_update_freq == 0:
        axpby(cfg.tau, st["critic"], F(1) - F(cfg.tau), st["critic_target"])
        axpby(cfg.tau, st["actor"], F(1) - F(cfg.tau), st["actor_target"])
    return float(loss), avg_q, taps

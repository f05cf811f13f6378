"""ctypes wrapper around the CPU oracle (oracle/dqn_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/dqn_oracle.h.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module.  PARITY UNPINNED:
the reference (mhauskn/dqn-hfo) ships no tests or golden vectors for this path.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdqn_oracle.so")
MAX_HIDDEN = 8


class OracleConfig(C.Structure):
    _fields_ = [
        ("state_size", C.c_int32),
        ("batch", C.c_int32),
        ("n_hidden", C.c_int32),
        ("hidden", C.c_int32 * MAX_HIDDEN),
        ("gamma", C.c_double),
        ("beta", C.c_double),
        ("tau", C.c_float),
        ("soft_update_freq", C.c_int32),
        ("actor_lr", C.c_float),
        ("critic_lr", C.c_float),
        ("momentum", C.c_float),
        ("momentum2", C.c_float),
        ("delta", C.c_float),
        ("clip_gradients", C.c_float),
        ("caffe_wasted_work", C.c_int32),
        ("use_blas", C.c_int32),
    ]


class OracleTaps(C.Structure):
    _names = ["q_next", "y", "q", "critic_grad", "a_pi", "q_pi", "d_raw", "d_inv",
              "actor_grad", "critic_gnorm", "actor_gnorm"]
    _fields_ = [(n, C.POINTER(C.c_float)) for n in _names]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (Makefile in this directory)."""
    src = os.path.join(_HERE, "dqn_oracle.c")
    hdr = os.path.join(_HERE, "dqn_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        cp = C.POINTER(OracleConfig)
        L.dqo_actor_param_count.restype = C.c_int64
        L.dqo_actor_param_count.argtypes = [cp]
        L.dqo_critic_param_count.restype = C.c_int64
        L.dqo_critic_param_count.argtypes = [cp]
        L.dqo_actor_forward.argtypes = [cp, fp, C.c_int32, fp, fp]
        L.dqo_critic_forward.argtypes = [cp, fp, C.c_int32, fp, fp, fp]
        L.dqo_update.argtypes = [cp] + [fp] * 8 + [C.POINTER(C.c_int32)] * 2 + [fp] * 4 + [
            C.POINTER(C.c_uint8), fp, fp, fp, C.POINTER(OracleTaps)]
        L.dqo_soft_update.argtypes = [C.c_int64, C.c_float, fp, fp]
        L.dqo_soft_update.restype = None
        L.dqo_label_transitions.argtypes = [C.c_int32, C.c_double, fp, fp]
        L.dqo_label_transitions.restype = None
        L.dqo_invert_gradients.argtypes = [C.c_int32, fp, fp]
        L.dqo_invert_gradients.restype = None
        L.dqo_replay_after_add_one.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.dqo_replay_after_add_many.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.dqo_get_action.argtypes = [fp, fp, fp]
        L.dqo_load_blas.argtypes = [C.c_char_p]
        L.dqo_set_threads.argtypes = [C.c_int]
        L.dqo_set_threads.restype = None
        _lib = L
    return _lib


def _f(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_config(state_size=58, batch=32, hidden=(1024, 512, 256, 128), gamma=0.99, beta=0.5,
                tau=0.001, soft_update_freq=1, actor_lr=1e-5, critic_lr=1e-3, momentum=0.95,
                momentum2=0.999, delta=1e-8, clip_gradients=10.0, caffe_wasted_work=0,
                use_blas=0) -> OracleConfig:
    """Defaults = the reference's flags (dqn.cpp:21-31, dqn_main.cpp:30-37)."""
    c = OracleConfig()
    c.state_size, c.batch, c.n_hidden = state_size, batch, len(hidden)
    for i, h in enumerate(hidden):
        c.hidden[i] = h
    c.gamma, c.beta, c.tau, c.soft_update_freq = gamma, beta, tau, soft_update_freq
    c.actor_lr, c.critic_lr, c.momentum, c.momentum2 = actor_lr, critic_lr, momentum, momentum2
    c.delta, c.clip_gradients = delta, clip_gradients
    c.caffe_wasted_work, c.use_blas = caffe_wasted_work, use_blas
    return c


def find_openblas() -> list:
    """An in-image OpenBLAS exporting cblas_sgemm (used only for the timed CPU baseline)."""
    import site
    pats = []
    for sp in site.getsitepackages():
        pats += [os.path.join(sp, "scipy.libs", "libscipy_openblas-*.so*"),          # scipy_cblas_sgemm, LP64
                 os.path.join(sp, "opencv_python_headless.libs", "libopenblas*.so*")]
    out = []
    for p in pats:
        out += sorted(glob.glob(p))
    return out


_blas_loaded = False


def load_blas() -> bool:
    """dlopen()s the first in-image OpenBLAS that exports (scipy_)cblas_sgemm with 32-bit ints."""
    global _blas_loaded
    if _blas_loaded:
        return True
    for p in find_openblas():
        if lib().dqo_load_blas(p.encode()) == 0:
            _blas_loaded = True
            return True
    return False


@dataclass
class OracleState:
    """All learner state of one dqn::DQN (dqn.hpp:183-193), Caffe param order."""
    cfg: OracleConfig
    actor: np.ndarray
    critic: np.ndarray
    actor_target: np.ndarray
    critic_target: np.ndarray
    actor_m: np.ndarray = None
    actor_v: np.ndarray = None
    critic_m: np.ndarray = None
    critic_v: np.ndarray = None
    actor_iter: int = 0
    critic_iter: int = 0
    last_taps: dict = field(default_factory=dict)

    def __post_init__(self):
        for n in ("actor", "critic", "actor_target", "critic_target"):
            setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.float32).copy())
        for n, ref in (("actor_m", self.actor), ("actor_v", self.actor),
                       ("critic_m", self.critic), ("critic_v", self.critic)):
            if getattr(self, n) is None:
                setattr(self, n, np.zeros_like(ref))
            else:
                setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.float32).copy())

    def update(self, s, act10, reward, mc, term, s_next, taps=False):
        """One UpdateActorCritic (dqn.cpp:828-972).  Returns (critic_loss, avg_q)."""
        cfg = self.cfg
        B, S = cfg.batch, cfg.state_size
        s = np.ascontiguousarray(s, np.float32).reshape(B, S)
        act10 = np.ascontiguousarray(act10, np.float32).reshape(B, 10)
        reward = np.ascontiguousarray(reward, np.float32).reshape(B)
        mc = np.ascontiguousarray(mc, np.float32).reshape(B)
        term = np.ascontiguousarray(term, np.uint8).reshape(B)
        s_next = np.ascontiguousarray(s_next, np.float32).reshape(B, S)
        ai, ci = C.c_int32(self.actor_iter), C.c_int32(self.critic_iter)
        loss, avgq = np.zeros(1, np.float32), np.zeros(1, np.float32)
        tp = None
        arrays = {}
        if taps:
            Pa, Pc = self.actor.size, self.critic.size
            shapes = dict(q_next=B, y=B, q=B, critic_grad=Pc, a_pi=B * 10, q_pi=B, d_raw=B * 10,
                          d_inv=B * 10, actor_grad=Pa, critic_gnorm=1, actor_gnorm=1)
            t = OracleTaps()
            for k, n in shapes.items():
                arrays[k] = np.zeros(n, np.float32)
                setattr(t, k, _f(arrays[k]))
            tp = C.byref(t)
        rc = lib().dqo_update(C.byref(cfg), _f(self.actor), _f(self.critic), _f(self.actor_target),
                              _f(self.critic_target), _f(self.actor_m), _f(self.actor_v),
                              _f(self.critic_m), _f(self.critic_v), C.byref(ai), C.byref(ci),
                              _f(s), _f(act10), _f(reward), _f(mc),
                              term.ctypes.data_as(C.POINTER(C.c_uint8)), _f(s_next), _f(loss),
                              _f(avgq), tp)
        if rc != 0:
            raise RuntimeError("dqo_update failed")
        self.actor_iter, self.critic_iter = ai.value, ci.value
        self.last_taps = arrays
        return float(loss[0]), float(avgq[0])

    def actor_forward(self, states, target=False):
        states = np.ascontiguousarray(states, np.float32)
        n = states.shape[0]
        out = np.zeros((n, 10), np.float32)
        w = self.actor_target if target else self.actor
        lib().dqo_actor_forward(C.byref(self.cfg), _f(w), n, _f(states), _f(out))
        return out

    def critic_forward(self, states, act10, target=False):
        states = np.ascontiguousarray(states, np.float32)
        act10 = np.ascontiguousarray(act10, np.float32)
        n = states.shape[0]
        q = np.zeros(n, np.float32)
        w = self.critic_target if target else self.critic
        lib().dqo_critic_forward(C.byref(self.cfg), _f(w), n, _f(states), _f(act10), _f(q))
        return q


def param_counts(cfg: OracleConfig):
    return int(lib().dqo_actor_param_count(C.byref(cfg))), int(lib().dqo_critic_param_count(C.byref(cfg)))


def net_blobs(cfg: OracleConfig, critic: bool):
    """[(name, offset, shape)] in Caffe learnable_params order (dqn.cpp:418-454)."""
    out, off = [], 0
    k = cfg.state_size + (10 if critic else 0)
    for l in range(cfg.n_hidden):
        h = cfg.hidden[l]
        out.append((f"ip{l+1}.W", off, (h, k))); off += h * k
        out.append((f"ip{l+1}.b", off, (h,))); off += h
        k = h
    heads = [("q_values_layer", 1)] if critic else [("action_layer", 4), ("actionpara_layer", 6)]
    for name, n in heads:
        out.append((f"{name}.W", off, (n, k))); off += n * k
        out.append((f"{name}.b", off, (n,))); off += n
    return out


def init_params(cfg: OracleConfig, critic: bool, rng: np.random.Generator, mode="caffe"):
    """Weight sets for tests: 'caffe' = gaussian std 0.01, b=0 (dqn.cpp:350-352; Caffe's own RNG is
    not reproducible, tests inject weights); 'warm' = 1/sqrt(fan_in) scale with small biases."""
    n = param_counts(cfg)[1 if critic else 0]
    p = np.zeros(n, np.float32)
    for name, off, shape in net_blobs(cfg, critic):
        cnt = int(np.prod(shape))
        if name.endswith(".W"):
            std = 0.01 if mode == "caffe" else 1.0 / np.sqrt(shape[1])
            p[off:off + cnt] = rng.normal(0, std, cnt).astype(np.float32)
        elif mode != "caffe":
            p[off:off + cnt] = rng.normal(0, 0.05, cnt).astype(np.float32)
    return p


def label_transitions(reward, gamma=0.99):
    reward = np.ascontiguousarray(reward, np.float32)
    mc = np.zeros_like(reward)
    lib().dqo_label_transitions(reward.size, gamma, _f(reward), _f(mc))
    return mc


def invert_gradients(a_pi, d10):
    a_pi = np.ascontiguousarray(a_pi, np.float32)
    d = np.ascontiguousarray(d10, np.float32).copy()
    lib().dqo_invert_gradients(a_pi.shape[0], _f(a_pi), _f(d))
    return d


def get_action(out10):
    o = np.ascontiguousarray(out10, np.float32)
    a1, a2 = C.c_float(), C.c_float()
    idx = lib().dqo_get_action(_f(o), C.byref(a1), C.byref(a2))
    return idx, a1.value, a2.value


def synth_batch(cfg: OracleConfig, rng: np.random.Generator, p_term=0.1):
    """Synthetic minibatch shaped like SURVEY §8(d): states U(-1,1), act10 per
    GetRandomActorOutput ranges (dqn.cpp:664-682), rewards N(0,0.1) with sparse +5 spikes."""
    B, S = cfg.batch, cfg.state_size
    s = rng.uniform(-1, 1, (B, S)).astype(np.float32)
    sn = rng.uniform(-1, 1, (B, S)).astype(np.float32)
    a = np.empty((B, 10), np.float32)
    a[:, 0:4] = rng.uniform(-1, 1, (B, 4))
    a[:, 4] = rng.uniform(-100, 100, B)
    a[:, 5:8] = rng.uniform(-180, 180, (B, 3))
    a[:, 8] = rng.uniform(0, 100, B)
    a[:, 9] = rng.uniform(-180, 180, B)
    term = (rng.uniform(size=B) < p_term).astype(np.uint8)
    r = rng.normal(0, 0.1, B).astype(np.float32)
    r[(term == 1) & (rng.uniform(size=B) < 0.3)] += 5.0
    mc = (r + rng.normal(0, 0.5, B)).astype(np.float32)
    return s, a, r, mc, term, sn

"""ctypes binding of include/dqn_b200.h (one method per C entry point, same names minus `dqnb_`)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_HIDDEN = 8
ACTOR, CRITIC, ACTOR_TARGET, CRITIC_TARGET = 0, 1, 2, 3
GEMM_TCGEN05_3XTF32, GEMM_SIMT_FP32 = 0, 1


class Config(C.Structure):
    """dqnb_config (include/dqn_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("state_size", C.c_int32),
        ("batch", C.c_int32), ("n_hidden", C.c_int32), ("hidden", C.c_int32 * MAX_HIDDEN),
        ("replay_capacity", C.c_int32), ("max_act_batch", C.c_int32),
        ("gamma", C.c_double), ("beta", C.c_double), ("tau", C.c_float),
        ("soft_update_freq", C.c_int32), ("actor_lr", C.c_float), ("critic_lr", C.c_float),
        ("momentum", C.c_float), ("momentum2", C.c_float), ("delta", C.c_float),
        ("clip_gradients", C.c_float), ("seed", C.c_uint64), ("gemm_mode", C.c_int32),
        ("use_graph", C.c_int32), ("world_size", C.c_int32), ("rank", C.c_int32),
    ]


def lib_path() -> str:
    return os.path.join(_HERE, "libdqn_b200.so")


def build_library(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a (csrc/Makefile); cross-compiles without a GPU."""
    args = ["make", "-C", os.path.join(_HERE, "csrc")]
    if force:
        args.append("-B")
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)
    return lib_path()


_lib = None


def lib():
    """Loads libdqn_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                           "dqn-hfo_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(p)
    fp, ip, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    H = C.c_void_p
    L.dqnb_last_error.restype = C.c_char_p
    L.dqnb_version.restype = C.c_char_p
    L.dqnb_default_config.argtypes = [C.POINTER(Config)]
    L.dqnb_default_config.restype = None
    L.dqnb_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.dqnb_destroy.argtypes = [H]
    L.dqnb_param_count.argtypes = [H, C.c_int]
    L.dqnb_param_count.restype = C.c_int64
    L.dqnb_set_params.argtypes = [H, C.c_int, fp]
    L.dqnb_get_params.argtypes = [H, C.c_int, fp]
    L.dqnb_init_params.argtypes = [H, C.c_uint64, C.c_float]
    L.dqnb_clone_targets.argtypes = [H]
    L.dqnb_copy_shared_layers.argtypes = [H, H, C.c_int32, C.c_int32]
    L.dqnb_set_opt_state.argtypes = [H, C.c_int, fp, fp, C.c_int32]
    L.dqnb_get_opt_state.argtypes = [H, C.c_int, fp, fp, ip]
    L.dqnb_iters.argtypes = [H, ip, ip]
    L.dqnb_add_transitions.argtypes = [H, C.c_int32, fp, fp, fp, fp, fp, u8p]
    L.dqnb_add_transition.argtypes = [H, fp, fp, C.c_float, C.c_float, fp, C.c_uint8]
    L.dqnb_memory_size.argtypes = [H]
    L.dqnb_clear_memory.argtypes = [H]
    L.dqnb_get_transitions.argtypes = [H, C.c_int32, C.c_int32, fp, fp, fp, fp, fp, u8p]
    L.dqnb_update.argtypes = [H, C.c_int32, fp, fp]
    L.dqnb_update_with_indices.argtypes = [H, ip, fp, fp]
    L.dqnb_update_async.argtypes = [H, C.c_int32, C.POINTER(C.c_int64)]
    L.dqnb_results.argtypes = [H, C.c_int64, C.c_int32, fp, fp]
    L.dqnb_benchmark.argtypes = [H, C.c_int32, fp]
    L.dqnb_benchmark_gemms.argtypes = [H, C.c_int32, fp, ip]
    L.dqnb_peek_sample_indices.argtypes = [H, ip]
    L.dqnb_select_actions.argtypes = [H, C.c_int32, fp, fp]
    L.dqnb_select_actions_async.argtypes = [H, C.c_int32, fp]
    L.dqnb_select_actions_wait.argtypes = [H, C.c_int32, fp]
    L.dqnb_evaluate.argtypes = [H, C.c_int32, fp, fp, fp]
    L.dqnb_comm_unique_id.argtypes = [C.c_void_p]
    L.dqnb_comm_init.argtypes = [H, C.c_void_p]
    L.dqnb_comm_p2p_handle.argtypes = [H, C.c_void_p]
    L.dqnb_comm_p2p_init.argtypes = [H, C.c_void_p]
    L.dqnb_comm_status.argtypes = [H]
    L.dqnb_sync.argtypes = [H]
    L.dqnb_kernel_launches.argtypes = [H]
    L.dqnb_kernel_launches.restype = C.c_int64
    L.dqnb_debug_read.argtypes = [H, C.c_char_p, fp, C.c_int64]
    L.dqnb_debug_read.restype = C.c_int64
    L.dqnb_gemm_test.argtypes = [C.c_int] * 8 + [fp, fp, fp, fp]
    _lib = L
    return L


EXPORTS = [
    "dqnb_default_config", "dqnb_last_error", "dqnb_version", "dqnb_create", "dqnb_destroy",
    "dqnb_param_count", "dqnb_set_params", "dqnb_get_params", "dqnb_init_params", "dqnb_clone_targets", "dqnb_copy_shared_layers",
    "dqnb_set_opt_state", "dqnb_get_opt_state", "dqnb_iters", "dqnb_add_transitions",
    "dqnb_add_transition", "dqnb_memory_size", "dqnb_clear_memory", "dqnb_get_transitions",
    "dqnb_update", "dqnb_update_async", "dqnb_results", "dqnb_update_with_indices", "dqnb_benchmark", "dqnb_benchmark_gemms",
    "dqnb_peek_sample_indices",
    "dqnb_select_actions", "dqnb_select_actions_async", "dqnb_select_actions_wait", "dqnb_evaluate",
    "dqnb_comm_unique_id", "dqnb_comm_init", "dqnb_comm_p2p_handle", "dqnb_comm_p2p_init", "dqnb_comm_status",
    "dqnb_sync", "dqnb_kernel_launches", "dqnb_debug_read",
    "dqnb_gemm_test",
]


def _f(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _err():
    return lib().dqnb_last_error().decode()


def _check(rc):
    if rc != 0:
        raise RuntimeError("libdqn_b200: " + _err())


def default_config(**kw) -> Config:
    c = Config()
    lib().dqnb_default_config(C.byref(c))
    hidden = kw.pop("hidden", None)
    if hidden is not None:
        c.n_hidden = len(hidden)
        for i, h in enumerate(hidden):
            c.hidden[i] = h
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().dqnb_comm_unique_id(buf))
    return buf.raw


def gemm_test(mode, a_mn, b_mn, M, N, K, splits, A, B, device=0):
    """C = A·B^T-like product through the layer GEMM kernels; returns (C, ms)."""
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    Cm = np.zeros((M, N), np.float32)
    ms = C.c_float()
    _check(lib().dqnb_gemm_test(device, mode, a_mn, b_mn, M, N, K, splits, _f(A), _f(B), _f(Cm),
                                C.byref(ms)))
    return Cm, ms.value


class DQNB:
    """Thin object wrapper over a dqnb_handle (mirrors dqn::DQN's hot-path members)."""

    def __init__(self, **kw):
        self.cfg = default_config(**kw)
        self._h = C.c_void_p()
        _check(lib().dqnb_create(C.byref(self.cfg), C.byref(self._h)))
        self.B, self.S = self.cfg.batch, self.cfg.state_size

    def close(self):
        if self._h:
            lib().dqnb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- parameters / solver state -------------------------------------------------------------
    def param_count(self, net):
        return int(lib().dqnb_param_count(self._h, net))

    def set_params(self, net, p):
        p = np.ascontiguousarray(p, np.float32)
        assert p.size == self.param_count(net)
        _check(lib().dqnb_set_params(self._h, net, _f(p)))

    def get_params(self, net):
        p = np.zeros(self.param_count(net), np.float32)
        _check(lib().dqnb_get_params(self._h, net, _f(p)))
        return p

    def init_params(self, seed=2, std=0.01):
        _check(lib().dqnb_init_params(self._h, seed, std))

    def clone_targets(self):
        _check(lib().dqnb_clone_targets(self._h))

    def copy_shared_layers_from(self, src, n_actor_layers, n_critic_layers):
        """ShareParameters write-through: the first n layers (and their target-net twins) of `src` overwrite ours."""
        _check(lib().dqnb_copy_shared_layers(self._h, src._h, n_actor_layers, n_critic_layers))

    def set_opt_state(self, net, m, v, it):
        m = np.ascontiguousarray(m, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        _check(lib().dqnb_set_opt_state(self._h, net, _f(m), _f(v), it))

    def get_opt_state(self, net):
        n = self.param_count(net)
        m, v, it = np.zeros(n, np.float32), np.zeros(n, np.float32), C.c_int32()
        _check(lib().dqnb_get_opt_state(self._h, net, _f(m), _f(v), C.byref(it)))
        return m, v, it.value

    def iters(self):
        a, c = C.c_int32(), C.c_int32()
        _check(lib().dqnb_iters(self._h, C.byref(a), C.byref(c)))
        return a.value, c.value

    # --- replay memory -------------------------------------------------------------------------
    def add_transitions(self, s, act10, reward, mc, s_next, term):
        s = np.ascontiguousarray(s, np.float32)
        n = s.shape[0]
        act10 = np.ascontiguousarray(act10, np.float32)
        reward = np.ascontiguousarray(reward, np.float32)
        mc = np.ascontiguousarray(mc, np.float32)
        s_next = np.ascontiguousarray(s_next, np.float32)
        term = np.ascontiguousarray(term, np.uint8)
        _check(lib().dqnb_add_transitions(self._h, n, _f(s), _f(act10), _f(reward), _f(mc), _f(s_next),
                                          term.ctypes.data_as(C.POINTER(C.c_uint8))))

    def add_transition(self, s, act10, reward, mc, s_next, term):
        s = np.ascontiguousarray(s, np.float32)
        act10 = np.ascontiguousarray(act10, np.float32)
        s_next = np.ascontiguousarray(s_next, np.float32)
        _check(lib().dqnb_add_transition(self._h, _f(s), _f(act10), float(reward), float(mc), _f(s_next),
                                         int(term)))

    def memory_size(self):
        return int(lib().dqnb_memory_size(self._h))

    def clear_memory(self):
        _check(lib().dqnb_clear_memory(self._h))

    def get_transitions(self, first, n):
        S = self.S
        s, a = np.zeros((n, S), np.float32), np.zeros((n, 10), np.float32)
        r, mc = np.zeros(n, np.float32), np.zeros(n, np.float32)
        sn, t = np.zeros((n, S), np.float32), np.zeros(n, np.uint8)
        _check(lib().dqnb_get_transitions(self._h, first, n, _f(s), _f(a), _f(r), _f(mc), _f(sn),
                                          t.ctypes.data_as(C.POINTER(C.c_uint8))))
        return s, a, r, mc, sn, t

    # --- learning ------------------------------------------------------------------------------
    def update(self, n=1):
        loss, avgq = np.zeros(n, np.float32), np.zeros(n, np.float32)
        _check(lib().dqnb_update(self._h, n, _f(loss), _f(avgq)))
        return loss, avgq

    def update_async(self, n=1):
        """Enqueue n updates; returns the 1-based sequence number of the last one (see results())."""
        last = C.c_int64(0)
        _check(lib().dqnb_update_async(self._h, n, C.byref(last)))
        return int(last.value)

    def results(self, first_step, n=1):
        loss, avgq = np.zeros(n, np.float32), np.zeros(n, np.float32)
        _check(lib().dqnb_results(self._h, first_step, n, _f(loss), _f(avgq)))
        return loss, avgq

    def update_with_indices(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        assert idx.size == self.B
        loss, avgq = np.zeros(1, np.float32), np.zeros(1, np.float32)
        _check(lib().dqnb_update_with_indices(self._h, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                              _f(loss), _f(avgq)))
        return float(loss[0]), float(avgq[0])

    def benchmark(self, n):
        ms = C.c_float()
        _check(lib().dqnb_benchmark(self._h, n, C.byref(ms)))
        return ms.value

    def peek_sample_indices(self):
        idx = np.zeros(self.B, np.int32)
        _check(lib().dqnb_peek_sample_indices(self._h, idx.ctypes.data_as(C.POINTER(C.c_int32))))
        return idx

    # --- acting --------------------------------------------------------------------------------
    def select_actions(self, states):
        states = np.ascontiguousarray(states, np.float32)
        n = states.shape[0]
        out = np.zeros((n, 10), np.float32)
        _check(lib().dqnb_select_actions(self._h, n, _f(states), _f(out)))
        return out

    def evaluate(self, states, act10):
        states = np.ascontiguousarray(states, np.float32)
        act10 = np.ascontiguousarray(act10, np.float32)
        n = states.shape[0]
        q = np.zeros(n, np.float32)
        _check(lib().dqnb_evaluate(self._h, n, _f(states), _f(act10), _f(q)))
        return q

    # --- misc ----------------------------------------------------------------------------------
    def comm_init(self, id128: bytes):
        buf = C.create_string_buffer(id128, 128)
        _check(lib().dqnb_comm_init(self._h, buf))

    def comm_p2p_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().dqnb_comm_p2p_handle(self._h, buf))
        return buf.raw

    def comm_p2p_init(self, handles: bytes):
        buf = C.create_string_buffer(handles, len(handles))
        _check(lib().dqnb_comm_p2p_init(self._h, buf))

    def comm_status(self) -> int:
        return int(lib().dqnb_comm_status(self._h))

    def sync(self):
        _check(lib().dqnb_sync(self._h))

    def kernel_launches(self):
        return int(lib().dqnb_kernel_launches(self._h))

    def debug_read(self, name, count):
        out = np.zeros(count, np.float32)
        n = lib().dqnb_debug_read(self._h, name.encode(), _f(out), count)
        if n < 0:
            raise RuntimeError(f"debug_read({name}) failed: {_err()}")
        return out[:n]

"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/dqn_b200.h declares,
the ctypes mirror of dqnb_config matches the C struct, and create() fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import ROOT, pkg


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dqn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dqnb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    P = pkg()
    L = P.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"libdqn_b200.so does not export {s}"
    from dqn_hfo_b200 import binding
    assert sorted(binding.EXPORTS) == syms


def test_default_config_matches_reference_flags():
    P = pkg()
    from dqn_hfo_b200 import binding
    c = binding.default_config()
    assert c.struct_size == C.sizeof(binding.Config)
    # dqn.cpp:21-31, dqn_main.cpp:30-37, dqn.hpp:19
    assert (c.state_size, c.batch, c.n_hidden) == (58, 32, 4)
    assert list(c.hidden)[:4] == [1024, 512, 256, 128]
    assert c.replay_capacity == 500000 and c.soft_update_freq == 1
    assert c.gamma == 0.99 and c.beta == 0.5
    assert c.tau == np.float32(0.001) and c.actor_lr == np.float32(1e-5) and c.critic_lr == np.float32(1e-3)
    assert c.momentum == np.float32(0.95) and c.momentum2 == np.float32(0.999)
    assert c.clip_gradients == 10.0 and c.delta == np.float32(1e-8)
    assert c.gemm_mode == P.GEMM_TCGEN05_3XTF32


def test_create_fails_loudly_without_gpu_or_with_bad_config():
    import torch
    P = pkg()
    from dqn_hfo_b200 import binding
    bad = binding.default_config()
    bad.struct_size = 4
    h = C.c_void_p()
    assert P.lib().dqnb_create(C.byref(bad), C.byref(h)) != 0
    assert b"struct_size" in P.lib().dqnb_last_error()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback|cuda"):
            P.DQNB()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The product GEMM must be tcgen05 + TMA (UTCHMMA / UTMALDG / LDTM in SASS), not mma.sync."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    P = pkg()
    sass = subprocess.run(["cuobjdump", "-sass", P.lib_path()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA.16" not in sass

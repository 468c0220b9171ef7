"""The C-ABI shared library loads and exports every symbol include/hmpc.h declares; the ctypes mirror of
hmpc_config matches the C struct; without a CUDA device creation fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from hopper_mpc_inertial_b200 import _lib
from tests.conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "hmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    assert sorted(_lib.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.hmpc_abi_version() == _lib.HMPC_ABI_VERSION


def test_config_struct_mirror_matches_c_defaults():
    from tests.emul import default_config
    cfg = _lib.default_config()                       # filled by the C library (no GPU needed)
    ref = default_config()                            # restated in Python
    for name, _ in _lib.HmpcConfig._fields_:
        a, b = getattr(cfg, name), getattr(ref, name)
        if hasattr(a, "__len__"):
            np.testing.assert_allclose(list(a), list(b), rtol=1e-12, atol=0, err_msg=name)
        else:
            assert a == pytest.approx(b, rel=1e-12), name
    # the trailing field is where a layout mismatch would show up
    assert cfg.ipm_tol == 1e-6 and cfg.polish_tol == 1e-9 and cfg.kkt_eps == 1e-9
    assert C.sizeof(_lib.HmpcConfig) == 22 * 4 + (5 + 9 + 9 + 3 + 3 + 3 + 5 + 3) * 8     # 21 int32 + padding


def test_bad_arguments_are_rejected_without_a_device_call():
    lib = _lib.load()
    cfg = _lib.default_config()
    h = C.c_void_p()
    cfg.N = 1
    assert lib.hmpc_create(C.byref(cfg), C.byref(h)) == -1 and h.value is None
    assert b"N must be" in lib.hmpc_last_error()
    cfg = _lib.default_config(); cfg.abi_version = 99
    assert lib.hmpc_create(C.byref(cfg), C.byref(h)) == -1
    cfg = _lib.default_config(); cfg.precision = 7
    assert lib.hmpc_create(C.byref(cfg), C.byref(h)) == -1          # unknown precision
    assert lib.hmpc_create(None, C.byref(h)) == -1
    assert lib.hmpc_destroy(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    lib = _lib.load()
    cfg = _lib.default_config()
    h = C.c_void_p()
    assert lib.hmpc_create(C.byref(cfg), C.byref(h)) == -4            # HMPC_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.hmpc_last_error()
    from hopper_mpc_inertial_b200.batch import BatchMpc
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchMpc(4)
    from hopper_mpc_inertial_b200 import mpc_cvx_euler_3f
    mpc = mpc_cvx_euler_3f.Mpc(t=0.02, N=10, m=7.5, g=9.807, mu=1, Jinv=np.eye(3), rh=np.zeros(3))
    with pytest.raises(RuntimeError):
        mpc.mpcontrol(np.zeros(12), np.zeros((10, 12)), np.zeros((10, 3)), np.ones(10), True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hopper_mpc_inertial_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "tests.emul" not in txt, f

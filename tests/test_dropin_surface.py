"""Drop-in surface of the Python modules: utils, Mpc attributes, gain validation, the CLI shims.
No GPU: nothing here creates a device handle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from .conftest import ROOT, golden
from hopper_mpc_inertial_b200 import mpc_cvx_euler_2f, mpc_cvx_euler_3f, utils
from hopper_mpc_inertial_b200.robotrunner import convert
from oracle import hopper_oracle as ho


def _rand_quats(n, seed=3):
    q = np.random.default_rng(seed).normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def test_utils_exports_the_reference_names():
    # utils.py:4-70 -- robotrunner.py:7 does `from utils import H, L, R, quat2euler`
    for name in ("H", "T", "projection", "hat", "L", "R", "rz", "quat2euler", "quat2rot"):
        assert hasattr(utils, name), name
    assert utils.H.shape == (4, 3) and utils.T.shape == (4, 4)
    assert np.array_equal(utils.T, np.diag([1.0, -1.0, -1.0, -1.0]))


def test_utils_match_the_reference_module(reference):
    ru = reference.utils
    rng = np.random.default_rng(0)
    assert np.array_equal(utils.H, ru.H) and np.array_equal(utils.T, ru.T)
    for q in _rand_quats(16):
        assert np.array_equal(utils.L(q), ru.L(q))
        assert np.array_equal(utils.R(q), ru.R(q))
        assert np.array_equal(utils.quat2rot(q), ru.quat2rot(q))
    for _ in range(8):
        w = rng.normal(size=3)
        assert np.array_equal(utils.hat(w), ru.hat(w))
        phi = rng.uniform(-4, 4)
        assert np.array_equal(utils.rz(phi), ru.rz(phi))
        p0, v = rng.normal(size=3) + [0, 0, 2.0], rng.normal(size=3) - [0, 0, 2.0]
        assert np.array_equal(utils.projection(p0, v), ru.projection(p0, v))


def test_utils_algebra():
    qs = _rand_quats(8)
    for q, p in zip(qs, _rand_quats(8, seed=5)):
        # L(q) p = q * p and R(q) p = p * q (Hamilton product)
        def qmul(a, b):
            return np.concatenate([[a[0] * b[0] - a[1:] @ b[1:]], a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:])])
        np.testing.assert_allclose(utils.L(q) @ p, qmul(q, p), atol=1e-15)
        np.testing.assert_allclose(utils.R(q) @ p, qmul(p, q), atol=1e-15)
        # rotation matrix two ways (robotrunner.py:22 vs utils.py:65-70)
        Rm = utils.H.T @ utils.L(q) @ utils.R(q).T @ utils.H
        np.testing.assert_allclose(Rm, utils.quat2rot(q), atol=1e-14)
        # Euler angles reproduce the rotation: R = Rz(yaw) Ry(pitch) Rx(roll)
        r, pch, y = utils.quat2euler(q)
        Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
        Ry = np.array([[np.cos(pch), 0, np.sin(pch)], [0, 1, 0], [-np.sin(pch), 0, np.cos(pch)]])
        np.testing.assert_allclose(utils.rz(y).T @ Ry @ Rx, Rm, atol=1e-13)
        np.testing.assert_allclose(utils.quat2euler(q), ho.quat2euler(q), atol=1e-15)
    v = np.array([0.3, -1.2, 2.0])
    np.testing.assert_allclose(utils.hat(v) @ qs[0, 1:], np.cross(v, qs[0, 1:]), atol=1e-15)
    # gimbal lock: pitch = +-90 deg takes the fallback branch (yaw = 0)
    q = np.array([np.cos(np.pi / 4), 0.0, np.sin(np.pi / 4), 0.0])
    e = utils.quat2euler(q)
    assert abs(e[1] - np.pi / 2) < 1e-7 and e[2] == 0.0


def test_host_convert_matches_reference_golden():
    g = golden("sim.npz")
    X, x = np.atleast_2d(g["X"]), np.atleast_2d(g["x"])      # the reference's convert() output (make_golden.py)
    for Xi, xi in zip(X, x):
        np.testing.assert_allclose(convert(Xi), xi, rtol=0, atol=1e-13)


@pytest.mark.parametrize("mod,dyn", [(mpc_cvx_euler_3f, "3f"), (mpc_cvx_euler_2f, "2f")])
def test_mpc_constructor_attributes(mod, dyn):
    # mpc_cvx_euler_3f.py:12-39 / mpc_cvx_euler_2f.py:12-38
    prm = ho.Params(dyn=dyn, N=7)
    mpc = mod.Mpc(t=0.02, N=7, m=7.5, g=9.807, mu=1, Jinv=prm.Jinv, rh=prm.rh)
    assert (mpc.n_x, mpc.n_u) == (12, 6)
    assert mpc.A.shape == (12, 12) and np.array_equal(mpc.A[0:3, 6:9], np.eye(3)) and mpc.A.sum() == 3
    assert mpc.B.shape == (12, 6)
    if dyn == "3f":
        assert np.array_equal(mpc.B[6:9, 0:3], np.eye(3) / 7.5)
    else:
        assert not mpc.B.any()
    assert mpc.G[8] == -9.807 and np.array_equal(mpc.Gd, mpc.G * 0.02)
    assert mpc.Ad.shape == (7, 12, 12) and mpc.Bd.shape == (7, 12, 6)
    assert np.array_equal(np.diag(mpc.Q), [50., 50., 2., 1., 1., 50., 1., 1., 1., 10., 10., 10.])
    assert np.array_equal(mpc.R, np.eye(6) * 0.001)
    assert np.array_equal(mpc.f_max, [352, 0, 206]) and np.array_equal(mpc.f_min, -mpc.f_max)
    assert mpc.x.shape == (8, 12) and mpc.u.shape == (7, 6) and mpc.x.value is None


def test_reference_constructor_agrees(reference):
    prm = ho.Params(dyn="3f", N=5)
    for mod, rmod in ((mpc_cvx_euler_3f, reference.mpc3f), (mpc_cvx_euler_2f, reference.mpc2f)):
        a = mod.Mpc(t=0.02, N=5, m=7.5, g=9.807, mu=1, Jinv=prm.Jinv, rh=prm.rh)
        b = rmod.Mpc(t=0.02, N=5, m=7.5, g=9.807, mu=1, Jinv=prm.Jinv, rh=prm.rh)
        for name in ("A", "B", "G", "Ad", "Bd", "Gd", "Q", "R", "f_max", "f_min"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_non_diagonal_gains_are_rejected_at_assignment():
    prm = ho.Params(N=5)
    mpc = mpc_cvx_euler_3f.Mpc(t=0.02, N=5, m=7.5, g=9.807, mu=1, Jinv=prm.Jinv, rh=prm.rh)
    Q = np.eye(12); Q[0, 1] = Q[1, 0] = 0.1
    with pytest.raises(ValueError, match="diagonal"):
        mpc.Q = Q
    with pytest.raises(ValueError, match="6x6"):
        mpc.R = np.eye(5)
    mpc.Q = np.eye(12) * 3.0                      # a plain diagonal replacement is fine
    assert mpc.Q[4, 4] == 3.0
    mpc.Q[0, 1] = 0.5                              # in-place edits are caught when the gains are pushed
    with pytest.raises(ValueError, match="diagonal"):
        mpc._push_gains(None)


def test_top_level_run_py_is_the_cli():
    # run.py:7-15: positional dyn, --curve, --N_run (and the README's --runtime)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "run.py"), "--help"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0
    for flag in ("{2f,3f}", "--curve", "--N_run", "--runtime"):
        assert flag in out.stdout

"""Import the reference's own modules with stand-ins for its missing third-party imports.

*** TEST INFRASTRUCTURE ONLY ***  Usable only where /root/reference exists (the build container);
nothing under ``tests -m gpu``, ``smoke()`` or ``bench.py`` may call it.  It is used by
``oracle/make_golden.py`` to freeze reference outputs into ``tests/golden/`` and by the CPU tests
that are skipped when the reference tree is absent.

Stand-ins: ``cvxpy`` -> oracle.minicvx (so mpc_cvx_euler_*.build_qp/solve_qp run
unchanged), ``transforms3d`` -> restated ``quat2euler(axes='rzyx')`` (oracle.hopper_oracle),
``plots`` -> no-op functions (plots.py:6-13 imports matplotlib at module load and
``Runner.run()`` blocks on ``plt.show()``, robotrunner.py:93,117-122).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(REF_SRC)


def load():
    """Returns a namespace with the reference modules: utils, robotrunner, mpc3f, mpc2f."""
    if not available():
        raise RuntimeError("reference tree not present")
    from . import minicvx, hopper_oracle

    sys.modules["cvxpy"] = minicvx
    t3d = types.ModuleType("transforms3d")
    t3d.euler = types.ModuleType("transforms3d.euler")

    def quat2euler(q, axes="rzyx"):
        assert axes == "rzyx"
        r, p, y = hopper_oracle.quat2euler(np.asarray(q, float))
        return (y, p, r)
    t3d.euler.quat2euler = quat2euler
    sys.modules["transforms3d"] = t3d
    sys.modules["transforms3d.euler"] = t3d.euler

    # plots.py (visualisation, out of scope) imports matplotlib at module load and blocks on
    # plt.show(); replace the whole module by no-op functions.
    plots = types.ModuleType("plots")
    for fn in ("fplot", "posplot", "posplot_animate", "posplot_animate_cube"):
        setattr(plots, fn, lambda *a, **k: None)
    sys.modules["plots"] = plots
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    # the reference's flat module names must not shadow / be shadowed by the product package
    mods = {}
    for name in ("utils", "mpc_cvx_euler_3f", "mpc_cvx_euler_2f", "robotrunner"):
        sys.modules.pop(name, None)
        mods[name] = importlib.import_module(name)
    mods["plots"] = plots
    ns = types.SimpleNamespace(utils=mods["utils"], robotrunner=mods["robotrunner"],
                               mpc3f=mods["mpc_cvx_euler_3f"], mpc2f=mods["mpc_cvx_euler_2f"],
                               plots=mods["plots"], minicvx=minicvx)
    return ns

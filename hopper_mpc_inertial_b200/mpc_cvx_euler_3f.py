"""Drop-in for the reference's mpc_cvx_euler_3f (world-frame force MPC, mpc_cvx_euler_3f.py:10-160)."""
from ._mpc_common import MpcBase


class Mpc(MpcBase):
    DYN = "3f"

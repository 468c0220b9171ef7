"""Drop-in for the reference's mpc_cvx_euler_2f (body-frame force, fy == 0; mpc_cvx_euler_2f.py:10-158)."""
from ._mpc_common import MpcBase


class Mpc(MpcBase):
    DYN = "2f"

"""Small host-side math used by the drop-in wrappers (reference: utils.py:4-62).

Only what the host side needs: the product's arithmetic runs on the GPU (csrc/hmpc_sim.cuh)."""
import numpy as np

H = np.zeros((4, 3))
H[1:4, 0:3] = np.eye(3)


def hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def rz(phi):
    c, s = np.cos(phi), np.sin(phi)
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])

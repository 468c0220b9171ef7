"""Host-side drop-in for the reference's ``utils`` module (utils.py:1-70): same names, same argument meaning,
same results -- ``H``, ``T``, ``projection``, ``hat``, ``L``, ``R``, ``rz``, ``quat2euler``, ``quat2rot``.

The product's arithmetic runs on the GPU (csrc/hmpc_sim.cuh holds the device versions of ``hat``, ``L``, ``R``,
``rz``, ``quat2euler``); these numpy versions exist so that code written against the reference
(``from utils import H, L, R, quat2euler``, robotrunner.py:7) keeps working.  ``quat2euler`` restates
transforms3d's ``quat2euler(Q, axes='rzyx')`` (not installed here), including its gimbal-lock
branch, and returns roll-pitch-yaw order like utils.py:54-62.
"""
import numpy as np

# quaternion "vector part" selector: v4 = H @ v3 (utils.py:4-5)
H = np.vstack([np.zeros((1, 3)), np.eye(3)])
# quaternion conjugation as a matrix (utils.py:7-8)
T = np.diag([1.0, -1.0, -1.0, -1.0])

_EPS4 = 4.0 * np.finfo(float).eps


def projection(p0, v):
    """Point where the ray p0 + s v meets the ground plane z = 0 (utils.py:11-18)."""
    s = (0 - p0[2]) / v[2]
    return np.array([p0[0] + s * v[0], p0[1] + s * v[1], 0])


def hat(w):
    """Skew-symmetric matrix: hat(w) @ x == cross(w, x) (utils.py:21-25)."""
    wx, wy, wz = w[0], w[1], w[2]
    return np.array([[0, -wz, wy], [wz, 0, -wx], [-wy, wx, 0]])


def _quat_matrix(Q, sign):
    """[[s, -v'], [v, s I + sign hat(v)]] for Q = (s, v): left (+1) / right (-1) multiplication matrix."""
    M = np.zeros((4, 4))
    M[0, 0] = Q[0]
    M[0, 1:] = -np.asarray(Q[1:4])
    M[1:, 0] = Q[1:4]
    M[1:, 1:] = Q[0] * np.eye(3) + sign * hat(Q[1:4])
    return M


def L(Q):
    """Left quaternion multiplication matrix: L(q) p == q * p (utils.py:28-34)."""
    return _quat_matrix(Q, 1.0)


def R(Q):
    """Right quaternion multiplication matrix: R(q) p == p * q (utils.py:37-43)."""
    return _quat_matrix(Q, -1.0)


def rz(phi):
    """World -> body yaw rotation Rz(phi)' used by the linearisation (utils.py:46-51)."""
    c, s = np.cos(phi), np.sin(phi)
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])


def quat2rot(Q):
    """Rotation matrix of a unit quaternion (w, x, y, z) (utils.py:65-70)."""
    w, x, y, z = Q
    return np.array([[2 * (w ** 2 + x ** 2) - 1, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 2 * (w ** 2 + y ** 2) - 1, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 2 * (w ** 2 + z ** 2) - 1]])


def quat2euler(Q):
    """ZYX Euler angles of quaternion (w, x, y, z), returned as [roll, pitch, yaw] (utils.py:54-62).

    transforms3d path restated: quat2mat (normalising by |Q|^2) then mat2euler(axes='rzyx'), i.e. static 'sxyz'
    extraction with the gimbal-lock fallback (yaw = 0) when hypot(M00, M10) <= 4 eps."""
    w, x, y, z = (float(Q[0]), float(Q[1]), float(Q[2]), float(Q[3]))
    nq = w * w + x * x + y * y + z * z
    if nq < np.finfo(float).eps:
        return np.zeros(3)          # transforms3d returns the identity rotation here
    s = 2.0 / nq
    m00 = 1.0 - s * (y * y + z * z)
    m10 = s * (x * y + w * z)
    m20 = s * (x * z - w * y)
    m21 = s * (y * z + w * x)
    m22 = 1.0 - s * (x * x + y * y)
    cy = np.hypot(m00, m10)
    xyz = np.zeros(3)
    if cy > _EPS4:
        xyz[0] = np.arctan2(m21, m22)
        xyz[1] = np.arctan2(-m20, cy)
        xyz[2] = np.arctan2(m10, m00)
    else:
        m11 = 1.0 - s * (x * x + z * z)
        m12 = s * (y * z - w * x)
        xyz[0] = np.arctan2(-m12, m11)
        xyz[1] = np.arctan2(-m20, cy)
        xyz[2] = 0.0
    return xyz

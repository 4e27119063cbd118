"""Minimal float64 GLM-style matrix helpers used to build scene fixtures.

The reference builds its matrices with an un-vendored `glm` module
(nim.cfg:1 -> ../nim-glm-fork).  Its conventions are pinned by the golden ray of
test/boxtest.nim:32-33 (see tests/test_oracle_golden.py): column vectors,
post-multiplying builders, i.e. ``mat4(1).rotate(X_AXIS, a).translate(v) = R*T``
(src/data/scenes/boxtest.nim:35-36).

Matrices here are numpy (4, 4) arrays indexed M[row, col]; `to_c` flattens them
to the C ABI layout m[col*4 + row] (include/nrt.h).
"""
from __future__ import annotations

import math

import numpy as np

X_AXIS = np.array([1.0, 0.0, 0.0])  # geom.nim:7-9
Y_AXIS = np.array([0.0, 1.0, 0.0])
Z_AXIS = np.array([0.0, 0.0, 1.0])


def mat4(d: float = 1.0) -> np.ndarray:
    return np.eye(4, dtype=np.float64) * d


def translate(m: np.ndarray, v) -> np.ndarray:
    """GLM translate: m * T(v) (the translation is applied first to a vector)."""
    v = np.asarray(v, dtype=np.float64)
    r = m.copy()
    r[:, 3] = m[:, 0] * v[0] + m[:, 1] * v[1] + m[:, 2] * v[2] + m[:, 3]
    return r


def rotate(m: np.ndarray, axis, angle: float) -> np.ndarray:
    """GLM rotate with the fork's argument order (m, axis, angle): m * R."""
    a = np.asarray(axis, dtype=np.float64)
    a = a * (1.0 / math.sqrt(float(a @ a)))
    c, s = math.cos(angle), math.sin(angle)
    t = (1.0 - c) * a
    rot = np.zeros((4, 4))  # rot[row, col]
    rot[0, 0] = c + t[0] * a[0]
    rot[1, 0] = t[0] * a[1] + s * a[2]
    rot[2, 0] = t[0] * a[2] - s * a[1]
    rot[0, 1] = t[1] * a[0] - s * a[2]
    rot[1, 1] = c + t[1] * a[1]
    rot[2, 1] = t[1] * a[2] + s * a[0]
    rot[0, 2] = t[2] * a[0] + s * a[1]
    rot[1, 2] = t[2] * a[1] - s * a[0]
    rot[2, 2] = c + t[2] * a[2]
    r = m.copy()
    for j in range(3):
        r[:, j] = m[:, 0] * rot[0, j] + m[:, 1] * rot[1, j] + m[:, 2] * rot[2, j]
    return r


def scale(m: np.ndarray, v) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    r = m.copy()
    for j in range(3):
        r[:, j] = m[:, j] * v[j]
    return r


def inverse(m: np.ndarray) -> np.ndarray:
    """Cofactor (adjugate / determinant) inverse, the GLM way (geom.nim:162)."""
    m = np.asarray(m, dtype=np.float64)
    cof = np.zeros((4, 4))
    for r in range(4):
        for c in range(4):
            minor = np.delete(np.delete(m, r, axis=0), c, axis=1)
            d = (
                minor[0, 0] * (minor[1, 1] * minor[2, 2] - minor[1, 2] * minor[2, 1])
                - minor[0, 1] * (minor[1, 0] * minor[2, 2] - minor[1, 2] * minor[2, 0])
                + minor[0, 2] * (minor[1, 0] * minor[2, 1] - minor[1, 1] * minor[2, 0])
            )
            cof[r, c] = d if (r + c) % 2 == 0 else -d
    det = np.float64(m[0, :] @ cof[0, :])
    with np.errstate(all="ignore"):   # (a singular matrix gives inf / nan entries, as the reference's float division does; no exception)
        return cof.T * (np.float64(1.0) / det)


def deg_to_rad(d: float) -> float:
    return d * (math.pi / 180.0)  # Nim math.degToRad


def normalize(v) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    with np.errstate(all="ignore"):   # (the zero vector normalises to nan, as in the reference; no exception)
        return v * (np.float64(1.0) / np.sqrt(np.float64(v @ v)))


def to_c(m: np.ndarray):
    """(4,4) M[row,col] -> 16 doubles m[col*4+row]."""
    return [float(x) for x in np.asarray(m, dtype=np.float64).T.reshape(-1)]

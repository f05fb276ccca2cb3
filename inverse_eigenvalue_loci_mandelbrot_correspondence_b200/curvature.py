"""Host side of the boundary-curvature consumer.

  compute_curvature_localpoly(P, neighbors=7, closed=True, stride=1)   boundary_curvature_localpoly.py:133-184
  load_points(path)                                                    boundary_curvature_localpoly.py:45-63

The per-point quadratic fits run on the GPU (liblm_b200.so: lm_curvature_localpoly); `stride > 1` keeps the
reference's meaning (evaluate every stride-th point, fill the rest by linear interpolation along the index).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import Stats

last_stats: dict = {}


def load_points(path: str) -> np.ndarray:
    """Two-column CSV with or without an `x,y` header -> float [N, 2] (the format <prefix>_boundary.csv has)."""
    try:
        pts = np.loadtxt(path, delimiter=",", dtype=float)
    except ValueError:
        pts = np.loadtxt(path, delimiter=",", dtype=float, skiprows=1)
    pts = np.atleast_2d(pts)
    if pts.shape[1] != 2:
        raise ValueError("Could not load 2D points from CSV (expect two columns 'x,y').")
    return pts


def compute_curvature_localpoly(P, neighbors: int = 7, closed: bool = True, stride: int = 1):
    """(kappa, kappa_signed, speed, aux) with aux = dict(xprime, yprime, x2, y2), as the reference returns them."""
    P = np.asarray(P, dtype=np.float64)
    if P.ndim != 2 or P.shape[1] != 2:
        raise ValueError("P must be an (N, 2) array of ordered boundary points")
    m = int(neighbors)
    if m < 2:
        raise ValueError("neighbors must be >= 2 for a meaningful quadratic fit.")
    N = P.shape[0]
    x = np.ascontiguousarray(P[:, 0]); y = np.ascontiguousarray(P[:, 1])
    out = [np.empty(N, dtype=np.float64) for _ in range(7)]
    st = Stats()
    _shim.call("lm_curvature_localpoly", _shim.ptr(x), _shim.ptr(y), N, m, int(bool(closed)),
               *[_shim.ptr(o) for o in out], C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    stride = max(1, int(stride))
    if stride > 1 and N:
        known = np.arange(0, N, stride)
        missing = np.setdiff1d(np.arange(N), known)
        for arr in out:
            arr[missing] = np.interp(missing, known, arr[known])
    kappa, kappa_s, speed, x1, y1, x2, y2 = out
    return kappa, kappa_s, speed, dict(xprime=x1, yprime=y1, x2=x2, y2=y2)

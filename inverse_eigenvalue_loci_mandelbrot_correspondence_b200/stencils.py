"""Host side of K4: 5-point stencils behind the reference's function surface.

  laplacian(U, h)      Laplacian_C-M.py:49-59
  laplacian_fd(U, h)   Iterative_Variogram_Laplacian.py:132-136
  smooth5(g)           the 5-point interior average of variograms_construct_mandelbrot.py:169-173
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import Stats

last_stats: dict = {}


def _run(name: str, U, *extra):
    U = np.ascontiguousarray(U, dtype=np.float64)
    if U.ndim != 2:
        raise ValueError("expected a 2-D field")
    out = np.empty_like(U)
    st = Stats()
    _shim.call(name, _shim.ptr(U), U.shape[0], U.shape[1], *extra, _shim.ptr(out), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return out


def laplacian(U, h: float):
    """(-4U + roll(U,1,0) + roll(U,-1,0) + roll(U,1,1) + roll(U,-1,1)) / h**2, bit-exact."""
    return _run("lm_laplacian5_periodic", U, float(h))


laplacian_fd = laplacian


def smooth5(g):
    """G = g.copy(); G[1:-1,1:-1] = (g + up + down + left + right)/5.0, bit-exact."""
    return _run("lm_smooth5_interior", g)

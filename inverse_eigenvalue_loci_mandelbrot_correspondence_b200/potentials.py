"""Host side of K4a / K1b: log-potentials of the Lucas Loci cloud and distance-estimator grids.

  log_potential(points, grid_x, grid_y)        Potentials.py:19-27
  construct_potential(Zx, Zy, C)               Laplacian_C-M.py:16-25
  log_potential (hypot form)                   Iterative_Variogram_Laplacian.py:102-112
  log_potential_from_points(grid, pts, eps)    variograms_construct_mandelbrot.py:128-146
  mandelbrot_distance_estimator                construct_stage1_clean.py:50-58,
                                               variograms_construct_mandelbrot.py:61-88
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import (DE_FINAL_DZ, DE_FINAL_DZ_NUMPY, DE_FIRST_ESCAPE, DE_SCALAR, LOGPOT_LOG_INV, LOGPOT_NEG_PERTERM, LOGPOT_SUM_HYPOT,
                    LOGPOT_SUM_SQRT, Stats)

last_stats: dict = {}


def _logpot(px, py, gx, gy, eps: float, variant: int) -> np.ndarray:
    px = np.ascontiguousarray(px, dtype=np.float64).ravel()
    py = np.ascontiguousarray(py, dtype=np.float64).ravel()
    gx = np.ascontiguousarray(gx, dtype=np.float64).ravel()
    gy = np.ascontiguousarray(gy, dtype=np.float64).ravel()
    U = np.empty((gy.size, gx.size), dtype=np.float64)
    st = Stats()
    _shim.call("lm_log_potential", _shim.ptr(px), _shim.ptr(py), px.size, _shim.ptr(gx), gx.size,
               _shim.ptr(gy), gy.size, float(eps), int(variant), _shim.ptr(U), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return U


def log_potential(points, grid_x, grid_y, eps: float = 1e-12, use_hypot: bool = False):
    """U = (1/N) sum log(|z - p| + eps) on the grid (Potentials.py:19-27; use_hypot=True for the
    np.hypot form of Iterative_Variogram_Laplacian.py:102-112)."""
    pts = np.asarray(points, dtype=np.float64)
    pts = pts.reshape(-1, pts.shape[-1])[:, :2]
    return _logpot(pts[:, 0], pts[:, 1], grid_x, grid_y, eps, LOGPOT_SUM_HYPOT if use_hypot else LOGPOT_SUM_SQRT)


def construct_potential(Zx, Zy, Cpts, eps: float = 1e-12):
    """U_C = -(1/N) sum log(|z - c_i| + eps) per term (Laplacian_C-M.py:16-25); Zx, Zy from np.meshgrid(x, y)."""
    Zx = np.asarray(Zx, dtype=float); Zy = np.asarray(Zy, dtype=float)
    xs = Zx[0, :]; ys = Zy[:, 0]
    if not (np.array_equal(Zx, np.broadcast_to(xs[None, :], Zx.shape)) and
            np.array_equal(Zy, np.broadcast_to(ys[:, None], Zy.shape))):
        raise ValueError("construct_potential expects Zx, Zy from np.meshgrid(x, y)")
    pts = np.asarray(Cpts, dtype=np.float64).reshape(-1, np.asarray(Cpts).shape[-1])[:, :2]
    return _logpot(pts[:, 0], pts[:, 1], xs, ys, eps, LOGPOT_NEG_PERTERM)


def log_potential_from_points(xs, ys, pts, eps: float = 1e-6):
    """U_C = (1/N) sum log(1/(|z - p_k| + eps)) (variograms_construct_mandelbrot.py:128-146);
    xs, ys are the grid axes (grid.x, grid.y there), pts complex."""
    pts = np.asarray(pts, dtype=np.complex128).ravel()
    if pts.size == 0:
        return np.zeros((np.size(ys), np.size(xs)), dtype=float)
    return _logpot(pts.real, pts.imag, xs, ys, eps, LOGPOT_LOG_INV)


def distance_grid(xs, ys, max_iter: int = 200, bailout: float = 1e6, eps: float = 1e-16, variant: int = DE_SCALAR):
    """Distance-estimator field and first-escape mask on the grid.

    variant=DE_SCALAR: construct_stage1_clean.py:50-58 (bailout 1e6);
    variant=DE_FIRST_ESCAPE: variograms_construct_mandelbrot.py:61-88 (R=4, eps=1e-14);
    variant=DE_FINAL_DZ: tci_construct_mandelbrot_v002_fixed.py:35-47 (R=250, eps=1e-12; dz from the end of the loop);
    variant=DE_FINAL_DZ_NUMPY: the same with numpy's SIMD complex multiply (FMA recipe) restated -- bit-exact masks.
    """
    xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
    ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
    d = np.empty((ys.size, xs.size), dtype=np.float64)
    e = np.empty((ys.size, xs.size), dtype=np.uint8)
    st = Stats()
    _shim.call("lm_distance_grid_f64", _shim.ptr(xs), xs.size, _shim.ptr(ys), ys.size, int(max_iter), float(bailout),
               float(eps), int(variant), _shim.ptr(d), _shim.ptr(e), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return d, e.astype(bool)


def mandelbrot_distance_estimator(c: complex, max_iter: int = 200, bailout: float = 1e6) -> float:
    """Scalar drop-in for construct_stage1_clean.py:50-58."""
    d, _ = distance_grid([complex(c).real], [complex(c).imag], max_iter, bailout, 1e-16, DE_SCALAR)
    return float(d[0, 0])


def nearest_match(X, Y):
    """For every X (complex) the FIRST index of the nearest Y and that distance: the selection
    argmax(exp(-cdist(X, Y)/const), axis=1) makes in entropic_ot_alignment
    (tci_construct_mandelbrot_v002_fixed.py:62-71).  Returns (index int64[n], distance float64[n])."""
    X = np.asarray(X, dtype=np.complex128).ravel(); Y = np.asarray(Y, dtype=np.complex128).ravel()
    xr = np.ascontiguousarray(X.real); xi = np.ascontiguousarray(X.imag)
    yr = np.ascontiguousarray(Y.real); yi = np.ascontiguousarray(Y.imag)
    idx = np.empty(X.size, dtype=np.int64); dist = np.empty(X.size, dtype=np.float64)
    st = Stats()
    _shim.call("lm_nearest_match", _shim.ptr(xr), _shim.ptr(xi), X.size, _shim.ptr(yr), _shim.ptr(yi), Y.size,
               _shim.ptr(idx), _shim.ptr(dist), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return idx, dist


def sample_mandelbrot_boundary(nx: int = 120, ny: int = 80, max_iter: int = 200, threshold_low: float = 1e-6,
                               threshold_high: float = 1e-1, nsamples: int = 800) -> np.ndarray:
    """sample_mandelbrot_boundary of construct_stage1_clean.py:60-80: the scalar distance estimator on the
    [-2.25, 1.25] x [-1.25, 1.25] grid (one device launch instead of nx*ny Python calls), candidates with
    threshold_low < d < threshold_high in the reference's y-outer / x-inner order, and -- when there are more than
    nsamples -- the same np.random.choice(len(cand), nsamples, replace=False, p=d / sum(d)) draw.  float [k, 2]."""
    xs = np.linspace(-2.25, 1.25, nx)
    ys = np.linspace(-1.25, 1.25, ny)
    d, _ = distance_grid(xs, ys, int(max_iter), 1e6, 1e-16, DE_SCALAR)
    jj, ii = np.nonzero((d > threshold_low) & (d < threshold_high))          # row-major = for y in ys: for x in xs
    cand = np.column_stack([xs[ii], ys[jj]]).astype(float)
    vals = d[jj, ii].astype(float)
    if cand.size == 0:
        return np.empty((0, 2), dtype=float)
    if len(cand) <= nsamples:
        return cand
    probs = vals / np.sum(vals)
    idx = np.random.choice(len(cand), size=nsamples, replace=False, p=probs)
    return cand[idx]


def mandelbrot_boundary_points(xmin: float = -2.25, xmax: float = 1.25, ymin: float = -1.75, ymax: float = 1.75,
                               N: int = 600, dist_thresh: float = 0.002, max_iter: int = 500) -> np.ndarray:
    """mandelbrot_boundary_points of variograms_construct_mandelbrot.py:90-104 (same in ...v2.py:97-111): escaped grid
    points whose first-escape distance estimate (R = 4, eps = 1e-14) is <= dist_thresh, as a complex array in the row-major
    order C[near] has."""
    xs = np.linspace(xmin, xmax, N)
    ys = np.linspace(ymin, ymax, N)
    d, esc = distance_grid(xs, ys, int(max_iter), 4.0, 1e-14, DE_FIRST_ESCAPE)
    jj, ii = np.nonzero(esc & (d <= dist_thresh))
    return xs[ii] + 1j * ys[jj]


"""Host side of K1: the escape-time grid / point kernels behind the reference's function surface.

Mirrors (same names, argument meaning, return layout, error behaviour):
  compute_grid, mandelbrot_dwell        mandelbrot_boundary_sample.py:22-39
  batch_potential,
  mandelbrot_parameter_potential        lucas_equipotential_test_v3.py:124-162
  escape_potential                      Potentials.py:32-47 / Iterative_Variogram_Laplacian.py:114-130
  mandelbrot_potential                  Laplacian_C-M.py:27-43
  mandelbrot_escape_potential           variograms_construct_mandelbrot.py:148-173

All arithmetic happens in liblm_b200.so on the GPU (ctypes over numpy buffers).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import FIELD_GREEN, FIELD_INV_K, FIELD_NONE, FIELD_POW2_ALWAYS, FIELD_POW2_FIRST, Stats

last_stats: dict = {}


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def escape_grid(xs, ys, max_iter: int, bailout: float = 2.0, field_mode: int = FIELD_NONE,
                want_dwell: str | None = "i32", pinned: bool = False):
    """Run K1 over the grid c = xs[i] + 1j*ys[j].

    want_dwell: "i32", "f64" (the reference's float Z) or None.
    Returns (dwell or None, field or None, stats dict).  dwell[j, i] is bit-exact with
    mandelbrot_dwell(xs[i], ys[j], max_iter) when bailout == 2 (test |z|^2 > 4.0).
    """
    xs = _f64(xs).ravel(); ys = _f64(ys).ravel()
    nx, ny = xs.size, ys.size
    alloc = _shim.pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dtype=dt))
    d32 = alloc((ny, nx), np.int32) if want_dwell == "i32" else None
    d64 = alloc((ny, nx), np.float64) if want_dwell == "f64" else None
    fld = alloc((ny, nx), np.float64) if field_mode != FIELD_NONE else None
    st = Stats()
    _shim.call("lm_escape_grid_f64", _shim.ptr(xs), nx, _shim.ptr(ys), ny, int(max_iter), float(bailout),
               int(field_mode), _shim.ptr(d32), _shim.ptr(d64), _shim.ptr(fld), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return (d32 if d32 is not None else d64), fld, last_stats


def escape_grid_f32(xs, ys, max_iter: int, bailout: float = 2.0):
    """The optional single-precision variant of K1 (no reference counterpart; BASELINE.json north_star piece 1): the same
    persistent kernel instantiated in binary32.  Returns (dwell int32 [ny, nx], stats).  NOT bit-exact: measured against
    the fp64 kernel, 0.2-0.3 % of the pixels of configs 1-3 differ (8-9 % in the deep zoom of config 4) and the
    interior mask agrees on >= 99.97 % of the pixels (tests/test_gpu_escape.py::test_f32_variant_tolerance)."""
    xs = _f64(xs).ravel(); ys = _f64(ys).ravel()
    d32 = np.empty((ys.size, xs.size), dtype=np.int32)
    st = Stats()
    _shim.call("lm_escape_grid_f32", _shim.ptr(xs), xs.size, _shim.ptr(ys), ys.size, int(max_iter), float(bailout),
               _shim.ptr(d32), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return d32, last_stats


def compute_grid(xlim, ylim, res: int, max_iter: int):
    """Drop-in for compute_grid (mandelbrot_boundary_sample.py:32-39): -> (xs, ys, Z float64[res,res])."""
    xs = np.linspace(xlim[0], xlim[1], res)
    ys = np.linspace(ylim[0], ylim[1], res)
    Z, _, _ = escape_grid(xs, ys, max_iter, 2.0, FIELD_NONE, want_dwell="f64")
    return xs, ys, Z


def mandelbrot_dwell(x: float, y: float, max_iter: int = 300) -> int:
    """Drop-in for the scalar mandelbrot_dwell (one-pixel grid; use compute_grid for fields)."""
    d, _, _ = escape_grid([x], [y], max_iter, 2.0, FIELD_NONE, want_dwell="i32")
    return int(d[0, 0])


def batch_potential(Cpts, max_iter: int = 4000, escape_radius: float = 2.0):
    """Drop-in for batch_potential (lucas_equipotential_test_v3.py:153-162) -> (g, it, phi)."""
    Cpts = np.asarray(Cpts, dtype=np.complex128).ravel()
    n = Cpts.size
    cre = _f64(Cpts.real); cim = _f64(Cpts.imag)
    g = np.empty(n, dtype=float)
    it = np.empty(n, dtype=np.int64)
    pr = np.empty(n); pi = np.empty(n)
    st = Stats()
    _shim.call("lm_escape_points_f64", _shim.ptr(cre), _shim.ptr(cim), n, int(max_iter), float(escape_radius),
               _shim.ptr(g), _shim.ptr(it), _shim.ptr(pr), _shim.ptr(pi), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    phi = np.empty(n, dtype=np.complex128)
    phi.real = pr; phi.imag = pi
    return g, it.astype(int), phi


def mandelbrot_parameter_potential(c: complex, max_iter: int = 4000, escape_radius: float = 2.0):
    """Drop-in for the scalar function (lucas_equipotential_test_v3.py:124-151) -> (g, k, phi)."""
    g, it, phi = batch_potential(np.array([c], dtype=np.complex128), max_iter, escape_radius)
    return float(g[0]), int(it[0]), complex(phi[0])


def escape_potential(grid_x, grid_y, max_iter: int = 200, R: float = 10, variant: str = "potentials"):
    """escape_potential of Potentials.py:32-47 (variant="potentials": log|z|/2**k, also evaluated
    when the orbit never escapes; raises OverflowError like the reference when k > 1023) or of
    Iterative_Variogram_Laplacian.py:114-130 (variant="iterative": log|z|/(k+1), 0 when bounded)."""
    mode = FIELD_POW2_ALWAYS if variant == "potentials" else FIELD_INV_K
    _, U, _ = escape_grid(grid_x, grid_y, max_iter, float(R), mode, want_dwell=None)
    return U


def mandelbrot_potential(Zx, Zy, max_iter: int = 200, R: float = 2.0):
    """mandelbrot_potential of Laplacian_C-M.py:27-43; Zx, Zy are np.meshgrid(x, y) arrays."""
    Zx = np.asarray(Zx, dtype=float); Zy = np.asarray(Zy, dtype=float)
    xs = Zx[0, :]; ys = Zy[:, 0]
    if not (np.array_equal(Zx, np.broadcast_to(xs[None, :], Zx.shape)) and
            np.array_equal(Zy, np.broadcast_to(ys[:, None], Zy.shape))):
        raise ValueError("mandelbrot_potential expects Zx, Zy from np.meshgrid(x, y)")
    _, U, _ = escape_grid(xs, ys, max_iter, float(R), FIELD_INV_K, want_dwell=None)
    return U


def green_potential_grid(xs, ys, max_iter: int, escape_radius: float = 2.0, want_dwell: str | None = "i32"):
    """g = log|z_k| 2^-k of lucas_equipotential_test_v3.py:124-151 evaluated on a grid (BASELINE config 2)."""
    return escape_grid(xs, ys, max_iter, float(escape_radius), FIELD_GREEN, want_dwell=want_dwell)


def mandelbrot_escape_potential_raw(xs, ys, max_iter: int = 500, R: float = 4.0):
    """variograms_construct_mandelbrot.py:148-167 before its smoothing step (unfused recurrence)."""
    _, g, _ = escape_grid(xs, ys, max_iter, float(R), FIELD_POW2_FIRST, want_dwell=None)
    return g

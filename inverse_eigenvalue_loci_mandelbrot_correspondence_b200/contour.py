"""Host side of K2: level set of the dwell field behind the reference's extract_contour.

  extract_contour(xs, ys, Z, max_iter, level_frac)   mandelbrot_boundary_sample.py:41-54,
                                                     mandelbrot_boundary_sample_spyder.py:35-43
The reference calls plt.contour(xs, ys, Z, levels=[level_frac*max_iter]) and keeps the line
with most vertices (cs.allsegs[0] semantics: one (N,2) array per connected line, closed loops
repeat their first vertex).  Here the quads are classified and the vertices computed on the
GPU (liblm_b200.so), and so is the ordered chaining of the lines (contourpy's mpl2014 rules).
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np

from . import _shim
from ._shim import Stats

last_stats: dict = {}


def _as_dwell_i32(Z) -> np.ndarray:
    Z = np.asarray(Z)
    if Z.dtype == np.int32:
        return np.ascontiguousarray(Z)
    Zi = np.ascontiguousarray(Z, dtype=np.int32)
    if not np.array_equal(Zi, Z):
        raise ValueError("the contour kernel takes an integer-valued dwell field (as compute_grid returns)")
    return Zi


class Polylines(Sequence):
    """The lines of one level as a read-only sequence of (N,2) arrays (what cs.allsegs[0] is in the reference),
    stored as one vertex array plus line offsets; the per-line arrays are views made on access, so a level with
    10^4 lines costs nothing until somebody iterates over it."""

    def __init__(self, verts: np.ndarray, offsets: np.ndarray):
        self.verts = verts              # [n_vertices, 2]
        self.offsets = offsets          # [n_lines + 1]

    def __len__(self) -> int:
        return int(self.offsets.size) - 1

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        n = len(self)
        if k < 0:
            k += n
        if not 0 <= k < n:
            raise IndexError("line index out of range")
        return self.verts[self.offsets[k]:self.offsets[k + 1]]

    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def longest(self):
        """The first of the lines with most vertices (max(paths, key=len) in the reference), or None."""
        if len(self) == 0:
            return None
        return self[int(np.argmax(self.lengths()))]


def _call_then_fetch(fn_name: str, head_args: tuple, level: float, tail_args: tuple = ()) -> Polylines:
    """Run a contour entry point with zero output capacity -- the library links the lines, keeps them and reports
    their sizes (LM_E_CAP) -- then fetch them into arrays of exactly that size: one copy, no over-allocation."""
    st = Stats()
    nv = C.c_int64(0); nl = C.c_int64(0)
    first = np.zeros(1, dtype=np.int64)
    lib = _shim.load()
    with _shim._lock:
        rc = getattr(lib, fn_name)(*head_args, float(level), *tail_args, None, 0, C.byref(nv),
                                   _shim.ptr(first), 0, C.byref(nl), C.byref(st))
        if rc == _shim.LM_E_CAP:
            verts = np.empty((nv.value, 2), dtype=np.float64)
            offs = np.empty(nl.value + 1, dtype=np.int64)
            rc = lib.lm_contour_fetch_last(_shim.ptr(verts), nv.value, C.byref(nv), _shim.ptr(offs), nl.value, C.byref(nl))
        else:
            verts = np.empty((0, 2), dtype=np.float64)
            offs = np.zeros(1, dtype=np.int64)
    _shim.check(rc)
    global last_stats
    last_stats = st.as_dict()
    return Polylines(verts, offs)


def contour_lines(xs, ys, Z, level: float):
    """All level-`level` lines of the dwell field Z[j,i] at (xs[i], ys[j]), in matplotlib's order."""
    xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
    ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
    d = _as_dwell_i32(Z)
    if d.shape != (ys.size, xs.size):
        raise ValueError("Z must have shape (len(ys), len(xs))")
    return _call_then_fetch("lm_contour_level", (_shim.ptr(d), _shim.ptr(xs), xs.size, _shim.ptr(ys), ys.size), level)


def contour_lines_dev(dwell_dev_ptr: int, xs, ys, level: float):
    """Same with the int32 dwell grid already resident on the current device."""
    xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
    ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
    return _call_then_fetch("lm_contour_level_dev",
                            (C.c_void_p(dwell_dev_ptr), _shim.ptr(xs), xs.size, _shim.ptr(ys), ys.size), level)


def boundary_sample(xs, ys, max_iter: int, level: float, dwell_out: np.ndarray | None = None,
                    potential_out: np.ndarray | None = None):
    """compute_grid + plt.contour in one host-buffer call (lm_boundary_sample): the dwell grid never
    leaves the GPU between K1 and K2.  dwell_out: optional int32 or float64 [ny, nx] array that
    receives the dwell grid (copied back while the GPU is still computing); potential_out: optional
    float64 [ny, nx] array for the smooth potential log|z_k| 2^-k of the same pass (config 2).
    Returns (lines, stats) with lines as in contour_lines."""
    xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
    ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
    d32 = d64 = None
    if dwell_out is not None:
        if dwell_out.shape != (ys.size, xs.size) or not dwell_out.flags["C_CONTIGUOUS"]:
            raise ValueError("dwell_out must be a C-contiguous [len(ys), len(xs)] array")
        if dwell_out.dtype == np.int32:
            d32 = dwell_out
        elif dwell_out.dtype == np.float64:
            d64 = dwell_out
        else:
            raise ValueError("dwell_out must be int32 or float64")
    if potential_out is not None:
        if (potential_out.shape != (ys.size, xs.size) or not potential_out.flags["C_CONTIGUOUS"]
                or potential_out.dtype != np.float64):
            raise ValueError("potential_out must be a C-contiguous float64 [len(ys), len(xs)] array")
        lines = _call_then_fetch("lm_boundary_sample_potential",
                                 (_shim.ptr(xs), xs.size, _shim.ptr(ys), ys.size, int(max_iter)), level,
                                 tail_args=(_shim.ptr(d32), _shim.ptr(d64), _shim.ptr(potential_out)))
        return lines, last_stats
    lines = _call_then_fetch("lm_boundary_sample", (_shim.ptr(xs), xs.size, _shim.ptr(ys), ys.size, int(max_iter)), level,
                             tail_args=(_shim.ptr(d32), _shim.ptr(d64)))
    return lines, last_stats


def longest(lines):
    """max(paths, key=#vertices): the first of the longest lines (mandelbrot_boundary_sample.py:53)."""
    if isinstance(lines, Polylines):
        return lines.longest()
    if not lines:
        return None
    return max(lines, key=lambda a: a.shape[0])


def extract_contour(xs, ys, Z, max_iter: int, level_frac: float = 0.96):
    """Drop-in for extract_contour: the longest isocontour of Z at level_frac*max_iter, or None."""
    target = level_frac * max_iter
    return longest(contour_lines(xs, ys, Z, target))


def _link(fn_name: str, head_args: tuple, xs, ys, level: float, tail_args: tuple = ()) -> Polylines:
    xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
    ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
    nv = C.c_int64(0); nl = C.c_int64(0)
    first = np.zeros(1, dtype=np.int64)
    lib = _shim.load()
    with _shim._lock:
        rc = getattr(lib, fn_name)(*head_args, _shim.ptr(xs), xs.size, _shim.ptr(ys), ys.size, float(level),
                                   None, 0, C.byref(nv), _shim.ptr(first), 0, C.byref(nl), *tail_args)
        if rc == _shim.LM_E_CAP:
            verts = np.empty((nv.value, 2), dtype=np.float64)
            offs = np.empty(nl.value + 1, dtype=np.int64)
            rc = lib.lm_contour_fetch_last(_shim.ptr(verts), nv.value, C.byref(nv), _shim.ptr(offs), nl.value, C.byref(nl))
        else:
            verts = np.empty((0, 2), dtype=np.float64)
            offs = np.zeros(1, dtype=np.int64)
    _shim.check(rc)
    return Polylines(verts, offs)


def link_records(records: np.ndarray, xs, ys, level: float) -> Polylines:
    """Chain raster-ordered crossing records (lm_contour_classify_dev; of one block, or of consecutive row blocks
    concatenated) into polylines in matplotlib's order.  The records are uploaded and linked on the GPU
    (lm_contour_link); xs, ys are the full grid coordinates."""
    records = np.ascontiguousarray(records, dtype=np.int64).reshape(-1, 8)
    return _link("lm_contour_link", (_shim.ptr(records), records.shape[0]), xs, ys, level)


def link_records_dev(records_dev_ptr: int, n_records: int, xs, ys, level: float, stream=None) -> Polylines:
    """Same with the records already in device memory (lm_contour_records_dev / the NCCL gather of the shards)."""
    return _link("lm_contour_link_dev", (C.c_void_p(records_dev_ptr), int(n_records)), xs, ys, level, tail_args=(stream,))

"""Host side of the binned pair statistics: same names, arguments and return values as

  empirical_variogram_field(locs, values, nbins, max_dist)        Variogram-Mandelbrot-Construct.py:106-130
  empirical_variogram_coords(locs, nbins, max_dist)               Variogram-Mandelbrot-Construct.py:132-152
  empirical_variogram_from_field_locs(locs, values, max_dist, nbins)   Iterative_Variogram_Laplacian.py:53-86
  pair_correlation(points, r_max, dr), ripley_K(points, r_max, dr)     spatial_stats_phase2.py:9-47

The reference materialises all N(N-1)/2 distances with scipy (5.7 GB for the tracker's 37 820-point cloud) and masks
them once per bin; here the O(N^2) part is one pass of liblm_b200.so:lm_pair_histogram (+ lm_pair_max_distance for the
default max_dist = 0.5 * D.max()).  Counts are bit-identical to the reference's; gamma agrees to rounding (1e-12: the
reference's np.mean is a pairwise sum).  The O(nbins) arithmetic around it is written as the reference writes it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import Stats

MAX_DIST_FACTOR = 0.5          # Variogram-Mandelbrot-Construct.py:24
MAX_BINS_PER_CALL = 2048
last_stats: dict = {}

_WEIGHT = {"none": _shim.PAIR_W_NONE, "value": _shim.PAIR_W_VALUE_SQDIFF, "dist2": _shim.PAIR_W_DIST_SQ}


def _xy(locs):
    P = np.asarray(locs, dtype=np.float64)
    if P.ndim != 2 or P.shape[1] != 2:
        raise ValueError("locs must have shape (N, 2)")
    return np.ascontiguousarray(P[:, 0]), np.ascontiguousarray(P[:, 1])


def pair_histogram(locs, lo, hi, values=None, weight: str = "none"):
    """counts[k] = #{i<j : lo[k] <= d_ij < hi[k]} and the sum of the per-pair weight over those pairs
    (weight: "none" | "value" -> (v_i - v_j)^2 | "dist2" -> d_ij^2).  Returns (counts uint64[nb], sums float64[nb])."""
    x, y = _xy(locs)
    lo = np.ascontiguousarray(lo, dtype=np.float64).ravel()
    hi = np.ascontiguousarray(hi, dtype=np.float64).ravel()
    if lo.size != hi.size:
        raise ValueError("lo and hi must have the same length")
    mode = _WEIGHT[weight]
    v = None
    if mode == _shim.PAIR_W_VALUE_SQDIFF:
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        if v.size != x.size:
            raise ValueError("one value per location expected")
    counts = np.zeros(lo.size, dtype=np.uint64)
    sums = np.zeros(lo.size, dtype=np.float64)
    global last_stats
    last_stats = {"work_units": 0, "items": int(x.size), "kernel_ms": 0.0, "launches": 0}
    for a in range(0, lo.size, MAX_BINS_PER_CALL):          # the library takes up to 2048 bins per pass
        b = min(a + MAX_BINS_PER_CALL, lo.size)
        clo, chi = np.ascontiguousarray(lo[a:b]), np.ascontiguousarray(hi[a:b])
        cc = np.zeros(b - a, dtype=np.uint64); cs = np.zeros(b - a, dtype=np.float64)
        st = Stats()
        _shim.call("lm_pair_histogram", _shim.ptr(x), _shim.ptr(y), _shim.ptr(v), x.size,
                   _shim.ptr(clo), _shim.ptr(chi), b - a, mode, _shim.ptr(cc), _shim.ptr(cs), C.byref(st))
        counts[a:b] = cc; sums[a:b] = cs
        last_stats["work_units"] += int(st.work_units); last_stats["kernel_ms"] += float(st.kernel_ms)
        last_stats["launches"] += int(st.launches)
    return counts, sums


def max_pair_distance(locs) -> float:
    """pdist(locs).max() without the distances."""
    x, y = _xy(locs)
    out = C.c_double(0.0)
    st = Stats()
    _shim.call("lm_pair_max_distance", _shim.ptr(x), _shim.ptr(y), x.size, C.byref(out), C.byref(st))
    return float(out.value)


def _binned_semivariance(locs, nbins, max_dist, values, weight, empty_max):
    if max_dist is None:
        n = np.asarray(locs).shape[0]
        max_dist = MAX_DIST_FACTOR * max_pair_distance(locs) if n >= 2 else empty_max
    bins = np.linspace(0.0, max_dist, nbins + 1)
    centers = 0.5 * (bins[:-1] + bins[1:])
    gamma = np.full(nbins, np.nan)
    counts = np.zeros(nbins, dtype=int)
    c, s = pair_histogram(locs, bins[:-1], bins[1:], values, weight)
    hit = c > 0
    gamma[hit] = 0.5 * (s[hit] / c[hit])
    counts[hit] = c[hit]
    return centers, gamma, counts


def empirical_variogram_field(locs, values, nbins: int = 50, max_dist=None):
    """Semivariogram of `values` sampled at `locs` -> (lag_centers, gamma, counts)."""
    locs = np.asarray(locs, dtype=np.float64)
    if locs.shape[0] < 2:
        return np.array([]), np.array([]), np.array([])
    return _binned_semivariance(locs, nbins, max_dist, np.asarray(values, dtype=np.float64), "value", None)


def empirical_variogram_coords(locs, nbins: int = 50, max_dist=None):
    """Variogram of the coordinates themselves (squared distance as the 'difference')."""
    locs = np.asarray(locs, dtype=np.float64)
    if locs.shape[0] < 2:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")   # D.max() of the reference
    return _binned_semivariance(locs, nbins, max_dist, None, "dist2", None)


def empirical_variogram_from_field_locs(locs, values=None, max_dist=None, nbins: int = 50):
    """Iterative_Variogram_Laplacian.py:53-86: coordinates variogram when values is None, else the field's."""
    locs = np.asarray(locs, dtype=np.float64)
    if values is None:
        return _binned_semivariance(locs, nbins, max_dist, None, "dist2", 1.0)
    return _binned_semivariance(locs, nbins, max_dist, np.asarray(values, dtype=np.float64), "value", 1.0)


def _density(points):
    N = len(points)
    area = (np.max(points[:, 0]) - np.min(points[:, 0])) * (np.max(points[:, 1]) - np.min(points[:, 1]))
    return N, N / area


def pair_correlation(points, r_max, dr):
    """g(r) on the shells [r, r+dr), r in np.arange(0, r_max, dr)."""
    points = np.asarray(points, dtype=np.float64)
    N, rho = _density(points)
    r_vals = np.arange(0, r_max, dr)
    if r_vals.size == 0:
        return r_vals, np.array([])
    counts, _ = pair_histogram(points, r_vals, r_vals + dr)
    g_r = []
    for r, count in zip(r_vals, counts.astype(np.int64)):
        norm = 2 * np.pi * r * dr * N * rho
        g_r.append(count / norm if norm > 0 else 0)
    return r_vals, np.array(g_r)


def ripley_K(points, r_max, dr):
    """K(r) = 2 * #{d < r} / (N * rho); #{d < r_k} is the number of pairs in the bins [r_j, r_{j+1}) with j < k."""
    points = np.asarray(points, dtype=np.float64)
    N, rho = _density(points)
    r_vals = np.arange(0, r_max, dr)
    if r_vals.size == 0:
        return r_vals, np.array([])
    counts, _ = pair_histogram(points, r_vals, np.append(r_vals[1:], np.inf))
    below = np.concatenate([[0], np.cumsum(counts.astype(np.int64))[:-1]])
    return r_vals, np.array([(2 * c) / (N * rho) for c in below])

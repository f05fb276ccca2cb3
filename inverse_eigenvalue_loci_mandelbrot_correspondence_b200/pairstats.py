"""Host side of the binned pair statistics: same names, arguments and return values as

  empirical_variogram_field(locs, values, nbins, max_dist)        Variogram-Mandelbrot-Construct.py:106-130
  empirical_variogram_coords(locs, nbins, max_dist)               Variogram-Mandelbrot-Construct.py:132-152
  empirical_variogram_from_field_locs(locs, values, max_dist, nbins)   Iterative_Variogram_Laplacian.py:53-86
  pair_correlation(points, r_max, dr), ripley_K(points, r_max, dr)     spatial_stats_phase2.py:9-47
  sample_semivariogram(field, grid, r_bins, max_pairs_per_bin),
  sample_cross_semivariogram(field1, field2, grid, r_bins, max_pairs_per_bin)   variograms_construct_mandelbrot.py:178-315

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


# ---- sub-sampled semivariograms on a regular grid (variograms_construct_mandelbrot.py:178-315) ------------------------
# The reference draws M <= 15000 pixel locations with np.random.choice, walks the pairs in 4000 x 4000 blocks and, per bin,
# keeps at most max_pairs_per_bin pairs: a block that would overflow a bin contributes a random subset drawn with
# np.random.choice(n_in_bin, size=room, replace=False) over dV2[np.where(mask)] (row-major).  Everything random happens on
# the host through numpy's global stream, in the reference's order and with the reference's arguments, so a seeded run
# consumes the same draws; the O(M^2) part -- per-block bin counts and sums, and the ordered dV2 list of the (at most one
# per bin) block that crosses the cap -- runs on the GPU.  gamma agrees to rounding (the reference's part.sum() is a
# pairwise sum over the block's list; the device accumulates in tile order): 1e-12 relative, counts exact.
SUBSAMPLE_POINTS = 15000      # M_target, :195 / :269
SUBSAMPLE_CHUNK = 4000        # chunk, :208 / :284


def _block_counts_sums(A, B, lo, hi, same_block: bool):
    """Bin counts / sums of (v_i - w_j)^2 over the pairs (i in A, j in B) of one block.  A, B = (x, y, v) triples.
    same_block: B is A and the reference's mask only removes the diagonal, so every unordered pair counts twice."""
    xa, ya, va = A
    if same_block:
        c, s = pair_histogram(np.column_stack([xa, ya]), lo, hi, va, "value")
        return 2 * c.astype(np.int64), 2.0 * s
    xb, yb, vb = B
    # cross pairs = pairs of the union - pairs inside A - pairs inside B (counts exactly, sums to a few ulps)
    cu, su = pair_histogram(np.column_stack([np.concatenate([xa, xb]), np.concatenate([ya, yb])]), lo, hi,
                            np.concatenate([va, vb]), "value")
    ca, sa = pair_histogram(np.column_stack([xa, ya]), lo, hi, va, "value")
    cb, sb = pair_histogram(np.column_stack([xb, yb]), lo, hi, vb, "value")
    return cu.astype(np.int64) - ca.astype(np.int64) - cb.astype(np.int64), su - sa - sb


def _select_sqdiff(A, B, lo: float, hi: float, skip_diag: bool, expect: int) -> np.ndarray:
    xa, ya, va = (np.ascontiguousarray(t, dtype=np.float64) for t in A)
    xb, yb, vb = (np.ascontiguousarray(t, dtype=np.float64) for t in B)
    out = np.empty(max(int(expect), 1), dtype=np.float64)
    n = C.c_int64(0)
    st = Stats()
    _shim.call("lm_pair_select_sqdiff", _shim.ptr(xa), _shim.ptr(ya), _shim.ptr(va), xa.size, _shim.ptr(xb), _shim.ptr(yb),
               _shim.ptr(vb), xb.size, float(lo), float(hi), 1 if skip_diag else 0, _shim.ptr(out), out.size, C.byref(n), C.byref(st))
    if n.value != expect:
        raise RuntimeError(f"pair selection returned {n.value} pairs, the block histogram counted {expect}")
    return out[: n.value]


def _capped_block_walk(P1, P2, r_bins, max_pairs_per_bin: int, symmetric: bool):
    """The reference's double loop over 4000-point chunks with its per-bin caps.  P1, P2 = (x, y, v) of the two
    sub-samples (P2 is P1 for the plain semivariogram, where only blocks b >= a are visited)."""
    r_bins = np.asarray(r_bins, dtype=np.float64)
    nbins = len(r_bins) - 1
    lo, hi = r_bins[:-1], r_bins[1:]
    sums = np.zeros(nbins, dtype=float)
    counts = np.zeros(nbins, dtype=int)
    M = P1[0].size
    chunk = SUBSAMPLE_CHUNK
    cut = lambda P, a, b: tuple(t[a:b] for t in P)
    for a in range(0, M, chunk):
        a_end = min(a + chunk, M)
        A = cut(P1, a, a_end)
        for b in range(a if symmetric else 0, M, chunk):
            b_end = min(b + chunk, M)
            B = cut(P2, b, b_end)
            same = symmetric and a == b
            if np.all(counts >= max_pairs_per_bin):
                continue                                   # every bin is full: the reference only evaluates masks here
            bc, bs = _block_counts_sums(A, B, lo, hi, same)
            for k in range(nbins):
                if bc[k] == 0:
                    continue
                room = max_pairs_per_bin - counts[k]
                if room <= 0:
                    continue
                if bc[k] > room:
                    vals = _select_sqdiff(A, B, lo[k], hi[k], same, int(bc[k]))
                    sel = np.random.choice(int(bc[k]), size=room, replace=False)     # the reference's draw, same arguments
                    part = vals[sel]
                    sums[k] += part.sum()
                    counts[k] += part.size
                else:
                    sums[k] += bs[k]
                    counts[k] += int(bc[k])
    gamma = np.zeros(nbins, dtype=float)
    nz = counts > 0
    gamma[nz] = 0.5 * (sums[nz] / counts[nz])
    return 0.5 * (r_bins[:-1] + r_bins[1:]), gamma


def sample_semivariogram(field, grid, r_bins, max_pairs_per_bin: int = 20000):
    """Drop-in for sample_semivariogram (variograms_construct_mandelbrot.py:178-250): grid has .X / .Y (np.meshgrid
    arrays of the field's shape).  Uses numpy's global random stream exactly like the reference (seed it the same way)."""
    field = np.asarray(field)
    X = np.asarray(grid.X, dtype=np.float64).ravel(); Y = np.asarray(grid.Y, dtype=np.float64).ravel()
    vals = field.ravel()
    M_target = min(SUBSAMPLE_POINTS, X.size)
    idx = np.random.choice(X.size, size=M_target, replace=False)
    P = (np.ascontiguousarray(X[idx]), np.ascontiguousarray(Y[idx]), np.ascontiguousarray(vals[idx], dtype=np.float64))
    return _capped_block_walk(P, P, r_bins, int(max_pairs_per_bin), symmetric=True)


def sample_cross_semivariogram(field1, field2, grid, r_bins, max_pairs_per_bin: int = 20000):
    """Drop-in for sample_cross_semivariogram (variograms_construct_mandelbrot.py:252-315)."""
    field1 = np.asarray(field1); field2 = np.asarray(field2)
    assert field1.shape == field2.shape
    X = np.asarray(grid.X, dtype=np.float64).ravel(); Y = np.asarray(grid.Y, dtype=np.float64).ravel()
    M_target = min(SUBSAMPLE_POINTS, X.size)
    idx1 = np.random.choice(X.size, size=M_target, replace=False)
    idx2 = np.random.choice(X.size, size=M_target, replace=False)
    P1 = (np.ascontiguousarray(X[idx1]), np.ascontiguousarray(Y[idx1]), np.ascontiguousarray(field1.ravel()[idx1], dtype=np.float64))
    P2 = (np.ascontiguousarray(X[idx2]), np.ascontiguousarray(Y[idx2]), np.ascontiguousarray(field2.ravel()[idx2], dtype=np.float64))
    return _capped_block_walk(P1, P2, r_bins, int(max_pairs_per_bin), symmetric=False)

"""Host side of K3: the Lucas Loci point cloud behind the reference's function surface.

Mirrors
  generate_lucas_companion, generate_companion_from_toprow, family_toprow,
  compute_inverse_eigenvalues, compute_inverse_eigenvalues_family
                                   lucas_equipotential_test_v3.py:58-118
  lucas_companion, construct_points   tci_construct_mandelbrot.py:5-19,
                                      tci_construct_mandelbrot_v002_fixed.py:24-33,
                                      variograms_construct_mandelbrot.py:40-56
  construct_points(maxN)              construct_stage1_clean.py:34-48

The eigenvalues are computed on the GPU as the roots of x^n - a_1 x^(n-1) - ... - a_n (batched
Aberth-Ehrlich, liblm_b200.so:lm_roots_batched).  The reference's order inside one n is
LAPACK's; here every n is returned sorted lexicographically by (real, imag) of the emitted
value, so compare as sets / after sorting.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import CloudStats, LOGPOT_SUM_SQRT, Stats

last_stats: dict = {}

FAMILIES = ("lucas_all_ones", "pell_like_all_twos", "sparser_gap_1_0_1_then_ones", "padovan_like_0_1_then_ones")


def generate_companion_from_toprow(n: int, top) -> np.ndarray:
    """n x n companion matrix: first row `top`, ones on the sub-diagonal (host-side helper)."""
    top = np.asarray(top, dtype=float).reshape(-1)
    if top.shape[0] != n:
        raise AssertionError("top row must have n entries")
    M = np.zeros((n, n), dtype=float)
    M[0, :] = top
    idx = np.arange(1, n)
    M[idx, idx - 1] = 1.0
    return M


def generate_lucas_companion(n: int) -> np.ndarray:
    return generate_companion_from_toprow(n, np.ones(n))


lucas_companion = generate_lucas_companion


def family_toprow(name: str, n: int) -> np.ndarray:
    """First rows of the reference's four families (lucas_equipotential_test_v3.py:76-91)."""
    top = np.ones(n)
    if name == "lucas_all_ones":
        return top
    if name == "pell_like_all_twos":
        return 2.0 * top
    if name == "sparser_gap_1_0_1_then_ones":
        if n >= 2:
            top[1] = 0.0
        return top
    if name == "padovan_like_0_1_then_ones":
        top[0] = 0.0
        return top
    raise ValueError(f"Unknown family '{name}'")


def roots_batched(toprows, deg, invert: bool = False, tol: float = 0.0, sort: bool = True):
    """Roots of x^d - sum_k a_k x^(d-k) for a zero-padded batch of first rows.

    toprows: float64 [npoly, maxdeg]; deg: int [npoly].
    Returns (values complex128 [npoly, maxdeg] NaN-padded, n_kept int32 [npoly], iters int32 [npoly]).
    With invert=True values are 1/lambda for |lambda| > tol (the Lucas Loci points).
    """
    toprows = np.ascontiguousarray(toprows, dtype=np.float64)
    if toprows.ndim != 2:
        raise ValueError("toprows must be [npoly, maxdeg]")
    npoly, maxdeg = toprows.shape
    deg = np.ascontiguousarray(deg, dtype=np.int32).reshape(-1)
    if deg.shape[0] != npoly:
        raise ValueError("deg must have one entry per polynomial")
    ore = np.empty((npoly, maxdeg), dtype=np.float64)
    oim = np.empty((npoly, maxdeg), dtype=np.float64)
    kept = np.empty(npoly, dtype=np.int32)
    iters = np.empty(npoly, dtype=np.int32)
    st = Stats()
    _shim.call("lm_roots_batched", _shim.ptr(toprows), _shim.ptr(deg), npoly, maxdeg, int(bool(invert)), float(tol),
               _shim.ptr(ore), _shim.ptr(oim), _shim.ptr(kept), _shim.ptr(iters), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    vals = np.empty((npoly, maxdeg), dtype=np.complex128)
    vals.real = ore
    vals.imag = oim
    if sort:
        vals = np.sort(vals, axis=1)      # lexicographic (real, imag); NaN padding sorts last
    return vals, kept, iters


def eigvals_toprow(top) -> np.ndarray:
    """Eigenvalues of the companion matrix with first row `top` (np.linalg.eigvals(companion) in the reference)."""
    top = np.asarray(top, dtype=np.float64).reshape(1, -1)
    vals, kept, _ = roots_batched(top, [top.shape[1]])
    return vals[0, : kept[0]]


def _inverse_eigs_for_rows(rows: list[np.ndarray], tol: float) -> np.ndarray:
    if not rows:
        return np.zeros(0, dtype=np.complex128)
    maxdeg = max(r.shape[0] for r in rows)
    tops = np.zeros((len(rows), maxdeg), dtype=np.float64)
    deg = np.empty(len(rows), dtype=np.int32)
    for k, r in enumerate(rows):
        tops[k, : r.shape[0]] = r
        deg[k] = r.shape[0]
    vals, kept, _ = roots_batched(tops, deg, invert=True, tol=tol)
    return np.concatenate([vals[k, : kept[k]] for k in range(len(rows))]).astype(np.complex128)


def compute_inverse_eigenvalues(n_min: int, n_max: int, tol: float = 1e-12) -> np.ndarray:
    """Lucas Loci for n = n_min..n_max (lucas_equipotential_test_v3.py:93-104)."""
    return _inverse_eigs_for_rows([np.ones(n) for n in range(n_min, n_max + 1)], tol)


def compute_inverse_eigenvalues_family(family: str, n_min: int, n_max: int, tol: float = 1e-12) -> np.ndarray:
    """Same for one of FAMILIES (lucas_equipotential_test_v3.py:106-118)."""
    return _inverse_eigs_for_rows([family_toprow(family, n) for n in range(n_min, n_max + 1)], tol)


def summarize_g(g: np.ndarray, label: str = "", quiet: bool = False) -> dict:
    """summarize_g of lucas_equipotential_test_v3.py:168-184: counts and order statistics of the exterior potentials
    (same keys, same numpy calls, same printed lines unless quiet)."""
    g = np.asarray(g, dtype=np.float64)
    outside = g > 0
    has = bool(outside.any())
    out = {
        "count": int(len(g)),
        "escaped": int(outside.sum()),
        "escaped_frac": float(outside.mean()) if len(g) else float("nan"),
        "g_median": float(np.median(g[outside])) if has else float("nan"),
        "g_mean": float(np.mean(g[outside])) if has else float("nan"),
        "g_std": float(np.std(g[outside])) if has else float("nan"),
        "g_p10": float(np.quantile(g[outside], 0.10)) if has else float("nan"),
        "g_p90": float(np.quantile(g[outside], 0.90)) if has else float("nan"),
    }
    if not quiet:
        print(f"{label}escaped: {out['escaped']}/{out['count']}  ({out['escaped_frac']*100:.2f}%)")
        if has:
            print(f"{label}g median={out['g_median']:.6g}  mean={out['g_mean']:.6g}  std={out['g_std']:.6g}")
            print(f"{label}g p10/p90={out['g_p10']:.6g}/{out['g_p90']:.6g}")
    return out


def _cloud_potentials_by_n(n_min: int, n_max: int, family: str | None, max_iter: int, escape_radius: float, tol: float):
    """ONE K3 batch over n = n_min..n_max and ONE K1d pass over the whole cloud -> (g of every point, points per n).
    The reference solves and re-evaluates per n (and, in cumulative_stats, re-evaluates the whole growing cloud for
    every N: O(N^2) potential evaluations); g of a point does not depend on the other points, so one pass serves all."""
    from . import escape
    rows = [np.ones(n) if family is None else family_toprow(family, n) for n in range(n_min, n_max + 1)]
    if not rows:
        return np.zeros(0), np.zeros(0, dtype=np.int64)
    maxdeg = max(r.shape[0] for r in rows)
    tops = np.zeros((len(rows), maxdeg), dtype=np.float64)
    deg = np.empty(len(rows), dtype=np.int32)
    for k, r in enumerate(rows):
        tops[k, : r.shape[0]] = r
        deg[k] = r.shape[0]
    vals, kept, _ = roots_batched(tops, deg, invert=True, tol=tol)
    cloud = np.concatenate([vals[k, : kept[k]] for k in range(len(rows))]).astype(np.complex128)
    g, _, _ = escape.batch_potential(cloud, max_iter=max_iter, escape_radius=escape_radius)
    return g, kept.astype(np.int64)


def per_n_stats(n_min: int, n_max: int, family: str | None = None, max_iter: int = 20000, escape_radius: float = 2.0,
                tol: float = 1e-12, quiet: bool = False) -> list[dict]:
    """per_n_stats of lucas_equipotential_test_v3.py:294-308 (MAX_ITER / ESCAPE_RADIUS / EIG_TOL of :41-44 as defaults):
    one row {"n", count, escaped, escaped_frac, g_median, g_mean, g_std, g_p10, g_p90} per order n."""
    g, counts = _cloud_potentials_by_n(n_min, n_max, family, max_iter, escape_radius, tol)
    offs = np.concatenate([[0], np.cumsum(counts)])
    return [{"n": n, **summarize_g(g[offs[k]:offs[k + 1]], label=f"[per-n n={n}] ", quiet=quiet)}
            for k, n in enumerate(range(n_min, n_max + 1))]


def cumulative_stats(n_min: int, n_max: int, family: str | None = None, max_iter: int = 20000, escape_radius: float = 2.0,
                     tol: float = 1e-12, quiet: bool = False) -> list[dict]:
    """cumulative_stats of lucas_equipotential_test_v3.py:310-327: row N summarises the cloud of orders n_min..N."""
    g, counts = _cloud_potentials_by_n(n_min, n_max, family, max_iter, escape_radius, tol)
    offs = np.concatenate([[0], np.cumsum(counts)])
    return [{"N": N, **summarize_g(g[: offs[k + 1]], label=f"[cum N={N}] ", quiet=quiet)}
            for k, N in enumerate(range(n_min, n_max + 1))]


def construct_points(ns, tol: float = 1e-10) -> np.ndarray:
    """construct_points(ns) of tci_construct_mandelbrot.py:11-19 (tol 1e-10 there and in
    tci_construct_mandelbrot_v002_fixed.py:27-33; variograms_construct_mandelbrot.py:48-56 uses 1e-14)."""
    return _inverse_eigs_for_rows([np.ones(int(n)) for n in ns], tol)


def construct_points_xy(maxN: int = 40) -> np.ndarray:
    """construct_points(maxN) of construct_stage1_clean.py:34-48: float [N,2] for n = 2..maxN, tol 1e-12."""
    pts = compute_inverse_eigenvalues(2, maxN, 1e-12)
    return np.column_stack([pts.real, pts.imag]).astype(float)


def cloud_fields(toprows, deg, grid_x=None, grid_y=None, tol: float = 1e-12, eps: float = 1e-12,
                 variant: int = LOGPOT_SUM_SQRT, h: float | None = None, potential: tuple[int, float] | None = None,
                 return_cloud: bool = True, laplacian: bool = True, cloud_out: tuple | None = None) -> dict:
    """The Lucas-Loci field stage in one call (BASELINE.json config 5), everything resident in HBM between
    the stages: roots of every polynomial (K3) -> cloud of 1/lambda in polynomial order
    (lucas_equipotential_test_v3.py:93-118) -> batch_potential at the cloud (K1d, :153-162, when
    `potential=(max_iter, R)`) -> log-potential on the grid (K4a, Potentials.py:19-27 by default) -> periodic
    5-point Laplacian (K4, Laplacian_C-M.py:49-59; h defaults to the x spacing as there).

    cloud_out=(re, im): optional preallocated float64 arrays (e.g. page-locked, `_shim.pinned_empty`) of at least
    sum(deg) entries that receive the cloud; "cloud" is then the pair of views (re[:n], im[:n]) instead of a
    complex array (no extra host pass).  Page-locked `toprows` / `deg` make the upload run at PCIe speed too.
    `toprows` of dtype int8 (generalized-Lucas rows are small integers) travel as 1 byte per coefficient and are
    widened on the device (lm_lucas_cloud_fields_i8): same results, 8x less upload.

    Returns {"cloud", "g", "it", "U", "lapU", "stats"}; entries not asked for are None.
    """
    compact = np.asarray(toprows).dtype == np.int8      # small-integer first rows: 1 byte per coefficient over PCIe
    toprows = np.ascontiguousarray(toprows, dtype=np.int8 if compact else np.float64)
    if toprows.ndim != 2:
        raise ValueError("toprows must be [npoly, maxdeg]")
    npoly, maxdeg = toprows.shape
    deg = np.ascontiguousarray(deg, dtype=np.int32).reshape(-1)
    if deg.shape[0] != npoly:
        raise ValueError("deg must have one entry per polynomial")
    want_pts = return_cloud or potential is not None
    if cloud_out is not None:
        cre, cim = cloud_out
        if (cre.dtype != np.float64 or cim.dtype != np.float64
                or not cre.flags["C_CONTIGUOUS"] or not cim.flags["C_CONTIGUOUS"]):
            raise ValueError("cloud_out must be two C-contiguous float64 arrays with at least sum(deg) entries")
        cap = int(min(cre.size, cim.size))       # the library reports LM_E_CAP (LmError) if the cloud is larger
        if potential is not None:
            cap = min(cap, max(int(deg.sum(dtype=np.int64)), 0))
        return_cloud = True
    else:
        cap = max(int(deg.sum(dtype=np.int64)), 0)   # a negative degree makes the call itself fail (LM_E_INVALID)
        cre = np.empty(cap, dtype=np.float64) if return_cloud else None
        cim = np.empty(cap, dtype=np.float64) if return_cloud else None
    g = np.empty(cap, dtype=np.float64) if potential is not None else None
    it = np.empty(cap, dtype=np.int64) if potential is not None else None
    U = lap = gx = gy = None
    nx = ny = 0
    if grid_x is not None and grid_y is not None:
        gx = np.ascontiguousarray(grid_x, dtype=np.float64).ravel()
        gy = np.ascontiguousarray(grid_y, dtype=np.float64).ravel()
        nx, ny = gx.size, gy.size
        U = np.empty((ny, nx), dtype=np.float64)
        if laplacian:
            lap = np.empty((ny, nx), dtype=np.float64)
            if h is None:
                h = float(gx[1] - gx[0])
    st = CloudStats()
    n = C.c_int64(0)
    pmi, pR = (int(potential[0]), float(potential[1])) if potential is not None else (1, 2.0)
    _shim.call("lm_lucas_cloud_fields_i8" if compact else "lm_lucas_cloud_fields", _shim.ptr(toprows), _shim.ptr(deg), npoly, maxdeg, float(tol),
               _shim.ptr(cre), _shim.ptr(cim), cap if want_pts else 0, C.byref(n), pmi, pR, _shim.ptr(g), _shim.ptr(it),
               _shim.ptr(gx), nx, _shim.ptr(gy), ny, float(eps), int(variant), float(h if h is not None else 1.0),
               _shim.ptr(U), _shim.ptr(lap), C.byref(st))
    k = int(n.value)
    cloud = None
    if cloud_out is not None:
        cloud = (cre[:k], cim[:k])
    elif return_cloud:
        cloud = np.empty(k, dtype=np.complex128)
        cloud.real = cre[:k]; cloud.imag = cim[:k]
    return {"cloud": cloud, "g": None if g is None else g[:k], "it": None if it is None else it[:k],
            "U": U, "lapU": lap, "n_points": k, "stats": st.as_dict()}

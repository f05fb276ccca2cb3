"""ctypes front-end of the CPU oracle (oracle/liboracle.so) plus the numpy root oracle.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

The C side (lm_oracle.c, lm_oracle_contour.c) restates the reference's functions; each
wrapper below names the reference function it stands for.  The root oracle IS the
reference's own call, numpy.linalg.eigvals on the companion matrix (LAPACK dgeev via the
numpy in this image; the reference pins no version: requirements.txt:1).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    srcs = [_HERE / "lm_oracle.c", _HERE / "lm_oracle_contour.c", _HERE / "Makefile"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_dwell_grid.restype = C.c_uint64
        L.oracle_dwell_grid.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, C.c_int32, C.c_double, _i32p]
        L.oracle_potential_grid.restype = C.c_int
        L.oracle_potential_grid.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, C.c_int32, C.c_double,
                                            C.c_int32, _i32p, _f64p]
        L.oracle_potential_points.restype = C.c_uint64
        L.oracle_potential_points.argtypes = [_f64p, _f64p, C.c_int64, C.c_int32, C.c_double,
                                              _f64p, _i64p, _f64p, _f64p]
        L.oracle_distance_grid.restype = None
        L.oracle_distance_grid.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, C.c_int32, C.c_double,
                                           C.c_double, C.c_int32, _f64p, _u8p]
        L.oracle_laplacian5_periodic.restype = None
        L.oracle_laplacian5_periodic.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_double, _f64p]
        L.oracle_smooth5_interior.restype = None
        L.oracle_smooth5_interior.argtypes = [_f64p, C.c_int64, C.c_int64, _f64p]
        L.oracle_log_potential.restype = None
        L.oracle_log_potential.argtypes = [_f64p, _f64p, C.c_int64, _f64p, C.c_int64, _f64p, C.c_int64,
                                           C.c_double, C.c_int32, _f64p]
        L.oracle_nearest_match.restype = None
        L.oracle_nearest_match.argtypes = [_f64p, _f64p, C.c_int64, _f64p, _f64p, C.c_int64, _i64p, _f64p]
        L.oracle_pair_histogram.restype = C.c_double
        L.oracle_pair_histogram.argtypes = [_f64p, _f64p, C.c_void_p, C.c_int64, _f64p, _f64p, C.c_int32, C.c_int32,
                                            np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS"), _f64p]
        L.oracle_contour_lines.restype = C.c_int
        L.oracle_contour_lines.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, _f64p, C.c_double,
                                           _f64p, C.c_int64, C.POINTER(C.c_int64),
                                           _i64p, C.c_int64, C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---- K1 -------------------------------------------------------------------------------
def dwell_grid(xs, ys, max_iter: int, bail2: float = 4.0):
    """compute_grid, mandelbrot_boundary_sample.py:32-39 -> (dwell int32[ny,nx], pixel_iters)."""
    xs = _c(xs, np.float64); ys = _c(ys, np.float64)
    out = np.empty((ys.size, xs.size), dtype=np.int32)
    work = lib().oracle_dwell_grid(xs, xs.size, ys, ys.size, int(max_iter), float(bail2), out)
    return out, int(work)


def compute_grid(xlim, ylim, res: int, max_iter: int):
    """Same signature/return as the reference compute_grid (Z float64)."""
    xs = np.linspace(xlim[0], xlim[1], res)
    ys = np.linspace(ylim[0], ylim[1], res)
    d, _ = dwell_grid(xs, ys, max_iter)
    return xs, ys, d.astype(np.float64)


FIELD_GREEN, FIELD_POW2_ALWAYS, FIELD_INV_K, FIELD_POW2_FIRST = 1, 2, 3, 4


def potential_grid(xs, ys, max_iter: int, R: float, mode: int):
    """Grid potentials (see lm_oracle.c).  Returns (dwell, field); raises OverflowError like the reference."""
    xs = _c(xs, np.float64); ys = _c(ys, np.float64)
    d = np.empty((ys.size, xs.size), dtype=np.int32)
    f = np.empty((ys.size, xs.size), dtype=np.float64)
    rc = lib().oracle_potential_grid(xs, xs.size, ys, ys.size, int(max_iter), float(R), int(mode), d, f)
    if rc:
        raise OverflowError("int too large to convert to float")
    return d, f


def batch_potential(Cpts, max_iter: int = 4000, escape_radius: float = 2.0):
    """batch_potential, lucas_equipotential_test_v3.py:153-162 -> (g, it, phi)."""
    Cpts = np.asarray(Cpts, dtype=np.complex128).ravel()
    cre = _c(Cpts.real, np.float64); cim = _c(Cpts.imag, np.float64)
    n = Cpts.size
    g = np.empty(n); it = np.empty(n, dtype=np.int64); pr = np.empty(n); pi = np.empty(n)
    lib().oracle_potential_points(cre, cim, n, int(max_iter), float(escape_radius), g, it, pr, pi)
    return g, it, pr + 1j * pi


def distance_grid(xs, ys, max_iter: int, bailout: float, eps: float, variant: int):
    xs = _c(xs, np.float64); ys = _c(ys, np.float64)
    d = np.empty((ys.size, xs.size), dtype=np.float64)
    e = np.empty((ys.size, xs.size), dtype=np.uint8)
    lib().oracle_distance_grid(xs, xs.size, ys, ys.size, int(max_iter), float(bailout), float(eps), int(variant), d, e)
    return d, e.astype(bool)


# ---- K4 / K4a -------------------------------------------------------------------------
def laplacian(U, h: float):
    """laplacian, Laplacian_C-M.py:49-59."""
    U = _c(U, np.float64)
    out = np.empty_like(U)
    lib().oracle_laplacian5_periodic(U, U.shape[0], U.shape[1], float(h), out)
    return out


def smooth5(g):
    """5-point interior average, variograms_construct_mandelbrot.py:169-173."""
    g = _c(g, np.float64)
    out = np.empty_like(g)
    lib().oracle_smooth5_interior(g, g.shape[0], g.shape[1], out)
    return out


def log_potential(points, grid_x, grid_y, eps: float, variant: int):
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    px = _c(pts[:, 0], np.float64); py = _c(pts[:, 1], np.float64)
    gx = _c(grid_x, np.float64); gy = _c(grid_y, np.float64)
    U = np.empty((gy.size, gx.size), dtype=np.float64)
    lib().oracle_log_potential(px, py, px.size, gx, gx.size, gy, gy.size, float(eps), int(variant), U)
    return U


def nearest_match(X, Y):
    """First index of the nearest Y for every X (complex arrays): the selection rule of entropic_ot_alignment,
    tci_construct_mandelbrot_v002_fixed.py:62-71.  Returns (index int64[n], distance float64[n])."""
    X = np.asarray(X, dtype=np.complex128).ravel(); Y = np.asarray(Y, dtype=np.complex128).ravel()
    idx = np.empty(X.size, dtype=np.int64); dist = np.empty(X.size, dtype=np.float64)
    lib().oracle_nearest_match(_c(X.real, np.float64), _c(X.imag, np.float64), X.size,
                               _c(Y.real, np.float64), _c(Y.imag, np.float64), Y.size, idx, dist)
    return idx, dist


def curvature_localpoly(P, neighbors: int = 7, closed: bool = True):
    """compute_curvature_localpoly with stride 1 (boundary_curvature_localpoly.py:65-184), restated: window indices,
    signed arclength from the centre, np.linalg.lstsq quadratics, curvature formula.  Returns [N, 7] columns
    kappa, kappa_signed, speed, x', y', x'', y''."""
    P = np.asarray(P, dtype=np.float64)
    N, m = P.shape[0], int(neighbors)
    out = np.zeros((N, 7))
    for i in range(N):
        idx = [(i + dlt) % N if closed else min(max(i + dlt, 0), N - 1) for dlt in range(-m, m + 1)]
        XY = P[idx]
        s = np.zeros(2 * m + 1)
        for k in range(m + 1, 2 * m + 1):
            s[k] = s[k - 1] + np.linalg.norm(XY[k] - XY[k - 1])
        for k in range(m - 1, -1, -1):
            s[k] = s[k + 1] - np.linalg.norm(XY[k + 1] - XY[k])
        A = np.c_[np.ones_like(s), s, s ** 2]
        ax = np.linalg.lstsq(A, XY[:, 0], rcond=None)[0]
        bx = np.linalg.lstsq(A, XY[:, 1], rcond=None)[0]
        x1, x2, y1, y2 = ax[1], 2.0 * ax[2], bx[1], 2.0 * bx[2]
        sp = np.sqrt(x1 * x1 + y1 * y1) + 1e-16
        ks = (x1 * y2 - y1 * x2) / sp ** 3
        out[i] = [abs(ks), ks, sp, x1, y1, x2, y2]
    return out


def weighted_log_sum(z, nodes, weights, eps: float = 1e-300):
    """sum_n w_n log(|z_m - zeta_n| + eps): the matrix-vector product inside g_real,
    lucas_to_cardioid_v40_reference.py:252-253, written as one numpy expression."""
    z = np.asarray(z, dtype=np.complex128).ravel(); b = np.asarray(nodes, dtype=np.complex128).ravel()
    return np.log(np.abs(z[:, None] - b[None, :]) + eps) @ np.asarray(weights, dtype=np.float64)


def weighted_cauchy_sum(z, nodes, weights, dz_eps: float = 1e-14):
    """sum_n w_n / dz with short differences clamped to dz_eps + 0j: the integral term of dPhi,
    lucas_to_cardioid_v40_reference.py:207-211."""
    z = np.asarray(z, dtype=np.complex128).ravel(); b = np.asarray(nodes, dtype=np.complex128).ravel()
    DZ = z[:, None] - b[None, :]
    DZ = np.where(np.abs(DZ) < dz_eps, dz_eps + 0j, DZ)
    return (np.asarray(weights, dtype=np.float64)[None, :] / DZ).sum(axis=1)


# ---- pair statistics --------------------------------------------------
def pair_histogram(locs, lo, hi, values=None, weight: str = "none"):
    """counts / weight sums of the pairs i<j with lo[k] <= d_ij < hi[k] (every pair tested against every bin, like the
    reference's masks) -> (counts uint64[nb], sums float64[nb], D.max()).  weight: "none" | "value" | "dist2"."""
    P = np.asarray(locs, dtype=np.float64).reshape(-1, 2)
    x = _c(P[:, 0], np.float64); y = _c(P[:, 1], np.float64)
    lo = _c(lo, np.float64); hi = _c(hi, np.float64)
    wmode = {"none": 0, "value": 1, "dist2": 2}[weight]
    v = _c(values, np.float64) if wmode == 1 else None
    counts = np.zeros(lo.size, dtype=np.uint64); sums = np.zeros(lo.size, dtype=np.float64)
    dmax = lib().oracle_pair_histogram(x, y, v.ctypes.data if v is not None else None, x.size, lo, hi, lo.size, wmode,
                                       counts, sums)
    return counts, sums, float(dmax)


def _variogram_from_bins(locs, nbins, max_dist, factor, values, weight):
    P = np.asarray(locs, dtype=np.float64).reshape(-1, 2)
    if max_dist is None:
        _, _, dmax = pair_histogram(P, [0.0], [np.inf])
        max_dist = factor * dmax
    bins = np.linspace(0.0, max_dist, nbins + 1)
    centers = 0.5 * (bins[:-1] + bins[1:])
    counts, sums, _ = pair_histogram(P, bins[:-1], bins[1:], values, weight)
    gamma = np.full(nbins, np.nan)
    nz = counts > 0
    gamma[nz] = 0.5 * (sums[nz] / counts[nz])
    return centers, gamma, counts.astype(int)


def empirical_variogram_field(locs, values, nbins: int = 50, max_dist=None, max_dist_factor: float = 0.5):
    """empirical_variogram_field, Variogram-Mandelbrot-Construct.py:106-130 (= empirical_variogram_from_field_locs with
    values, Iterative_Variogram_Laplacian.py:53-86)."""
    if np.asarray(locs).shape[0] < 2:
        return np.array([]), np.array([]), np.array([])
    return _variogram_from_bins(locs, nbins, max_dist, max_dist_factor, np.asarray(values, dtype=np.float64), "value")


def empirical_variogram_coords(locs, nbins: int = 50, max_dist=None, max_dist_factor: float = 0.5):
    """empirical_variogram_coords, Variogram-Mandelbrot-Construct.py:132-152."""
    return _variogram_from_bins(locs, nbins, max_dist, max_dist_factor, None, "dist2")


def pair_correlation(points, r_max: float, dr: float):
    """pair_correlation, spatial_stats_phase2.py:9-28."""
    P = np.asarray(points, dtype=np.float64)
    N = len(P)
    area = (np.max(P[:, 0]) - np.min(P[:, 0])) * (np.max(P[:, 1]) - np.min(P[:, 1]))
    rho = N / area
    r_vals = np.arange(0, r_max, dr)
    counts, _, _ = pair_histogram(P, r_vals, r_vals + dr)
    g_r = []
    for r, count in zip(r_vals, counts.astype(np.int64)):
        norm = 2 * np.pi * r * dr * N * rho
        g_r.append(count / norm if norm > 0 else 0)
    return r_vals, np.array(g_r)


def ripley_K(points, r_max: float, dr: float):
    """ripley_K, spatial_stats_phase2.py:30-47: count(d < r) = pairs in the bins [r_j, r_{j+1}) below r."""
    P = np.asarray(points, dtype=np.float64)
    N = len(P)
    area = (np.max(P[:, 0]) - np.min(P[:, 0])) * (np.max(P[:, 1]) - np.min(P[:, 1]))
    rho = N / area
    r_vals = np.arange(0, r_max, dr)
    counts, _, _ = pair_histogram(P, r_vals, np.append(r_vals[1:], np.inf))
    below = np.concatenate([[0], np.cumsum(counts.astype(np.int64))[:-1]])
    return r_vals, np.array([(2 * c) / (N * rho) for c in below])


# ---- the tracker's density stage ------------------------------------------
def mollified_histogram(domain, eps: float, cloud, bins: int, sigma_bins: float) -> np.ndarray:
    """mollified_histogram, gi_assumption_tracker_v3.py:109-127 (mod.domain / mod.eps passed explicitly): the same
    numpy / scipy calls -- they are the reference's arithmetic for this function."""
    cloud = np.asarray(cloud, dtype=np.complex128)
    H, _, _ = np.histogram2d(cloud.real, cloud.imag, bins=(bins, bins),
                             range=[[domain[0], domain[1]], [domain[2], domain[3]]])
    H = np.maximum(H, eps)
    if sigma_bins and sigma_bins > 0:
        from scipy.ndimage import gaussian_filter
        H = gaussian_filter(H, sigma=float(sigma_bins), mode="nearest")
        H = np.maximum(H, eps)
    return H / H.sum()


def tv_distance(p, q) -> float:
    """gi_assumption_tracker_v3.py:91-92."""
    return 0.5 * float(np.sum(np.abs(p - q)))


def overlap_mass(p, q) -> float:
    """gi_assumption_tracker_v3.py:95-96."""
    return float(np.sum(np.minimum(p, q)))


def KL(P, X, eps: float = 1e-12) -> float:
    """KL of the stock module, tci_construct_mandelbrot_v002_fixed.py:84-86."""
    P_ = np.clip(P, eps, None); X_ = np.clip(X, eps, None)
    return float(np.sum(P_ * (np.log(P_) - np.log(X_))))


def gi_flow(P_target, X0, alpha: float, max_steps: int, min_steps: int = 1, kl_threshold=None, eps: float = 1e-12):
    """gi_flow_to_threshold (kl_threshold given) / gi_flow_fixed_T (None: exactly max_steps sweeps),
    gi_assumption_tracker_v3.py:130-151 -> (X_T, T, kl0, klT)."""
    X = np.array(X0, dtype=np.float64, copy=True)
    kl0 = KL(P_target, X, eps)
    kl, T = kl0, 0
    for t in range(1, int(max_steps) + 1):
        X = (1.0 - alpha) * X + alpha * P_target
        T = t
        if kl_threshold is None:
            continue
        kl = KL(P_target, X, eps)
        if t >= int(min_steps) and kl <= float(kl_threshold):
            break
    if kl_threshold is None:
        kl = KL(P_target, X, eps)
    return X, T, kl0, kl


# ---- alpha-shape edge filter ------------------------------------------------
def circumradius(p, q, r) -> float:
    """circumradius, construct_boundary_alpha.py:45-55 (same numpy calls: np.linalg.norm is the reference's arithmetic)."""
    a = np.linalg.norm(q - r)
    b = np.linalg.norm(p - r)
    c = np.linalg.norm(p - q)
    s = (a + b + c) / 2.0
    A = max(s * (s - a) * (s - b) * (s - c), 0.0)
    if A == 0.0:
        return np.inf
    area = np.sqrt(A)
    return (a * b * c) / (4.0 * area + 1e-16)


def alpha_shape_edges(P, simplices, alpha: float):
    """alpha_shape_edges, construct_boundary_alpha.py:57-82, for given Delaunay simplices -> (keep bool[T], radius[T],
    boundary edges [(i, j)] in dict order)."""
    P = np.asarray(P, dtype=np.float64)
    inv_alpha = 1.0 / alpha
    radius = np.array([circumradius(P[t[0]], P[t[1]], P[t[2]]) for t in simplices], dtype=np.float64).reshape(-1)
    keep = radius < inv_alpha
    edge_count: dict = {}
    for t in np.asarray(simplices)[keep]:
        for i, j in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            key = (int(i), int(j)) if i < j else (int(j), int(i))
            edge_count[key] = edge_count.get(key, 0) + 1
    return keep, radius, [e for e, c in edge_count.items() if c == 1]


# ---- K2 -------------------------------------------------------------------------------
def contour_lines(xs, ys, Z, level: float):
    """plt.contour(xs, ys, Z, levels=[level]).allsegs[0] restated (mpl2014) -> list of (N,2) arrays."""
    xs = _c(xs, np.float64); ys = _c(ys, np.float64); Z = _c(Z, np.float64)
    assert Z.shape == (ys.size, xs.size)
    cap_v, cap_l = 1 << 16, 1 << 12
    while True:
        verts = np.empty((cap_v, 2), dtype=np.float64)
        offs = np.empty(cap_l + 1, dtype=np.int64)
        nv = C.c_int64(0); nl = C.c_int64(0)
        rc = lib().oracle_contour_lines(xs, xs.size, ys, ys.size, Z, float(level),
                                        verts.reshape(-1), cap_v, C.byref(nv), offs, cap_l, C.byref(nl))
        if rc < 0:
            raise MemoryError("oracle_contour_lines")
        if rc == 0:
            break
        cap_v = max(cap_v, nv.value + 16); cap_l = max(cap_l, nl.value + 16)
    return [verts[offs[k]:offs[k + 1]].copy() for k in range(nl.value)]


def extract_contour(xs, ys, Z, max_iter: int, level_frac: float = 0.96):
    """extract_contour, mandelbrot_boundary_sample_spyder.py:35-43 (longest line)."""
    segs = contour_lines(xs, ys, Z, level_frac * max_iter)
    if not segs:
        return None
    return max(segs, key=lambda a: a.shape[0])


# ---- K3 (numpy = the reference's own arithmetic) -----------------------------------------
def companion_from_toprow(top) -> np.ndarray:
    """generate_companion_from_toprow, lucas_equipotential_test_v3.py:66-74."""
    top = np.asarray(top, dtype=float).reshape(-1)
    n = top.shape[0]
    Cm = np.zeros((n, n), dtype=float)
    Cm[0, :] = top
    for i in range(1, n):
        Cm[i, i - 1] = 1.0
    return Cm


def family_toprow(name: str, n: int) -> np.ndarray:
    """family_toprow, lucas_equipotential_test_v3.py:76-91."""
    if name == "lucas_all_ones":
        return np.ones(n)
    if name == "pell_like_all_twos":
        return 2.0 * np.ones(n)
    if name == "sparser_gap_1_0_1_then_ones":
        top = np.ones(n)
        if n >= 2:
            top[1] = 0.0
        return top
    if name == "padovan_like_0_1_then_ones":
        top = np.ones(n)
        top[0] = 0.0
        return top
    raise ValueError(f"Unknown family '{name}'")


def eigvals_toprow(top) -> np.ndarray:
    return np.linalg.eigvals(companion_from_toprow(top))


def inverse_eigenvalues_toprow(top, tol: float = 1e-12) -> np.ndarray:
    """One n of compute_inverse_eigenvalues_family, lucas_equipotential_test_v3.py:106-118."""
    eigs = eigvals_toprow(top)
    return (1.0 / eigs[np.abs(eigs) > tol]).astype(np.complex128)


def compute_inverse_eigenvalues_family(family: str, n_min: int, n_max: int, tol: float = 1e-12) -> np.ndarray:
    out = [inverse_eigenvalues_toprow(family_toprow(family, n), tol) for n in range(n_min, n_max + 1)]
    return np.concatenate(out).astype(np.complex128)


def roots_batched(toprows: np.ndarray, deg: np.ndarray):
    """eigvals for a zero-padded batch; returns list of complex arrays (one per polynomial)."""
    return [eigvals_toprow(toprows[k, : deg[k]]) for k in range(len(deg))]

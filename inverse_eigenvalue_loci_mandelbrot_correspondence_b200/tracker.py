"""Host side of the tracker's density stage (second half): the functions gi_assumption_tracker_v3.py
defines around its --module plug-in, with the same names, arguments and return values:

  tv_distance(p, q), overlap_mass(p, q), fraction_outside_domain(cloud, domain)      gi_assumption_tracker_v3.py:91-104
  mollified_histogram(mod, cloud, bins, sigma_bins)                                  gi_assumption_tracker_v3.py:109-127
  gi_flow_fixed_T(KL_fn, P_target, X0, alpha, T)                                     gi_assumption_tracker_v3.py:130-136
  gi_flow_to_threshold(KL_fn, P_target, X0, alpha, kl_threshold, max_steps, min_steps)   :139-151
  KL(P, X) / make_KL(eps)       the stock module's KL, tci_construct_mandelbrot_v002_fixed.py:84-86

Histogram, blur, normalisation, divergences and the whole GI flow (up to 800 sweeps over bins^2 cells with a KL after
each) run in liblm_b200.so (csrc/lm_density.cu).  Bit-identical to numpy / scipy except for the logarithm inside KL
(values agree to ~1e-15 absolute; X_T and T_n are exact).  There is no CPU fallback: the flow functions accept only a
KL_fn created here (they need its eps, and the KL itself is evaluated on the device).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import Stats

last_stats: dict = {}


def _set_stats(st: Stats) -> None:
    global last_stats
    last_stats = st.as_dict()


def _flat(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64).ravel()


def gaussian_kernel1d(sigma: float, truncate: float = 4.0):
    """scipy.ndimage's 1-D kernel for order 0 (_gaussian_kernel1d): radius = int(truncate*sigma + 0.5), normalised."""
    sigma = float(sigma)
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def histogram2d(x, y, bins, range):
    """np.histogram2d(x, y, bins=(bx, by), range=[[x0, x1], [y0, y1]]) for integer bins -> (H, xedges, yedges)."""
    bx, by = (int(bins), int(bins)) if np.ndim(bins) == 0 else (int(bins[0]), int(bins[1]))
    xe = np.linspace(float(range[0][0]), float(range[0][1]), bx + 1)
    ye = np.linspace(float(range[1][0]), float(range[1][1]), by + 1)
    x = _flat(x); y = _flat(y)
    if x.size != y.size:
        raise ValueError("x and y must have the same length")
    H = np.empty((bx, by), dtype=np.float64)
    st = Stats()
    _shim.call("lm_histogram2d", _shim.ptr(x), _shim.ptr(y), x.size, _shim.ptr(xe), bx, _shim.ptr(ye), by, _shim.ptr(H), C.byref(st))
    _set_stats(st)
    return H, xe, ye


def gaussian_filter(H, sigma: float, mode: str = "nearest", truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage.gaussian_filter for a 2-D float64 array, mode="nearest" (what the tracker uses)."""
    if mode != "nearest":
        raise ValueError("only mode='nearest' is implemented (gi_assumption_tracker_v3.py:122)")
    H = np.ascontiguousarray(H, dtype=np.float64)
    if H.ndim != 2:
        raise ValueError("2-D array expected")
    w, radius = gaussian_kernel1d(sigma, truncate)
    if radius < 1:
        return H.copy()
    out = np.empty_like(H)
    st = Stats()
    _shim.call("lm_gaussian_filter_nearest", _shim.ptr(H), H.shape[0], H.shape[1], _shim.ptr(w), radius, _shim.ptr(out), C.byref(st))
    _set_stats(st)
    return out


def sum_pairwise(a) -> float:
    """np.sum of a float64 array on the device, in numpy's pairwise order."""
    a = _flat(a)
    out = C.c_double(0.0)
    st = Stats()
    _shim.call("lm_sum_pairwise", _shim.ptr(a), a.size, C.byref(out), C.byref(st))
    _set_stats(st)
    return float(out.value)


def mollified_histogram(mod, cloud, bins: int, sigma_bins: float) -> np.ndarray:
    cloud = np.asarray(cloud, dtype=np.complex128).ravel()
    bins = int(bins)
    xe = np.linspace(float(mod.domain[0]), float(mod.domain[1]), bins + 1)
    ye = np.linspace(float(mod.domain[2]), float(mod.domain[3]), bins + 1)
    eps = float(getattr(mod, "eps", 1e-12))
    w, radius = (None, 0)
    if sigma_bins and sigma_bins > 0:
        w, radius = gaussian_kernel1d(float(sigma_bins))
    x = np.ascontiguousarray(cloud.real); y = np.ascontiguousarray(cloud.imag)
    P = np.empty((bins, bins), dtype=np.float64)
    st = Stats()
    _shim.call("lm_mollified_histogram", _shim.ptr(x), _shim.ptr(y), x.size, _shim.ptr(xe), bins, _shim.ptr(ye), bins, eps,
               _shim.ptr(w), radius, _shim.ptr(P), C.byref(st))
    _set_stats(st)
    return P


def density_compare(p, q, eps: float = 1e-12):
    """(sum|p-q|, sum min(p,q), KL(p, q)) in one device pass."""
    p = _flat(p); q = _flat(q)
    if p.size != q.size:
        raise ValueError("p and q must have the same size")
    a, b, c = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
    st = Stats()
    _shim.call("lm_density_compare", _shim.ptr(p), _shim.ptr(q), p.size, float(eps), C.byref(a), C.byref(b), C.byref(c), C.byref(st))
    _set_stats(st)
    return float(a.value), float(b.value), float(c.value)


def tv_distance(p, q) -> float:
    return 0.5 * density_compare(p, q)[0]


def overlap_mass(p, q) -> float:
    return density_compare(p, q)[1]


def fraction_outside_domain(cloud, domain) -> float:
    xmin, xmax, ymin, ymax = domain
    cloud = np.asarray(cloud)
    x = cloud.real
    y = cloud.imag
    inside = (x >= xmin) & (x <= xmax) & (y >= ymin) & (y <= ymax)
    return float(1.0 - np.mean(inside))


def make_KL(eps: float = 1e-12):
    """KL(P, X) = sum P_ (log P_ - log X_), P_ = clip(P, eps), X_ = clip(X, eps), evaluated on the device."""
    def KL(P, X):
        return density_compare(P, X, eps)[2]
    KL.lm_eps = float(eps)
    return KL


KL = make_KL(1e-12)


def _flow(KL_fn, P_target, X0, alpha, max_steps, min_steps, kl_threshold, fixed):
    eps = getattr(KL_fn, "lm_eps", None)
    eps = eps() if callable(eps) else eps
    if eps is None:
        raise TypeError("the device GI flow evaluates the stock module's KL itself: pass tracker.KL / tracker.make_KL(eps) "
                        "(or the KL of tci_construct_mandelbrot_b200) as KL_fn")
    P = np.ascontiguousarray(P_target, dtype=np.float64)
    X = np.ascontiguousarray(X0, dtype=np.float64)
    if P.shape != X.shape:
        raise ValueError("P_target and X0 must have the same shape")
    out = np.empty_like(X)
    T = C.c_int32(0)
    kl0, klT = C.c_double(0.0), C.c_double(0.0)
    st = Stats()
    _shim.call("lm_gi_flow", _shim.ptr(P.ravel()), _shim.ptr(X.ravel()), P.size, float(alpha), float(eps), int(max_steps),
               int(min_steps), float(kl_threshold), int(fixed), _shim.ptr(out.reshape(-1)), C.byref(T), C.byref(kl0), C.byref(klT),
               None, C.byref(st))
    _set_stats(st)
    return out, int(T.value), float(kl0.value), float(klT.value)


def gi_flow_fixed_T(KL_fn, P_target, X0, alpha: float, T: int):
    return _flow(KL_fn, P_target, X0, alpha, int(T), 1, 0.0, True)


def gi_flow_to_threshold(KL_fn, P_target, X0, alpha: float, kl_threshold: float, max_steps: int, min_steps: int = 1):
    return _flow(KL_fn, P_target, X0, alpha, int(max_steps), int(min_steps), float(kl_threshold), False)

"""Host side of the dense boundary-integral sums of the Lucas-domain Green function.

  RiemannMapDisk_GreenModulus.g_real(z)   lucas_to_cardioid_v40_reference.py:240-257
  RiemannMapDisk_GreenModulus.dPhi(z)     lucas_to_cardioid_v40_reference.py:201-211

The O(M*N) sums over the boundary nodes run on the GPU (liblm_b200.so: lm_weighted_log_sum,
lm_weighted_cauchy_sum); the O(M) pole / constant terms are added here exactly as the reference writes them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import Stats

last_stats: dict = {}


def _split(z):
    z = np.asarray(z, dtype=np.complex128).ravel()
    return z, np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)


def weighted_log_sum(z, nodes, weights, eps: float = 1e-300) -> np.ndarray:
    """sum_n weights[n] * log(|z_m - nodes[n]| + eps) for every z_m."""
    z, zr, zi = _split(z)
    _, br, bi = _split(nodes)
    w = np.ascontiguousarray(weights, dtype=np.float64).ravel()
    if w.size != br.size:
        raise ValueError("one weight per node expected")
    out = np.empty(z.size, dtype=np.float64)
    st = Stats()
    _shim.call("lm_weighted_log_sum", _shim.ptr(zr), _shim.ptr(zi), z.size, _shim.ptr(br), _shim.ptr(bi), _shim.ptr(w), w.size,
               float(eps), _shim.ptr(out), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return out


def weighted_cauchy_sum(z, nodes, weights, dz_eps: float = 1e-14) -> np.ndarray:
    """sum_n weights[n] / (z_m - nodes[n]) (differences shorter than dz_eps replaced by dz_eps + 0j), complex."""
    z, zr, zi = _split(z)
    _, br, bi = _split(nodes)
    w = np.ascontiguousarray(weights, dtype=np.float64).ravel()
    if w.size != br.size:
        raise ValueError("one weight per node expected")
    ore = np.empty(z.size, dtype=np.float64); oim = np.empty(z.size, dtype=np.float64)
    st = Stats()
    _shim.call("lm_weighted_cauchy_sum", _shim.ptr(zr), _shim.ptr(zi), z.size, _shim.ptr(br), _shim.ptr(bi), _shim.ptr(w), w.size,
               float(dz_eps), _shim.ptr(ore), _shim.ptr(oim), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return ore + 1j * oim


def g_real(z, bdy_z, sigma, ds, a: complex, C_const: float, g_shift: float = 0.0) -> np.ndarray:
    """g(z) = -log|z-a| + sum_n sigma_n ds_n log|z - zeta_n| + C + g_shift   (g_real, :240-257)."""
    z = np.asarray(z, dtype=np.complex128).ravel()
    sigw = (np.asarray(sigma) * np.asarray(ds)).astype(float)
    sl = weighted_log_sum(z, bdy_z, sigw, 1e-300)
    return -np.log(np.abs(z - a) + 1e-300) + sl + C_const + g_shift


def dPhi(z, bdy_z, sigma, ds, a: complex, dz_eps: float = 1e-14) -> np.ndarray:
    """Phi'(z) = -1/(z-a) + sum_n sigma_n ds_n / (z - zeta_n)   (dPhi, :201-211)."""
    z = np.asarray(z, dtype=np.complex128).ravel()
    DZ0 = z - a
    DZ0 = np.where(np.abs(DZ0) < dz_eps, dz_eps + 0j, DZ0)
    return -1.0 / DZ0 + weighted_cauchy_sum(z, bdy_z, np.asarray(sigma) * np.asarray(ds), dz_eps)

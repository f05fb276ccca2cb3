#!/usr/bin/env python3
"""GPU-backed module for `gi_assumption_tracker_v3.py --module <this file>`.

The tracker loads a script by path (gi_assumption_tracker_v3.py:84-90, 193), overwrites the
attributes `domain`, `mandelbrot_grid`, `mandelbrot_samples` per level (:194, 208-209) and calls
`construct_points(ns)`, `sample_mandelbrot_boundary()`, `entropic_ot_alignment(C, M)`,
`procrustes_align_no_scale(X, Y)`, `KL(P, X)` and reads `eps` (:212-228, 109-120).  The stock module is
tci_construct_mandelbrot_v002_fixed.py; this one keeps its attribute names, defaults and call
signatures and moves the two generators onto the B200:

  construct_points(ns)            -> K3  lm_roots_batched            (…_v002_fixed.py:27-33)
  mandelbrot_distance_estimator   -> K1b lm_distance_grid_f64, LM_DE_FINAL_DZ_NUMPY (…_v002_fixed.py:35-47 as numpy
                                     evaluates it: FMA complex multiply; escape mask and d == 0 pattern bit-exact)
  sample_mandelbrot_boundary()    -> K1b + the same quantile mask / np.random.choice on the host
                                     (…_v002_fixed.py:49-59), so a seeded run draws the same sample
  entropic_ot_alignment(X, Y)     -> lm_nearest_match: the argmax of exp(-cdist/const) is the nearest
                                     point (…_v002_fixed.py:62-71), same np.random.choice subsampling
  to_prob(cloud, bins), KL(P, X)  -> lm_mollified_histogram / lm_density_compare (…_v002_fixed.py:78-86);
                                     KL carries `lm_eps`, so tracker.gi_flow_* run the GI flow on the device

The Procrustes helper is host numpy on <= 4*10^4 points, as in the stock module.
Differences a user can see: inside one n the cloud is sorted by (re, im) instead of in
LAPACK's order, and `mandelbrot_distance_estimator` wants a meshgrid (`X + 1j*Y`) and returns
`last = None` (z of the first escape stays on the device).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

_ROOT = Path(__file__).resolve().parents[1]
if str(_ROOT) not in sys.path:           # loaded by file path, not as a package member
    sys.path.insert(0, str(_ROOT))

from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import lucas as _lucas  # noqa: E402
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import potentials as _potentials  # noqa: E402
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import tracker as _tracker  # noqa: E402

# ---------- CONFIG (names and defaults of tci_construct_mandelbrot_v002_fixed.py:12-22) ----------
np.random.seed(7)
construct_ns = list(range(20, 301, 20))
mandelbrot_grid = 600
mandelbrot_samples = 25000
escape_R, max_iter = 250, 250
grid_bins = 128
domain = (-2.25, 1.25, -1.75, 1.75)
alpha, T, eps = 0.2, 60, 1e-12
sinkhorn_eps, sinkhorn_iter = 0.8, 600
# ---------------------------------------------------------------------------------------------------

lucas_companion = _lucas.lucas_companion


def construct_points(ns):
    """Lucas Loci for the given orders, |lambda| > 1e-10 kept (K3 on the GPU)."""
    return _lucas.construct_points(ns, tol=1e-10)


def _axes_of_meshgrid(c: np.ndarray):
    c = np.asarray(c, dtype=np.complex128)
    if c.ndim != 2:
        raise ValueError("mandelbrot_distance_estimator expects a 2-D meshgrid X + 1j*Y")
    xs = np.ascontiguousarray(c[0, :].real)
    ys = np.ascontiguousarray(c[:, 0].imag)
    if not (np.array_equal(c.real, np.broadcast_to(xs[None, :], c.shape)) and
            np.array_equal(c.imag, np.broadcast_to(ys[:, None], c.shape))):
        raise ValueError("mandelbrot_distance_estimator expects a 2-D meshgrid X + 1j*Y")
    return xs, ys


def mandelbrot_distance_estimator(c):
    """(escaped mask, distance estimate, None) on the meshgrid c; module-level max_iter/escape_R/eps apply."""
    xs, ys = _axes_of_meshgrid(c)
    d, esc = _potentials.distance_grid(xs, ys, int(max_iter), float(escape_R), float(eps), _potentials.DE_FINAL_DZ_NUMPY)
    return esc, d, None


def sample_mandelbrot_boundary():
    xs = np.linspace(domain[0], domain[1], mandelbrot_grid)
    ys = np.linspace(domain[2], domain[3], mandelbrot_grid)
    d, esc = _potentials.distance_grid(xs, ys, int(max_iter), float(escape_R), float(eps), _potentials.DE_FINAL_DZ_NUMPY)
    if not esc.any():
        raise RuntimeError("No escape points")
    q = np.quantile(d[esc], 0.25)
    jj, ii = np.nonzero(esc & (d <= q))                   # row-major, the order C[mask].ravel() has
    pts = xs[ii] + 1j * ys[jj]
    if pts.size > mandelbrot_samples:
        pts = np.random.choice(pts, mandelbrot_samples, replace=False)
    return pts


def entropic_ot_alignment(X, Y):
    """The stock module's 'simplified Sinkhorn': equalise the sizes by random subsampling (same draws from
    numpy's global stream), then match every X to the Y with the largest exp(-dist/(mean dist * sinkhorn_eps)),
    i.e. its nearest Y, first index on ties -- on the GPU (lm_nearest_match), so the n x m distance matrix the
    stock module builds (5 GB at 25 000 points) never exists."""
    X = np.asarray(X); Y = np.asarray(Y)
    n, m = len(X), len(Y)
    if n > m:
        X = np.random.choice(X, m, replace=False)
    if m > n:
        Y = np.random.choice(Y, n, replace=False)
    if len(X) == 0:
        return Y[:0], X
    match, _ = _potentials.nearest_match(X, Y)
    return Y[match], X


def procrustes_align_no_scale(Xc, Yc):
    """Rotation (no scaling) of the centred X onto the centred Y, translated to Y's centroid."""
    A = np.column_stack([Xc.real, Xc.imag]); B = np.column_stack([Yc.real, Yc.imag])
    muA, muB = A.mean(axis=0), B.mean(axis=0)
    U, _, Vt = np.linalg.svd((B - muB).T @ (A - muA), full_matrices=False)
    out = (A - muA) @ (U @ Vt) + muB
    return out[:, 0] + 1j * out[:, 1]


class _Self:
    """what tracker.mollified_histogram reads from a module: the current `domain` and `eps` of this one"""
    domain = property(lambda self: domain)
    eps = property(lambda self: eps)


def to_prob(cloud, bins=grid_bins):
    """histogram2d -> max(eps) -> / sum on the device (…_v002_fixed.py:78-82; bit-identical to numpy)."""
    return _tracker.mollified_histogram(_Self(), cloud, bins, 0.0)


def KL(P, X):
    """sum P_ (log P_ - log X_) with both clipped at eps, on the device (…_v002_fixed.py:84-86)."""
    return _tracker.density_compare(P, X, eps)[2]


KL.lm_eps = lambda: eps      # lets tracker.gi_flow_* run the whole flow on the device with this module's KL


def tci_flow(P, X0):
    X = X0.copy()
    kls, traj = [KL(P, X)], [X]
    for _ in range(T):
        X = (1 - alpha) * X + alpha * P
        kls.append(KL(P, X)); traj.append(X.copy())
    return np.array(kls), traj


if __name__ == "__main__":
    import json
    import time
    t0 = time.time()
    Cpts = construct_points(construct_ns)
    Mpts = sample_mandelbrot_boundary()
    Mmatch, Ctrim = entropic_ot_alignment(Cpts, Mpts)
    Caligned = procrustes_align_no_scale(Ctrim, Mmatch)
    kls, _ = tci_flow(to_prob(Mpts), to_prob(Caligned))
    out = {"n_construct_pts": int(Cpts.size), "n_mandel_pts": int(Mpts.size), "KL_initial": float(kls[0]),
           "KL_final": float(kls[-1]), "runtime_sec": time.time() - t0}
    json.dump(out, open("tci_results.json", "w"), indent=2)
    print("Done. Results:", out)

"""Small instances of every kernel family, for compute-sanitizer (one --tool per gpurun call)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import (alpha_shape, contour, curvature, escape, lucas, nystrom, pairstats, potentials,
                                                                    stencils, tracker)
import types

rng = np.random.default_rng(0)
xs = np.linspace(-2.1, 0.9, 256); ys = np.linspace(-1.5, 1.5, 200)
lines, st = contour.boundary_sample(xs, ys, 120, 115.2)                       # K1 chunked + bulk-copy mark + emit + link
d, f, _ = escape.escape_grid(np.linspace(-2.1, 0.9, 301), ys, 120, 2.0, 1)    # field mode, ragged width
l2 = contour.contour_lines(np.linspace(-2.1, 0.9, 301), ys, d, 115.2)         # register-load mark kernel
pts = rng.uniform(-2.2, 1.0, 20000) + 1j * rng.uniform(-1.5, 1.5, 20000)
g, it, phi = escape.batch_potential(pts, 600, 2.0)                            # two-pass point schedule
deg = rng.integers(2, 26, size=2000).astype(np.int32)
top = rng.integers(0, 3, size=(2000, 25)).astype(np.float64)
top[np.arange(25)[None, :] >= deg[:, None]] = 0.0
top[np.arange(2000), deg - 1] = np.maximum(top[np.arange(2000), deg - 1], 1.0)
gx = np.linspace(-2, 2, 50)
out = lucas.cloud_fields(top, deg, gx, gx, potential=(300, 2.0))              # K3 generations, compaction, K1d, K4a, K4
hi = lucas.construct_points([40, 130])                                        # 8-lane and 32-lane solver classes
U = potentials._logpot(pts.real[:3000], pts.imag[:3000], gx, gx, 1e-6, 3)     # general-eps K4a path
L = stencils.laplacian(rng.standard_normal((64, 96)), 0.1); S = stencils.smooth5(rng.standard_normal((33, 41)))
dd, esc = potentials.distance_grid(xs[:100], ys[:80], 100, 250.0, 1e-12, 2)
i, dist = potentials.nearest_match(pts[:3000], pts[3000:5000])
w = nystrom.weighted_log_sum(pts[:500], pts[500:900], rng.uniform(0, 1, 400))
c = nystrom.weighted_cauchy_sum(pts[:500], pts[500:900], rng.uniform(0, 1, 400))
k = curvature.compute_curvature_localpoly(contour.longest(lines), 7, True)
out8 = lucas.cloud_fields(top.astype(np.int8), deg)                             # int8 first rows, widened on the device
P2 = np.c_[pts.real[:1500], pts.imag[:1500]]
e = np.linspace(0, 1.5, 61)
for wt in ("none", "value", "dist2"):
    pairstats.pair_histogram(P2, e[:-1], e[1:], P2[:, 0], wt)                 # partition bins
pairstats.pair_histogram(P2, e[:-1], e[:-1] + 0.03, P2[:, 0], "value")        # overlapping shells (general edges)
ee = np.linspace(0, 3.0, 1801)
pairstats.pair_histogram(P2, ee[:-1], ee[1:], None, "dist2")                  # 1800 bins: a single histogram copy
dm = pairstats.max_pair_distance(P2)
mod = types.SimpleNamespace(domain=(-2.2, 1.2, -1.6, 1.6), eps=1e-12)
PM = tracker.mollified_histogram(mod, pts, 96, 1.0); PC = tracker.mollified_histogram(mod, pts[:4000] * 0.9, 96, 2.5)
XT, T, kl0, klT = tracker.gi_flow_to_threshold(tracker.KL, PM, PC, 0.1, 1e-6, 200, 5)
tv = tracker.tv_distance(PM, PC); sm = tracker.sum_pairwise(rng.standard_normal(12345))
tri = alpha_shape.delaunay_simplices(P2)
edges = alpha_shape.alpha_shape_edges(P2, 8.0, simplices=tri)
# round 2: device linker on a field full of saddles / border-cut lines, fp32 K1, ordered pair selection, stats drop-ins
fld = rng.integers(0, 3, size=(97, 130)).astype(np.int32)
ll = contour.contour_lines(np.linspace(0, 1, 130), np.linspace(0, 1, 97), fld, 0.5)
d32, _ = escape.escape_grid_f32(np.linspace(-2.1, 0.9, 301), ys, 120)
grid = types.SimpleNamespace(X=np.meshgrid(np.linspace(-2, 1, 70), np.linspace(-1.4, 1.4, 66))[0],
                             Y=np.meshgrid(np.linspace(-2, 1, 70), np.linspace(-1.4, 1.4, 66))[1])
np.random.seed(5)
rcs, gam = pairstats.sample_semivariogram(np.sin(grid.X) * np.cos(grid.Y), grid, np.linspace(0, 2, 11), max_pairs_per_bin=5000)
rows = lucas.per_n_stats(2, 12, max_iter=300, quiet=True)
print("round-2 rows ok:", len(ll), int(d32.sum()), float(gam.sum()), len(rows))
print("new rows ok:", out8["n_points"], dm, T, klT, tv, len(edges))
print("sanitize_small ok:", len(lines), len(l2), int(it.sum()), out["n_points"], hi.size, float(U.sum()), i[:3], float(w[0]))

#!/usr/bin/env python3
"""Generate tests/golden/reference_vectors.npz by running the REFERENCE's own Python functions.

Run in the build container only (it reads /root/reference, which does not exist on the GPU
box); the resulting small fixture is committed and is what pins oracle/ and the CUDA path:

    python oracle/gen_golden.py [/root/reference]

Functions are pulled out of the reference scripts with ast (their `def` blocks only, executed
in a namespace that has numpy/math), because several scripts load CSV files or plot at import
time and matplotlib is not installed here.  Nothing from the reference is copied into the
repository: only the numeric outputs are stored.
"""
from __future__ import annotations

import ast
import math
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "reference_vectors.npz"


def load_defs(fname: str, names: list[str], extra: dict | None = None) -> dict:
    """exec the named top-level function / class definitions of a reference script."""
    src = (REF / fname).read_text().replace("\r\n", "\n")
    tree = ast.parse(src)
    picked = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise RuntimeError(f"{fname}: missing definitions {missing}")
    import numpy.linalg as la
    from dataclasses import dataclass
    from scipy.linalg import eigvals
    ns = {"np": np, "numpy": np, "math": math, "la": la, "dataclass": dataclass, "eigvals": eigvals}
    ns.update(extra or {})
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, str(REF / fname), "exec"), ns)
    return ns


def main() -> None:
    g: dict[str, np.ndarray] = {}

    # ---- K1: mandelbrot_dwell / compute_grid (mandelbrot_boundary_sample.py:22-39)
    mbs = load_defs("mandelbrot_boundary_sample.py", ["mandelbrot_dwell", "compute_grid"])
    for tag, xlim, ylim, res, mi in [("cfg1", (-2.1, 0.9), (-1.5, 1.5), 72, 500),
                                     ("seahorse", (-0.755, -0.735), (0.10, 0.12), 40, 300),
                                     ("tip", (-2.05, -1.7), (-0.05, 0.05), 33, 200)]:
        xs, ys, Z = mbs["compute_grid"](xlim, ylim, res, mi)
        g[f"dwell_{tag}_args"] = np.array([xlim[0], xlim[1], ylim[0], ylim[1], res, mi], dtype=np.float64)
        g[f"dwell_{tag}_xs"] = xs
        g[f"dwell_{tag}_ys"] = ys
        g[f"dwell_{tag}_Z"] = Z.astype(np.int32)
        assert np.array_equal(Z, Z.astype(np.int32))
    pts = np.array([[0.0, 0.0], [-2.0, 0.0], [0.25, 0.0], [0.26, 0.0], [-0.75, 0.1], [0.3, 0.5], [-1.75, 0.0],
                    [2.5, 2.5], [-0.1, 0.651], [-1.25, 0.0], [0.0, 1.0], [-0.743643887, 0.131825904]])
    g["dwell_points_xy"] = pts
    g["dwell_points_mi"] = np.array([400])
    g["dwell_points_out"] = np.array([mbs["mandelbrot_dwell"](x, y, 400) for x, y in pts], dtype=np.int32)

    # ---- K3: Lucas Loci (lucas_equipotential_test_v3.py:58-118, tci_construct_mandelbrot.py:5-19)
    lucas = load_defs("lucas_equipotential_test_v3.py",
                      ["generate_lucas_companion", "generate_companion_from_toprow", "family_toprow",
                       "compute_inverse_eigenvalues", "compute_inverse_eigenvalues_family",
                       "mandelbrot_parameter_potential", "batch_potential"], {"print": lambda *a, **k: None})
    fam_names = ["lucas_all_ones", "pell_like_all_twos", "sparser_gap_1_0_1_then_ones", "padovan_like_0_1_then_ones"]
    for fam in fam_names:
        chunks, counts = [], []
        for n in range(2, 26):
            v = lucas["compute_inverse_eigenvalues_family"](fam, n, n, 1e-12)
            chunks.append(np.sort(v)); counts.append(len(v))
        g[f"family_{fam}_values"] = np.concatenate(chunks)
        g[f"family_{fam}_counts"] = np.array(counts, dtype=np.int32)
    cloud = lucas["compute_inverse_eigenvalues"](2, 40, 1e-12)
    g["lucas_2_40_sorted_per_n"] = np.concatenate(
        [np.sort(lucas["compute_inverse_eigenvalues"](n, n, 1e-12)) for n in range(2, 41)])
    assert len(cloud) == len(g["lucas_2_40_sorted_per_n"])
    # ---- per-n / cumulative statistics (lucas_equipotential_test_v3.py:168-184, 294-327), MAX_ITER lowered to 3000
    stats = load_defs("lucas_equipotential_test_v3.py",
                      ["generate_lucas_companion", "generate_companion_from_toprow", "family_toprow",
                       "mandelbrot_parameter_potential", "batch_potential", "summarize_g", "per_n_stats", "cumulative_stats"],
                      {"print": lambda *a, **k: None, "MAX_ITER": 3000, "ESCAPE_RADIUS": 2.0, "EIG_TOL": 1e-12})
    cols = ["count", "escaped", "escaped_frac", "g_median", "g_mean", "g_std", "g_p10", "g_p90"]
    for tag, fam in (("lucas", None), ("pell", "pell_like_all_twos")):
        rows = stats["per_n_stats"](2, 30, family=fam)
        g[f"per_n_stats_{tag}_2_30_mi3000"] = np.array([[r[c] for c in cols] for r in rows], dtype=np.float64)
        rows = stats["cumulative_stats"](2, 30, family=fam)
        g[f"cumulative_stats_{tag}_2_30_mi3000"] = np.array([[r[c] for c in cols] for r in rows], dtype=np.float64)
    tci = load_defs("tci_construct_mandelbrot.py", ["lucas_companion", "construct_points"])
    cp = tci["construct_points"](range(20, 301, 20))
    g["tci_construct_points_count"] = np.array([len(cp)])            # 2400, v3_T25_sigma3_dense.csv:2
    g["tci_construct_points_n300_sorted"] = np.sort(tci["construct_points"]([300]))
    g["tci_construct_points_n20_sorted"] = np.sort(tci["construct_points"]([20]))

    # ---- K1d: batch_potential on part of the cloud + some exterior/interior points
    sample = np.concatenate([cloud[::7], np.array([0.3 + 0.5j, -0.75 + 0.1j, 0.0 + 0.0j, -2.5 + 0.1j, 0.26 + 0.0j])])
    gg, it, phi = lucas["batch_potential"](sample, max_iter=1500, escape_radius=2.0)
    g["potential_points_c"] = sample
    g["potential_points_g"] = gg
    g["potential_points_it"] = it.astype(np.int64)
    g["potential_points_phi"] = phi

    # ---- K1c: grid potentials
    gx = np.linspace(-2, 2, 28); gy = np.linspace(-2, 2, 26)
    pot = load_defs("Potentials.py", ["log_potential", "escape_potential"])
    g["potgrid_x"] = gx; g["potgrid_y"] = gy
    g["potentials_escape_R10_mi60"] = pot["escape_potential"](gx, gy, max_iter=60, R=10)
    lap = load_defs("Laplacian_C-M.py", ["construct_potential", "mandelbrot_potential", "laplacian"])
    X, Y = np.meshgrid(gx, gy)
    g["laplacian_cm_potential_R2_mi80"] = lap["mandelbrot_potential"](X, Y, max_iter=80, R=2.0)
    itv = load_defs("Iterative_Variogram_Laplacian.py", ["log_potential", "escape_potential", "laplacian_fd"])
    g["iterative_escape_R10_mi70"] = itv["escape_potential"](gx, gy, max_iter=70, R=10.0)
    vg = load_defs("variograms_construct_mandelbrot.py",
                   ["Grid", "make_grid", "log_potential_from_points", "mandelbrot_escape_potential",
                    "mandelbrot_distance_estimator"])
    # ---- sub-sampled semivariograms (variograms_construct_mandelbrot.py:178-315) with a seeded global stream:
    # an 84 x 80 grid (6720 pixels -> two chunks of 4000 / 2720 sampled points, three blocks) and caps that some
    # bins reach and others do not, so that both the "take all" and the "draw a subset" branches are recorded
    sv = load_defs("variograms_construct_mandelbrot.py",
                   ["Grid", "make_grid", "sample_semivariogram", "sample_cross_semivariogram"])
    sgrid = sv["make_grid"](-2.0, 1.0, -1.4, 1.4, 84, 80)
    rng_f = np.random.default_rng(2024)
    f1 = np.sin(3.0 * sgrid.X) * np.cos(2.0 * sgrid.Y) + 0.1 * rng_f.standard_normal(sgrid.X.shape)
    f2 = np.cos(1.5 * sgrid.X + 0.3) * np.sin(2.5 * sgrid.Y) + 0.1 * rng_f.standard_normal(sgrid.X.shape)
    g["semivario_field1"] = f1; g["semivario_field2"] = f2
    g["semivario_grid_args"] = np.array([-2.0, 1.0, -1.4, 1.4, 84, 80], dtype=np.float64)
    for tag, bins, cap in (("a", np.linspace(0.0, 2.0, 21), 20000), ("b", np.linspace(0.05, 3.5, 13), 150000)):
        g[f"semivario_{tag}_bins"] = bins
        g[f"semivario_{tag}_cap"] = np.array([cap])
        np.random.seed(777)
        rc, gam = sv["sample_semivariogram"](f1, sgrid, bins, max_pairs_per_bin=cap)
        g[f"semivario_{tag}_centers"] = rc; g[f"semivario_{tag}_gamma"] = gam
        np.random.seed(778)
        rc, gam = sv["sample_cross_semivariogram"](f1, f2, sgrid, bins, max_pairs_per_bin=cap)
        g[f"semivario_{tag}_cross_gamma"] = gam
    grid = vg["make_grid"](-2.25, 1.25, -1.75, 1.75, 30, 27)
    g["vario_grid_x"] = grid.x; g["vario_grid_y"] = grid.y
    g["vario_escape_potential_mi90"] = vg["mandelbrot_escape_potential"](grid, max_iter=90, R=4.0)
    esc, dist, lastz, lastdz = vg["mandelbrot_distance_estimator"](grid.Z, max_iter=90, R=4.0, eps=1e-14)
    g["vario_de_escaped"] = esc
    g["vario_de_dist"] = dist

    # ---- K4a: log-potentials of a small cloud
    cpts = np.column_stack([cloud.real, cloud.imag])[::11]
    g["logpot_points"] = cpts
    g["logpot_potentials"] = pot["log_potential"](cpts, gx, gy)
    g["logpot_laplacian_cm"] = lap["construct_potential"](X, Y, cpts)
    g["logpot_iterative"] = itv["log_potential"](cpts, gx, gy)
    g["logpot_vario_eps1e-6"] = vg["log_potential_from_points"](grid, cloud[::11], eps=1e-6)

    # ---- K4: stencils
    rng = np.random.default_rng(12345)
    U = rng.standard_normal((23, 31))
    g["stencil_U"] = U
    g["stencil_h"] = np.array([gx[1] - gx[0]])
    g["stencil_laplacian"] = lap["laplacian"](U, gx[1] - gx[0])
    g["stencil_laplacian_fd"] = itv["laplacian_fd"](U, gx[1] - gx[0])
    sm = U.copy()
    sm[1:-1, 1:-1] = (U[1:-1, 1:-1] + U[:-2, 1:-1] + U[2:, 1:-1] + U[1:-1, :-2] + U[1:-1, 2:]) / 5.0  # variograms_...py:169-173
    g["stencil_smooth5"] = sm

    # ---- K1b: scalar distance estimator (construct_stage1_clean.py:50-58)
    cs = load_defs("construct_stage1_clean.py", ["mandelbrot_distance_estimator", "construct_points"])
    dxs = np.linspace(-2.25, 1.25, 24); dys = np.linspace(-1.25, 1.25, 17)
    g["de_scalar_x"] = dxs; g["de_scalar_y"] = dys
    g["de_scalar_dist"] = np.array([[cs["mandelbrot_distance_estimator"](complex(x, y), max_iter=200) for x in dxs]
                                    for y in dys])
    g["stage1_construct_points_maxN12"] = cs["construct_points"](12)

    # ---- the tracker module's DE (tci_construct_mandelbrot_v002_fixed.py:35-47): dz read after the loop
    fx = load_defs("tci_construct_mandelbrot_v002_fixed.py", ["mandelbrot_distance_estimator"],
                   {"max_iter": 250, "escape_R": 250, "eps": 1e-12})
    fxs = np.linspace(-2.25, 1.25, 61); fys = np.linspace(-1.75, 1.75, 53)
    FX, FY = np.meshgrid(fxs, fys)
    esc, dd, last = fx["mandelbrot_distance_estimator"](FX + 1j * FY)
    g["tci_fixed_x"] = fxs; g["tci_fixed_y"] = fys
    g["tci_fixed_escaped"] = esc
    g["tci_fixed_dist"] = dd
    # a zoom on the boundary, where some points escape in the last iterations and dz stays finite
    zxs = np.linspace(-0.7600, -0.7400, 96); zys = np.linspace(0.0900, 0.1100, 90)
    ZX, ZY = np.meshgrid(zxs, zys)
    esc, dd, last = fx["mandelbrot_distance_estimator"](ZX + 1j * ZY)
    g["tci_fixed_zoom_x"] = zxs; g["tci_fixed_zoom_y"] = zys
    g["tci_fixed_zoom_escaped"] = esc
    g["tci_fixed_zoom_dist"] = dd

    # sample_mandelbrot_boundary() of the same module at a tracker-like level (grid 150, no subsampling)
    fxm = load_defs("tci_construct_mandelbrot_v002_fixed.py", ["mandelbrot_distance_estimator", "sample_mandelbrot_boundary"],
                    {"max_iter": 250, "escape_R": 250, "eps": 1e-12, "domain": (-2.25, 1.25, -1.75, 1.75),
                     "mandelbrot_grid": 150, "mandelbrot_samples": 10 ** 9})
    g["tci_fixed_boundary_sample_grid150"] = fxm["sample_mandelbrot_boundary"]()

    # ---- boundary-integral Green function (lucas_to_cardioid_v40_reference.py:184-257): g_real and dPhi of the
    # reference's own dataclass on a synthetic closed boundary (an ellipse with a smooth density)
    rm = load_defs("lucas_to_cardioid_v40_reference.py", ["gauss_legendre_01", "RiemannMapDisk_GreenModulus"],
                   {"PATH_GAUSS_N": 16, "EPS_POLE": 1e-6, "DZ_EPS": 1e-14, "G_CHUNK": 600})
    nb = 257
    th = 2 * np.pi * (np.arange(nb) + 0.5) / nb
    bdy = 1.3 * np.cos(th) + 0.8j * np.sin(th) + 0.1
    dsb = np.abs(np.roll(bdy, -1) - bdy)
    sig = 0.5 + 0.2 * np.cos(3 * th) + 0.05 * np.sin(th)
    obj = rm["RiemannMapDisk_GreenModulus"](bdy_z=bdy, ds=dsb, sigma=sig, a=0.15 + 0.05j, C=0.3, g_shift=-0.07)
    rng = np.random.default_rng(11)
    zt = np.concatenate([rng.uniform(-2, 2, 700) + 1j * rng.uniform(-1.5, 1.5, 700), bdy[:5], [0.15 + 0.05j]])
    g["green_bdy"] = bdy; g["green_ds"] = dsb; g["green_sigma"] = sig
    g["green_params"] = np.array([0.15, 0.05, 0.3, -0.07])
    g["green_targets"] = zt
    g["green_g_real"] = obj.g_real(zt)
    g["green_dPhi"] = obj.dPhi(zt)

    # ---- boundary consumer: local-polynomial curvature (boundary_curvature_localpoly.py:65-184) on a closed and an
    # open curve at pixel-like spacing, including a repeated closing vertex like <prefix>_boundary.csv has
    cv = load_defs("boundary_curvature_localpoly.py",
                   ["local_arclength_parameters", "quadratic_design", "fit_quadratic", "curvature_from_param_quadratic",
                    "index_window", "compute_curvature_localpoly"])
    tt = np.linspace(0, 2 * np.pi, 181)
    loop = np.c_[-0.75 + 3e-3 * (1 + 0.3 * np.cos(5 * tt)) * np.cos(tt), 0.1 + 3e-3 * (1 + 0.3 * np.cos(5 * tt)) * np.sin(tt)]
    loop[-1] = loop[0]                                       # closed polyline repeats its first vertex
    k, ks, sp, aux = cv["compute_curvature_localpoly"](loop, neighbors=7, closed=True, stride=1)
    g["curv_closed_P"] = loop
    g["curv_closed_out"] = np.c_[k, ks, sp, aux["xprime"], aux["yprime"], aux["x2"], aux["y2"]]
    arc = np.c_[np.linspace(-1, 1, 90), 0.3 * np.linspace(-1, 1, 90) ** 3]
    k, ks, sp, aux = cv["compute_curvature_localpoly"](arc, neighbors=4, closed=False, stride=3)
    g["curv_open_P"] = arc
    g["curv_open_out"] = np.c_[k, ks, sp, aux["xprime"], aux["yprime"], aux["x2"], aux["y2"]]

    # ---- pair statistics: variograms and pair correlation of a small cloud, plus a lattice whose
    # distances fall exactly on bin edges (3-4-5 triangles against integer edges)
    from scipy.spatial.distance import pdist
    from scipy.spatial import distance_matrix
    import warnings
    warnings.simplefilter("ignore", DeprecationWarning)
    vm = load_defs("Variogram-Mandelbrot-Construct.py", ["empirical_variogram_field", "empirical_variogram_coords"],
                   {"pdist": pdist, "MAX_DIST_FACTOR": 0.5})
    iv = load_defs("Iterative_Variogram_Laplacian.py", ["empirical_variogram_from_field_locs"], {"pdist": pdist})
    sp2 = load_defs("spatial_stats_phase2.py", ["pair_correlation", "ripley_K"], {"distance_matrix": distance_matrix})
    pc = np.column_stack([cloud.real, cloud.imag])[::2]                    # 410 points of the Lucas cloud
    pv = np.hypot(pc[:, 0] + 0.25, pc[:, 1]) + 0.1 * np.sin(7 * pc[:, 0])   # a smooth "matching distance"-like field
    g["pair_cloud"] = pc; g["pair_values"] = pv
    for tag, out in [("field", vm["empirical_variogram_field"](pc, pv, nbins=60)),
                     ("field_maxd", vm["empirical_variogram_field"](pc, pv, nbins=17, max_dist=0.9)),
                     ("coords", vm["empirical_variogram_coords"](pc, nbins=60)),
                     ("iter_values", iv["empirical_variogram_from_field_locs"](pc, values=pv, nbins=50)),
                     ("iter_coords", iv["empirical_variogram_from_field_locs"](pc, values=None, nbins=50))]:
        g[f"pair_vario_{tag}"] = np.vstack([out[0], out[1], out[2].astype(np.float64)])
    r_pc, g_pc = sp2["pair_correlation"](pc, 1.5, 0.01)
    r_k, K_k = sp2["ripley_K"](pc, 1.5, 0.01)
    g["pair_correlation_r1.5_dr0.01"] = np.vstack([r_pc, g_pc])
    g["pair_ripley_r1.5_dr0.01"] = np.vstack([r_k, K_k])
    lat = np.array([[i, j] for i in range(13) for j in range(11)], dtype=np.float64)
    g["pair_lattice"] = lat
    lv = (lat[:, 0] * 3 - lat[:, 1] * 2) % 7
    g["pair_lattice_values"] = lv
    out = vm["empirical_variogram_field"](lat, lv, nbins=16, max_dist=16.0)    # edges 0, 1, 2, ...: d = 5, 10, 13 sit on edges
    g["pair_lattice_vario"] = np.vstack([out[0], out[1], out[2].astype(np.float64)])
    r_l, g_l = sp2["pair_correlation"](lat, 8.0, 0.5)
    r_lk, K_l = sp2["ripley_K"](lat, 8.0, 0.5)
    g["pair_lattice_correlation"] = np.vstack([r_l, g_l])
    g["pair_lattice_ripley"] = np.vstack([r_lk, K_l])

    # ---- the tracker's density stage (gi_assumption_tracker_v3.py:91-151) with the stock module's KL
    # (tci_construct_mandelbrot_v002_fixed.py:84-86): Lucas cloud vs the module's boundary sample, as main() pairs them
    from scipy.ndimage import gaussian_filter
    from typing import Tuple
    trk = load_defs("gi_assumption_tracker_v3.py",
                    ["tv_distance", "overlap_mass", "fraction_outside_domain", "mollified_histogram", "gi_flow_fixed_T",
                     "gi_flow_to_threshold"], {"gaussian_filter": gaussian_filter, "Tuple": Tuple})
    kmod = load_defs("tci_construct_mandelbrot_v002_fixed.py", ["KL"], {"eps": 1e-12})

    class _Mod:
        domain = (-2.2, 1.2, -1.6, 1.6)         # the tracker's default --domain
        eps = 1e-12
    Mb = g["tci_fixed_boundary_sample_grid150"]
    g["density_domain_eps"] = np.array([*_Mod.domain, _Mod.eps])
    g["density_cloud_C"] = cloud
    for bins, sig in [(64, 1.0), (64, 0.0), (50, 2.5)]:
        tag = f"b{bins}_s{sig}"
        P_C = trk["mollified_histogram"](_Mod, cloud, bins, sig)
        P_M = trk["mollified_histogram"](_Mod, Mb, bins, sig)
        g[f"density_PC_{tag}"] = P_C; g[f"density_PM_{tag}"] = P_M
        X_T, Tn, kl0, klT = trk["gi_flow_to_threshold"](kmod["KL"], P_M, P_C, 0.1, 1e-6, 800, 5)
        X_F, TF, kl0f, klF = trk["gi_flow_fixed_T"](kmod["KL"], P_M, P_C, 0.1, 25)
        g[f"density_flow_XT_{tag}"] = X_T; g[f"density_flow_XF_{tag}"] = X_F
        g[f"density_scalars_{tag}"] = np.array([trk["tv_distance"](P_C, P_M), trk["overlap_mass"](P_C, P_M), kmod["KL"](P_M, P_C),
                                                Tn, kl0, klT, TF, kl0f, klF, trk["tv_distance"](X_T, P_M),
                                                trk["fraction_outside_domain"](cloud, _Mod.domain),
                                                trk["fraction_outside_domain"](Mb, _Mod.domain)])

    # ---- alpha-shape boundary of a point cloud (construct_boundary_alpha.py:45-125): the reference's functions with the
    # Delaunay simplices recorded, so that the fixture does not depend on the Qhull build of the machine running the tests
    from scipy.spatial import Delaunay as _Delaunay
    seen = {}

    def _recording_delaunay(P):
        tri = _Delaunay(P)
        seen["simplices"] = tri.simplices.copy()
        return tri
    ab = load_defs("construct_boundary_alpha.py", ["circumradius", "alpha_shape_edges", "order_boundary"], {"Delaunay": _recording_delaunay})
    rng = np.random.default_rng(77)
    tt = rng.uniform(0, 2 * np.pi, 1500)
    rr = np.sqrt(rng.uniform(0.0, 1.0, 1500)) * (1.0 + 0.35 * np.cos(3 * tt))       # a filled trefoil-like blob
    Pa = np.c_[rr * np.cos(tt), rr * np.sin(tt)]
    for tag, alpha in [("a6", 6.0), ("a12", 12.0)]:
        edges = ab["alpha_shape_edges"](Pa, alpha)
        g[f"alpha_{tag}_edges"] = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
        g[f"alpha_{tag}_ordered"] = np.asarray(ab["order_boundary"](Pa, edges), dtype=np.int32)
    g["alpha_points"] = Pa
    g["alpha_simplices"] = seen["simplices"].astype(np.int32)
    g["alpha_radius"] = np.array([ab["circumradius"](Pa[t[0]], Pa[t[1]], Pa[t[2]]) for t in seen["simplices"]])

    # ---- the two grid samplers built on the distance estimators (construct_stage1_clean.py:60-80, seeded;
    # variograms_construct_mandelbrot.py:90-104)
    st1 = load_defs("construct_stage1_clean.py", ["mandelbrot_distance_estimator", "sample_mandelbrot_boundary"])
    np.random.seed(11)
    g["stage1_boundary_sample_seed11"] = st1["sample_mandelbrot_boundary"](nx=120, ny=80, max_iter=200, nsamples=300)
    g["stage1_boundary_sample_all"] = st1["sample_mandelbrot_boundary"](nx=60, ny=40, max_iter=150, nsamples=10 ** 6)
    vb = load_defs("variograms_construct_mandelbrot.py", ["mandelbrot_distance_estimator", "mandelbrot_boundary_points"])
    g["vario_boundary_points_N160"] = vb["mandelbrot_boundary_points"](N=160, dist_thresh=0.01, max_iter=200)

    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.1f} KiB, {len(g)} arrays)")


if __name__ == "__main__":
    main()

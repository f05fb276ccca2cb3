"""CPU: the oracle (oracle/) against outputs of the reference's own Python functions.

tests/golden/reference_vectors.npz was produced by oracle/gen_golden.py, which executes the
function definitions of the scripts under /root/reference.  These tests pin the oracle; the
GPU tests then compare the CUDA path with the oracle and with the same fixtures.
"""
import numpy as np
import pytest

from conftest import match_sorted_complex

FAMILIES = ["lucas_all_ones", "pell_like_all_twos", "sparser_gap_1_0_1_then_ones", "padovan_like_0_1_then_ones"]


@pytest.mark.parametrize("tag", ["cfg1", "seahorse", "tip"])
def test_dwell_grid_bit_exact(oracle, golden, tag):
    xlim0, xlim1, ylim0, ylim1, res, mi = golden[f"dwell_{tag}_args"]
    xs = np.linspace(xlim0, xlim1, int(res)); ys = np.linspace(ylim0, ylim1, int(res))
    assert np.array_equal(xs, golden[f"dwell_{tag}_xs"]) and np.array_equal(ys, golden[f"dwell_{tag}_ys"])
    d, work = oracle.dwell_grid(xs, ys, int(mi))
    Z = golden[f"dwell_{tag}_Z"]
    assert np.array_equal(d, Z)
    assert work == int(np.minimum(Z.astype(np.int64) + 1, int(mi)).sum())


@pytest.mark.parametrize("tag", ["cfg1", "seahorse", "tip"])
def test_python_restatement_rows(golden, tag):
    """oracle/pyref.py (the pure-Python loop bench.py times as the reference-speed CPU baseline) on sample rows of the
    grids produced by the reference's own compute_grid."""
    from oracle import pyref
    xs, ys, Z = golden[f"dwell_{tag}_xs"], golden[f"dwell_{tag}_ys"], golden[f"dwell_{tag}_Z"]
    mi = int(golden[f"dwell_{tag}_args"][5])
    rows = [0, len(ys) // 3, len(ys) // 2, len(ys) - 1]
    assert np.array_equal(pyref.dwell_rows(xs, ys[rows], mi), Z[rows])


def test_dwell_points(oracle, golden):
    mi = int(golden["dwell_points_mi"][0])
    for (x, y), want in zip(golden["dwell_points_xy"], golden["dwell_points_out"]):
        d, _ = oracle.dwell_grid([x], [y], mi)
        assert d[0, 0] == want
    # analytic facts: dwell(0)=max_iter, |c|>2 escapes at n=0
    assert oracle.dwell_grid([0.0], [0.0], 123)[0][0, 0] == 123
    assert oracle.dwell_grid([2.5], [0.0], 123)[0][0, 0] == 0


def test_batch_potential(oracle, golden):
    c = golden["potential_points_c"]
    g, it, phi = oracle.batch_potential(c, 1500, 2.0)
    assert np.array_equal(it, golden["potential_points_it"])
    np.testing.assert_allclose(g, golden["potential_points_g"], rtol=1e-14, atol=0)
    want = golden["potential_points_phi"]
    assert np.array_equal(np.isnan(phi.real), np.isnan(want.real))
    m = ~np.isnan(want.real)
    np.testing.assert_allclose(phi[m], want[m], rtol=1e-14)


def test_grid_potentials(oracle, golden):
    gx, gy = golden["potgrid_x"], golden["potgrid_y"]
    _, f = oracle.potential_grid(gx, gy, 60, 10.0, oracle.FIELD_POW2_ALWAYS)      # Potentials.py:32-47
    np.testing.assert_allclose(f, golden["potentials_escape_R10_mi60"], rtol=1e-14, atol=0)
    assert (golden["potentials_escape_R10_mi60"] < 0).any()     # bounded orbits with |z|<1 give negative values
    _, f = oracle.potential_grid(gx, gy, 80, 2.0, oracle.FIELD_INV_K)             # Laplacian_C-M.py:27-43
    np.testing.assert_allclose(f, golden["laplacian_cm_potential_R2_mi80"], rtol=1e-14, atol=0)
    _, f = oracle.potential_grid(gx, gy, 70, 10.0, oracle.FIELD_INV_K)            # Iterative_...py:114-130
    np.testing.assert_allclose(f, golden["iterative_escape_R10_mi70"], rtol=1e-14, atol=0)


def test_vario_escape_potential(oracle, golden):
    """variograms_construct_mandelbrot.py:148-173.  The reference iterates numpy complex ARRAYS, whose
    multiply is FMA-fused on AVX hosts, so a few pixels may escape one step apart;
    everything else must agree to rounding."""
    gx, gy = golden["vario_grid_x"], golden["vario_grid_y"]
    _, raw = oracle.potential_grid(gx, gy, 90, 4.0, oracle.FIELD_POW2_FIRST)
    got = oracle.smooth5(raw)
    want = golden["vario_escape_potential_mi90"]
    close = np.isclose(got, want, rtol=1e-12, atol=1e-300)
    assert close.mean() > 0.99
    assert ((got == 0) == (want == 0)).mean() > 0.99


def test_distance_estimators(oracle, golden):
    d, _ = oracle.distance_grid(golden["de_scalar_x"], golden["de_scalar_y"], 200, 1e6, 1e-16, 0)
    np.testing.assert_allclose(d, golden["de_scalar_dist"], rtol=1e-13, atol=0)
    d, esc = oracle.distance_grid(golden["vario_grid_x"], golden["vario_grid_y"], 90, 4.0, 1e-14, 1)
    want_esc = golden["vario_de_escaped"]
    assert (esc == want_esc).mean() > 0.99           # numpy array arithmetic is FMA-contaminated, see above
    both = esc & want_esc
    close = np.isclose(d[both], golden["vario_de_dist"][both], rtol=1e-9, atol=0)
    assert close.mean() > 0.98


def test_distance_estimator_tracker_module(oracle, golden):
    """tci_construct_mandelbrot_v002_fixed.py:35-47 (the module gi_assumption_tracker_v3.py loads): z of the
    first escape, dz read after the loop.  dz normally overflows (d = 0); it stays finite only for points
    that escape in the last iterations, and there the FMA-fused numpy array recurrence has
    been amplified by ~240 chaotic steps, hence the loose tolerance on those few values."""
    for tag in ("tci_fixed", "tci_fixed_zoom"):
        d, esc = oracle.distance_grid(golden[tag + "_x"], golden[tag + "_y"], 250, 250.0, 1e-12, 2)
        want, want_esc = golden[tag + "_dist"], golden[tag + "_escaped"]
        assert (esc == want_esc).mean() > 0.999
        assert np.array_equal(d != 0, want != 0)
        m = want != 0
        np.testing.assert_allclose(d[m], want[m], rtol=1e-4)
    assert (golden["tci_fixed_zoom_dist"] != 0).sum() >= 10
    # variant 3 restates numpy's SIMD complex multiply (fma(ar, br, -(ai*bi)), fma(ar, bi, ai*br)): the same fixtures to
    # rounding (log / hypot of numpy vs libm), masks exact, and the stock module's boundary sample as an exact set
    for tag in ("tci_fixed", "tci_fixed_zoom"):
        d, esc = oracle.distance_grid(golden[tag + "_x"], golden[tag + "_y"], 250, 250.0, 1e-12, 3)
        want, want_esc = golden[tag + "_dist"], golden[tag + "_escaped"]
        assert np.array_equal(esc, want_esc) and np.array_equal(d != 0, want != 0)
        np.testing.assert_allclose(d, want, rtol=2e-15, atol=0)
    xs = np.linspace(-2.25, 1.25, 150); ys = np.linspace(-1.75, 1.75, 150)
    d, esc = oracle.distance_grid(xs, ys, 250, 250.0, 1e-12, 3)
    jj, ii = np.nonzero(esc & (d <= np.quantile(d[esc], 0.25)))
    assert np.array_equal(xs[ii] + 1j * ys[jj], golden["tci_fixed_boundary_sample_grid150"])


def test_stencils_bit_exact(oracle, golden):
    U = golden["stencil_U"]; h = float(golden["stencil_h"][0])
    assert np.array_equal(oracle.laplacian(U, h), golden["stencil_laplacian"])
    assert np.array_equal(oracle.laplacian(U, h), golden["stencil_laplacian_fd"])
    assert np.array_equal(oracle.smooth5(U), golden["stencil_smooth5"])


def test_log_potentials(oracle, golden):
    P = golden["logpot_points"]; gx, gy = golden["potgrid_x"], golden["potgrid_y"]
    np.testing.assert_allclose(oracle.log_potential(P, gx, gy, 1e-12, 0), golden["logpot_potentials"], rtol=1e-12)
    np.testing.assert_allclose(oracle.log_potential(P, gx, gy, 1e-12, 1), golden["logpot_laplacian_cm"], rtol=1e-12)
    np.testing.assert_allclose(oracle.log_potential(P, gx, gy, 1e-12, 2), golden["logpot_iterative"], rtol=1e-12)
    np.testing.assert_allclose(oracle.log_potential(P, golden["vario_grid_x"], golden["vario_grid_y"], 1e-6, 3),
                               golden["logpot_vario_eps1e-6"], rtol=1e-12)


@pytest.mark.parametrize("fam", FAMILIES)
def test_families(oracle, golden, fam):
    vals, counts = golden[f"family_{fam}_values"], golden[f"family_{fam}_counts"]
    off = 0
    for n, cnt in zip(range(2, 26), counts):
        got = oracle.inverse_eigenvalues_toprow(oracle.family_toprow(fam, n), 1e-12)
        assert len(got) == cnt
        assert match_sorted_complex(got, vals[off:off + cnt]) < 1e-12
        off += cnt


def test_known_answers(oracle, golden):
    # n = 2 Lucas: roots of x^2 - x - 1 -> loci {1/phi, -phi}   
    phi = (1 + 5 ** 0.5) / 2
    got = np.sort(oracle.inverse_eigenvalues_toprow([1.0, 1.0]).real)
    np.testing.assert_allclose(got, [-phi, 1 / phi], rtol=1e-15)
    # shipped artefact: 2400 points for n = 20..300 step 20 (v3_T25_sigma3_dense.csv:2)
    assert int(golden["tci_construct_points_count"][0]) == 2400
    assert sum(range(20, 301, 20)) == 2400


def test_green_function_sums(oracle, golden):
    """The oracle's Nystrom sums against g_real / dPhi of the reference's own RiemannMapDisk_GreenModulus
    (lucas_to_cardioid_v40_reference.py:184-257) on a synthetic boundary."""
    bdy, ds, sig, z = golden["green_bdy"], golden["green_ds"], golden["green_sigma"], golden["green_targets"]
    ar, ai, Cc, sh = golden["green_params"]
    a = complex(ar, ai)
    got = -np.log(np.abs(z - a) + 1e-300) + oracle.weighted_log_sum(z, bdy, sig * ds) + Cc + sh
    np.testing.assert_allclose(got, golden["green_g_real"], rtol=1e-12, atol=1e-13)
    DZ0 = np.where(np.abs(z - a) < 1e-14, 1e-14 + 0j, z - a)
    got = -1.0 / DZ0 + oracle.weighted_cauchy_sum(z, bdy, sig * ds, 1e-14)
    np.testing.assert_allclose(got, golden["green_dPhi"], rtol=1e-12, atol=1e-12 * np.abs(golden["green_dPhi"]).max())
    assert np.isfinite(golden["green_g_real"][:-1]).all()


def test_curvature_localpoly(oracle, golden):
    """The oracle's restatement against compute_curvature_localpoly of the reference (boundary_curvature_localpoly.py)."""
    from helpers import assert_columns_close
    got = oracle.curvature_localpoly(golden["curv_closed_P"], 7, True)
    assert_columns_close(got, golden["curv_closed_out"], 1e-9, 1e-9)
    got = oracle.curvature_localpoly(golden["curv_open_P"], 4, False)
    want = golden["curv_open_out"]                         # stride 3: every third point is a fit, the rest interpolated
    assert_columns_close(got[::3], want[::3], 1e-9, 1e-9)


def _check_vario(out, ref):
    centers, gamma, counts = out
    assert np.array_equal(centers, ref[0])
    assert np.array_equal(counts, ref[2].astype(int))              # pair counts per bin: exact
    assert np.array_equal(np.isnan(gamma), np.isnan(ref[1]))
    np.testing.assert_allclose(gamma, ref[1], rtol=1e-12)          # the reference's np.mean is a pairwise sum


def test_pair_statistics(oracle, golden):
    """the pair-histogram restatement against the reference's variogram / pair-correlation functions
    (Variogram-Mandelbrot-Construct.py:106-152, Iterative_Variogram_Laplacian.py:53-86, spatial_stats_phase2.py:9-47)."""
    pc, pv = golden["pair_cloud"], golden["pair_values"]
    _check_vario(oracle.empirical_variogram_field(pc, pv, 60), golden["pair_vario_field"])
    _check_vario(oracle.empirical_variogram_field(pc, pv, 17, 0.9), golden["pair_vario_field_maxd"])
    _check_vario(oracle.empirical_variogram_coords(pc, 60), golden["pair_vario_coords"])
    _check_vario(oracle.empirical_variogram_field(pc, pv, 50), golden["pair_vario_iter_values"])
    _check_vario(oracle.empirical_variogram_coords(pc, 50), golden["pair_vario_iter_coords"])
    for pts, rmax, dr, kc, kk in [(pc, 1.5, 0.01, "pair_correlation_r1.5_dr0.01", "pair_ripley_r1.5_dr0.01"),
                                  (golden["pair_lattice"], 8.0, 0.5, "pair_lattice_correlation", "pair_lattice_ripley")]:
        r, g_r = oracle.pair_correlation(pts, rmax, dr)
        assert np.array_equal(r, golden[kc][0]) and np.array_equal(g_r, golden[kc][1])      # integer counts / same norm
        r, K = oracle.ripley_K(pts, rmax, dr)
        assert np.array_equal(r, golden[kk][0]) and np.array_equal(K, golden[kk][1])
    # distances exactly on the integer bin edges (3-4-5 triangles): [lo, hi) semantics of np.digitize
    _check_vario(oracle.empirical_variogram_field(golden["pair_lattice"], golden["pair_lattice_values"], 16, 16.0),
                 golden["pair_lattice_vario"])


def test_tracker_density_stage(oracle, golden):
    """mollified_histogram / tv / overlap / KL / GI flows restated (gi_assumption_tracker_v3.py:91-151,
    tci_construct_mandelbrot_v002_fixed.py:84-86) against the reference's own functions."""
    dom = tuple(golden["density_domain_eps"][:4]); eps = float(golden["density_domain_eps"][4])
    Cc, Mb = golden["density_cloud_C"], golden["tci_fixed_boundary_sample_grid150"]
    for bins, sig in [(64, 1.0), (64, 0.0), (50, 2.5)]:
        tag = f"b{bins}_s{sig}"
        P_C = oracle.mollified_histogram(dom, eps, Cc, bins, sig)
        P_M = oracle.mollified_histogram(dom, eps, Mb, bins, sig)
        assert np.array_equal(P_C, golden[f"density_PC_{tag}"]) and np.array_equal(P_M, golden[f"density_PM_{tag}"])
        sc = golden[f"density_scalars_{tag}"]
        assert oracle.tv_distance(P_C, P_M) == sc[0] and oracle.overlap_mass(P_C, P_M) == sc[1]
        assert oracle.KL(P_M, P_C, eps) == sc[2]
        X, T, kl0, klT = oracle.gi_flow(P_M, P_C, 0.1, 800, 5, 1e-6, eps)
        assert T == int(sc[3]) and kl0 == sc[4] and klT == sc[5] and np.array_equal(X, golden[f"density_flow_XT_{tag}"])
        X, T, kl0, klT = oracle.gi_flow(P_M, P_C, 0.1, 25, eps=eps)
        assert T == int(sc[6]) and kl0 == sc[7] and klT == sc[8] and np.array_equal(X, golden[f"density_flow_XF_{tag}"])


def test_alpha_shape_edges(oracle, golden):
    """circumradius / alpha_shape_edges restated (construct_boundary_alpha.py:45-82) on the recorded Delaunay
    simplices, against the reference's own outputs."""
    P, S = golden["alpha_points"], golden["alpha_simplices"]
    for tag, alpha in (("a6", 6.0), ("a12", 12.0)):
        keep, radius, edges = oracle.alpha_shape_edges(P, S, alpha)
        assert np.array_equal(radius, golden["alpha_radius"])
        assert np.array_equal(np.asarray(edges, dtype=np.int32).reshape(-1, 2), golden[f"alpha_{tag}_edges"])


@pytest.mark.parametrize("tag,family", [("lucas", None), ("pell", "pell_like_all_twos")])
def test_per_n_stats_composition(oracle, golden, tag, family):
    """The oracle's eigvals + batch_potential composed like per_n_stats / cumulative_stats
    (lucas_equipotential_test_v3.py:294-327) reproduces the rows of the reference's own functions."""
    per, cum, acc = [], [], []
    for n in range(2, 31):
        top = np.ones(n) if family is None else oracle.family_toprow(family, n)
        inv = oracle.inverse_eigenvalues_toprow(top, 1e-12)
        acc.append(inv)
        for rows, pts in ((per, inv), (cum, np.concatenate(acc))):
            g, _, _ = oracle.batch_potential(pts, 3000, 2.0)
            o = g > 0
            rows.append([len(g), o.sum(), o.mean(), np.median(g[o]), np.mean(g[o]), np.std(g[o]), np.quantile(g[o], 0.1), np.quantile(g[o], 0.9)])
    np.testing.assert_allclose(np.array(per, dtype=float), golden[f"per_n_stats_{tag}_2_30_mi3000"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(np.array(cum, dtype=float), golden[f"cumulative_stats_{tag}_2_30_mi3000"], rtol=1e-12, atol=0)

"""GPU: binned pair statistics -- counts bit-exact, weight sums to 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ps(gpu):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import pairstats
    return pairstats


def _check_vario(out, ref):
    centers, gamma, counts = out
    assert np.array_equal(centers, ref[0])
    assert np.array_equal(counts, ref[2].astype(int))
    assert np.array_equal(np.isnan(gamma), np.isnan(ref[1]))
    np.testing.assert_allclose(gamma, ref[1], rtol=1e-12)


def test_golden_reference_functions(ps, golden):
    pc, pv = golden["pair_cloud"], golden["pair_values"]
    _check_vario(ps.empirical_variogram_field(pc, pv, nbins=60), golden["pair_vario_field"])
    _check_vario(ps.empirical_variogram_field(pc, pv, nbins=17, max_dist=0.9), golden["pair_vario_field_maxd"])
    _check_vario(ps.empirical_variogram_coords(pc, nbins=60), golden["pair_vario_coords"])
    _check_vario(ps.empirical_variogram_from_field_locs(pc, values=pv, nbins=50), golden["pair_vario_iter_values"])
    _check_vario(ps.empirical_variogram_from_field_locs(pc, values=None, nbins=50), golden["pair_vario_iter_coords"])
    _check_vario(ps.empirical_variogram_field(golden["pair_lattice"], golden["pair_lattice_values"], nbins=16, max_dist=16.0),
                 golden["pair_lattice_vario"])
    for pts, rmax, dr, kc, kk in [(pc, 1.5, 0.01, "pair_correlation_r1.5_dr0.01", "pair_ripley_r1.5_dr0.01"),
                                  (golden["pair_lattice"], 8.0, 0.5, "pair_lattice_correlation", "pair_lattice_ripley")]:
        r, g_r = ps.pair_correlation(pts, rmax, dr)
        assert np.array_equal(r, golden[kc][0]) and np.array_equal(g_r, golden[kc][1])
        r, K = ps.ripley_K(pts, rmax, dr)
        assert np.array_equal(r, golden[kk][0]) and np.array_equal(K, golden[kk][1])


@pytest.mark.parametrize("n", [2, 3, 31, 127, 128, 129, 255, 257, 1000, 4100])
@pytest.mark.parametrize("weight", ["none", "value", "dist2"])
def test_histogram_against_oracle(ps, oracle, n, weight):
    rng = np.random.default_rng(n * 3 + len(weight))
    P = np.c_[rng.uniform(-2.2, 1.2, n), rng.uniform(-1.6, 1.6, n)]
    P[n // 2] = P[0]                                           # a coincident pair: d = 0 lands in the first bin
    v = rng.standard_normal(n)
    bins = np.linspace(0.0, 2.0, 61)
    c_ref, s_ref, dmax_ref = oracle.pair_histogram(P, bins[:-1], bins[1:], v, weight)
    c, s = ps.pair_histogram(P, bins[:-1], bins[1:], v, weight)
    assert np.array_equal(c, c_ref)
    if weight != "none":
        np.testing.assert_allclose(s, s_ref, rtol=1e-12, atol=0)
    assert ps.max_pair_distance(P) == dmax_ref
    assert ps.last_stats["work_units"] == n * (n - 1) // 2


def test_edges_and_overlapping_shells(ps, oracle):
    """Integer lattice against integer / half-integer edges (many distances exactly on an edge), shells whose upper
    edge overlaps the next shell, a last bin open to +inf, more bins than one warp-private histogram holds."""
    lat = np.array([[i, j] for i in range(40) for j in range(33)], dtype=np.float64)
    v = (lat[:, 0] * 5 + lat[:, 1] * 3) % 11
    for lo, hi in [(np.arange(0.0, 50.0, 1.0), np.arange(1.0, 51.0, 1.0)),
                   (np.arange(0.0, 50.0, 0.5), np.arange(0.0, 50.0, 0.5) + 0.5000000000000001),
                   (np.arange(0.0, 30.0, 1.0), np.arange(0.0, 30.0, 1.0) + 1.75),            # overlaps the next shell
                   (np.arange(0.0, 20.0, 1.0), np.append(np.arange(1.0, 20.0, 1.0), np.inf)),
                   (np.linspace(0.0, 52.0, 1500)[:-1], np.linspace(0.0, 52.0, 1500)[1:]),    # 1499 bins: shared histogram
                   (np.array([3.0]), np.array([5.0])),                                        # one bin, d = 5 excluded
                   (np.array([0.0, 1.0, 1.5, 4.0, 4.25, 30.0]), np.array([1.0, 1.5, 4.0, 4.25, 30.0, 31.0]))]:  # uneven
        c_ref, s_ref, _ = oracle.pair_histogram(lat, lo, hi, v, "value")
        c, s = ps.pair_histogram(lat, lo, hi, v, "value")
        assert np.array_equal(c, c_ref)
        np.testing.assert_allclose(s, s_ref, rtol=1e-12)
    n = len(lat)
    c, _ = ps.pair_histogram(lat, [0.0], [np.inf])
    assert int(c[0]) == n * (n - 1) // 2
    # 3000 bins: the host splits the edge arrays over two passes
    e = np.linspace(0.0, 52.0, 3001)
    c, s = ps.pair_histogram(lat, e[:-1], e[1:], v, "value")
    c_ref, s_ref, _ = oracle.pair_histogram(lat, e[:-1], e[1:], v, "value")
    assert np.array_equal(c, c_ref)


def test_degenerate_inputs(ps):
    assert all(a.size == 0 for a in ps.empirical_variogram_field(np.zeros((1, 2)), np.zeros(1)))
    c, s = ps.pair_histogram(np.zeros((1, 2)), [0.0, 1.0], [1.0, 2.0])
    assert not c.any() and not s.any()
    c, s = ps.pair_histogram(np.zeros((0, 2)), [0.0], [1.0])
    assert not c.any()
    assert ps.max_pair_distance(np.zeros((1, 2))) == 0.0
    with pytest.raises(ValueError):
        ps.pair_histogram(np.zeros((4, 2)), [0.0, 0.0], [1.0, 1.0])          # lower edges must increase
    with pytest.raises(ValueError):
        ps.pair_histogram(np.array([[0.0, np.nan], [1.0, 1.0]]), [0.0], [1.0])
    with pytest.raises(ValueError):
        ps.pair_histogram(np.zeros((4, 2)), [0.0], [1.0], np.zeros(3), "value")
    # all points coincide: every pair has d = 0 -> first bin; D.max() = 0 -> bins collapse like the reference's linspace(0, 0)
    P = np.ones((50, 2))
    c, _ = ps.pair_histogram(P, [0.0, 1.0], [1.0, 2.0])
    assert c.tolist() == [50 * 49 // 2, 0]


def test_tracker_cloud_full_size(ps):
    """The 37 820-point construct cloud size of the tracker's last level (v3_T25_sigma3_dense.csv:5): 7.15e8 pairs.
    Size-independent properties: every pair lands in exactly one bin of a partition of [0, inf); the cumulated
    ripley counts are monotone; a point-order permutation leaves counts unchanged; dist2 sums match counts * d^2
    bounds bin by bin."""
    n = 37820
    rng = np.random.default_rng(2)
    th = rng.uniform(0, 2 * np.pi, n)
    P = np.c_[-0.5 + 1.2 * np.cos(th) * (1 - 0.5 * np.cos(th)), 1.2 * np.sin(th) * (1 - 0.5 * np.cos(th))]
    P += 0.02 * rng.standard_normal(P.shape)
    e = np.append(np.linspace(0.0, 1.5, 61), np.inf)
    c, s = ps.pair_histogram(P, e[:-1], e[1:], None, "dist2")
    assert int(c.sum()) == n * (n - 1) // 2
    assert ps.last_stats["work_units"] == n * (n - 1) // 2
    inner = slice(0, 60)
    assert np.all(s[inner] >= c[inner] * e[:-1][inner] ** 2 * (1 - 1e-12))
    assert np.all(s[inner] <= c[inner] * e[1:][inner] ** 2 * (1 + 1e-12))
    perm = rng.permutation(n)
    c2, s2 = ps.pair_histogram(P[perm], e[:-1], e[1:], None, "dist2")
    assert np.array_equal(c, c2)
    np.testing.assert_allclose(s2, s, rtol=1e-11)
    dmax = ps.max_pair_distance(P)
    assert dmax == ps.max_pair_distance(P[perm])
    # exact check of one row block against numpy: pairs (i, j>i) for the first 64 points
    sub = 64
    dx = P[:sub, None, 0] - P[None, :, 0]; dy = P[:sub, None, 1] - P[None, :, 1]
    D = np.sqrt(dx * dx + dy * dy)
    assert D.max() <= dmax


@pytest.mark.parametrize("tag", ["a", "b"])
def test_subsampled_semivariograms_golden(gpu, golden, tag):
    """sample_semivariogram / sample_cross_semivariogram (variograms_construct_mandelbrot.py:178-315) with the seeded
    global stream the reference was run with: per-block bin counts / sums from lm_pair_histogram, the ordered pair list
    of the block that crosses a bin's cap from lm_pair_select_sqdiff, the reference's own np.random.choice draws on the
    host.  gamma to 1e-12 against the reference's output (counts are exact, or the subsets would differ)."""
    import types
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import pairstats
    x0, x1, y0, y1, nx, ny = golden["semivario_grid_args"]
    X, Y = np.meshgrid(np.linspace(x0, x1, int(nx)), np.linspace(y0, y1, int(ny)), indexing="xy")
    grid = types.SimpleNamespace(X=X, Y=Y)
    bins = golden[f"semivario_{tag}_bins"]; cap = int(golden[f"semivario_{tag}_cap"][0])
    np.random.seed(777)
    rc, gam = pairstats.sample_semivariogram(golden["semivario_field1"], grid, bins, max_pairs_per_bin=cap)
    assert np.array_equal(rc, golden[f"semivario_{tag}_centers"])
    np.testing.assert_allclose(gam, golden[f"semivario_{tag}_gamma"], rtol=1e-12, atol=0)
    np.random.seed(778)
    rc, gam = pairstats.sample_cross_semivariogram(golden["semivario_field1"], golden["semivario_field2"], grid, bins,
                                                   max_pairs_per_bin=cap)
    np.testing.assert_allclose(gam, golden[f"semivario_{tag}_cross_gamma"], rtol=1e-12, atol=0)


def test_pair_select_matches_numpy(gpu):
    """lm_pair_select_sqdiff: row-major order, diagonal exclusion, edge-inclusive lower bound (3-4-5 lattice)."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import pairstats
    rng = np.random.default_rng(3)
    xa = rng.integers(0, 12, 300).astype(float); ya = rng.integers(0, 12, 300).astype(float); va = rng.standard_normal(300)
    xb = rng.integers(0, 12, 257).astype(float); yb = rng.integers(0, 12, 257).astype(float); vb = rng.standard_normal(257)
    for lo, hi in ((5.0, 6.0), (0.0, 1.0), (4.999999, 5.0), (13.0, 1e9)):
        D = np.sqrt((xa[:, None] - xb[None, :]) ** 2 + (ya[:, None] - yb[None, :]) ** 2)
        m = (D >= lo) & (D < hi)
        want = ((va[:, None] - vb[None, :]) ** 2)[np.where(m)]
        got = pairstats._select_sqdiff((xa, ya, va), (xb, yb, vb), lo, hi, False, int(m.sum()))
        assert np.array_equal(got, want)
    D = np.sqrt((xa[:, None] - xa[None, :]) ** 2 + (ya[:, None] - ya[None, :]) ** 2)
    m = (D >= 0.0) & (D < 2.0) & ~np.eye(300, dtype=bool)
    got = pairstats._select_sqdiff((xa, ya, va), (xa, ya, va), 0.0, 2.0, True, int(m.sum()))
    assert np.array_equal(got, ((va[:, None] - va[None, :]) ** 2)[np.where(m)])

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

"""The GPU-backed module for `gi_assumption_tracker_v3.py --module`: loaded by file path
exactly as the tracker does (gi_assumption_tracker_v3.py:84-90), attributes overwritten per level (:194,
208-209), generators compared with the stock module's outputs (tests/golden, from
tci_construct_mandelbrot_v002_fixed.py) and with the shipped artefact counts (v3_T25_sigma3_dense.csv:2-5)."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

MODULE = (Path(__file__).resolve().parents[1] / "inverse_eigenvalue_loci_mandelbrot_correspondence_b200"
          / "tci_construct_mandelbrot_b200.py")


def load_module(path, name="tci_fixed_import"):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_module_contract_cpu(shim):
    mod = load_module(MODULE)
    for attr in ("domain", "eps", "mandelbrot_grid", "mandelbrot_samples", "max_iter", "escape_R", "grid_bins"):
        assert hasattr(mod, attr)
    for fn in ("construct_points", "sample_mandelbrot_boundary", "entropic_ot_alignment", "procrustes_align_no_scale",
               "KL", "to_prob", "tci_flow", "mandelbrot_distance_estimator", "lucas_companion"):
        assert callable(getattr(mod, fn))
    assert mod.domain == (-2.25, 1.25, -1.75, 1.75) and mod.eps == 1e-12 and mod.mandelbrot_grid == 600
    # host helpers: rigid alignment, histogram probabilities, KL
    rng = np.random.default_rng(3)
    Y = rng.standard_normal(50) + 1j * rng.standard_normal(50)
    X = Y[rng.permutation(50)] + 1e-3
    back = mod.procrustes_align_no_scale(Y + (2 - 1j), Y)                # pure translation is undone
    assert np.allclose(back, Y, atol=1e-12)
    moved = mod.procrustes_align_no_scale(Y * np.exp(0.3j) + (2 - 1j), Y)  # always a rigid motion onto Y's centroid
    assert abs(moved.mean() - Y.mean()) < 1e-12
    assert np.allclose(np.abs(moved[:, None] - moved[None, :]), np.abs(Y[:, None] - Y[None, :]), atol=1e-12)
    assert mod.KL.lm_eps() == mod.eps                                    # what tracker.gi_flow_* read
    if shim.device_count() < 1:          # no CPU fallback behind the module either
        with pytest.raises(RuntimeError):
            mod.construct_points([3])
        with pytest.raises(RuntimeError):
            mod.entropic_ot_alignment(X, Y)
        with pytest.raises(RuntimeError):
            mod.to_prob(Y, 16)
        with pytest.raises(RuntimeError):
            mod.KL(np.full((4, 4), 1 / 16), np.full((4, 4), 1 / 16))


def test_nearest_match_oracle_is_the_reference_rule(oracle):
    """The oracle's nearest_match against the stock module's own expression
    argmax(exp(-cdist/mean/eps), axis=1) (tci_construct_mandelbrot_v002_fixed.py:66-70), restated with numpy."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal(300) + 1j * rng.standard_normal(300)
    Y = np.concatenate([rng.standard_normal(200) + 1j * rng.standard_normal(200), X[:40], X[:40]])   # exact ties
    M = np.sqrt((X.real[:, None] - Y.real[None, :]) ** 2 + (X.imag[:, None] - Y.imag[None, :]) ** 2)
    K = np.nan_to_num(np.exp(-(M / M.mean()) / 0.8))
    idx, dist = oracle.nearest_match(X, Y)
    assert np.array_equal(idx, np.argmax(K, axis=1))
    assert np.array_equal(dist, M[np.arange(300), idx])


def test_exp_rule_near_ties_pick_the_first_index(oracle):
    """Known gap: exp() maps scaled distances that differ by an ulp to the same double, so the stock
    module's argmax(exp(-M/mean/eps)) returns the FIRST of two near-equidistant points, while nearest_match returns the
    true minimum.  The four pairs below are the ones where the two rules differ in the reference's own level-4 run
    (real-axis cloud points against the mirror pair of boundary samples -1.2483 +- 0.0509i, captured from that run)."""
    X = np.array([-1.0183764595711713 - 0j, -1.0110099247483106 - 0j, -1.0078593635360475 - 0j, -1.0068756134485202 - 0j])
    pair = np.array([-1.248298572996707 + 0.05093304061470927j, -1.248298572996707 - 0.050933040614709046j])   # first: 1 ulp farther
    Y = np.concatenate([pair, 2.0 * np.array([0.9 + 0.7j, -1.2 - 1.3j, 0.1 + 1.1j, 1.0 - 0.9j])])
    M = np.sqrt((X.real[:, None] - Y.real[None, :]) ** 2 + (X.imag[:, None] - Y.imag[None, :]) ** 2)
    assert np.all(M[:, 1] < M[:, 0]) and np.all(M[:, 0] - M[:, 1] < 1e-16)
    K = np.nan_to_num(np.exp(-(M / M.mean()) / 0.8))
    assert np.sum(np.argmax(K, axis=1) == 0) >= 3            # the reference's rule: first index of the tied maximum
    assert np.array_equal(oracle.nearest_match(X, Y)[0], [1, 1, 1, 1])      # true nearest


@pytest.mark.gpu
def test_tracker_levels(gpu, golden):
    mod = load_module(MODULE)
    mod.domain = (-2.25, 1.25, -1.75, 1.75)
    # n_construct_pts of the shipped tracker runs: 2400 / 6000 / 14820 / 37820 for construct_max_n 300/480/760/1220
    for nmax, want in ((300, 2400), (480, 6000), (760, 14820), (1220, 37820)):
        pts = mod.construct_points(list(range(20, nmax + 1, 20)))
        assert pts.size == want and np.isfinite(pts).all()
    # sample_mandelbrot_boundary at a level the stock module was run at for the fixture
    mod.mandelbrot_grid = 150; mod.mandelbrot_samples = 10 ** 9
    got = mod.sample_mandelbrot_boundary()
    want = golden["tci_fixed_boundary_sample_grid150"]
    common = np.intersect1d(got, want).size
    assert common >= 0.999 * max(got.size, want.size)
    if got.size == want.size:
        assert (got == want).mean() > 0.99             # same row-major order
    # seeded subsampling draws through numpy's global stream like the stock module
    mod.mandelbrot_samples = 500
    np.random.seed(7); a = mod.sample_mandelbrot_boundary()
    np.random.seed(7); b = mod.sample_mandelbrot_boundary()
    assert a.size == 500 and np.array_equal(a, b)
    # the matching rule on the GPU: first index of the nearest point, bit-identical to the oracle (exact ties included)
    rng = np.random.default_rng(9)
    X = rng.standard_normal(5000) + 1j * rng.standard_normal(5000)
    Y = np.concatenate([rng.standard_normal(3000) + 1j * rng.standard_normal(3000), X[:500], X[:500]])
    from oracle import oracle as orc
    gi, gd = gpu.potentials.nearest_match(X, Y)
    oi, od = orc.nearest_match(X, Y)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    Ym, Xs = mod.entropic_ot_alignment(X[:2000], Y[:2000])          # equal sizes: no subsampling
    assert np.array_equal(Ym, Y[:2000][orc.nearest_match(X[:2000], Y[:2000])[0]]) and np.array_equal(Xs, X[:2000])
    # one tracker level end to end
    mod.mandelbrot_grid = 600; mod.mandelbrot_samples = 25000
    np.random.seed(7)
    Cp = mod.construct_points(list(range(20, 301, 20))); M = mod.sample_mandelbrot_boundary()
    Mmatch, Csub = mod.entropic_ot_alignment(Cp, M)
    Cal = mod.procrustes_align_no_scale(Csub, Mmatch)
    assert Cal.size == 2400 and Mmatch.size == 2400
    assert np.isfinite(mod.KL(mod.to_prob(Mmatch, 64), mod.to_prob(Cal, 64)))
    # histogram probabilities, KL and the module's own flow (device KL)
    rng = np.random.default_rng(3)
    Y = rng.standard_normal(50) + 1j * rng.standard_normal(50)
    X = Y[rng.permutation(50)] + 1e-3
    P = mod.to_prob(Y, 16); Q = mod.to_prob(X, 16)
    assert P.shape == (16, 16) and abs(P.sum() - 1) < 1e-12
    H = np.maximum(np.histogram2d(Y.real, Y.imag, bins=(16, 16), range=[[-2.25, 1.25], [-1.75, 1.75]])[0], mod.eps)
    assert np.array_equal(P, H / H.sum())
    assert mod.KL(P, P) == 0.0 and mod.KL(P, Q) > 0
    kls, traj = mod.tci_flow(P, Q)
    assert len(kls) == mod.T + 1 and kls[-1] < kls[0] * 1e-4 and np.all(np.diff(kls) < 0)


def test_tracker_refuses_a_module_with_another_KL():
    """The device GI flow evaluates the stock KL; a plug-in whose KL differs must not be replaced silently."""
    import types
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import gi_assumption_tracker_v3 as trk
    stock = types.SimpleNamespace(eps=1e-12)
    stock.KL = lambda P, X: float(np.sum(np.clip(P, 1e-12, None) * (np.log(np.clip(P, 1e-12, None)) - np.log(np.clip(X, 1e-12, None)))))
    assert trk._is_stock_KL(stock)
    other_eps = types.SimpleNamespace(eps=1e-12)
    other_eps.KL = lambda P, X: float(np.sum(np.clip(P, 1e-9, None) * (np.log(np.clip(P, 1e-9, None)) - np.log(np.clip(X, 1e-9, None)))))
    assert not trk._is_stock_KL(other_eps)
    sym = types.SimpleNamespace(eps=1e-12)
    sym.KL = lambda P, X: 0.5 * (stock.KL(P, X) + stock.KL(X, P))
    assert not trk._is_stock_KL(sym)
    base2 = types.SimpleNamespace(eps=1e-12)
    base2.KL = lambda P, X: stock.KL(P, X) / np.log(2.0)
    assert not trk._is_stock_KL(base2)
    with pytest.raises(SystemExit):
        trk._device_ops(sym)

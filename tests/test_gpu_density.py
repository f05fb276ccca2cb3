"""GPU: the tracker's density stage -- histogram, blur, numpy-order sums, divergences, GI flow.
Everything but the logarithm is bit-exact against numpy / scipy; KL values to 1e-13 absolute."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KL_ATOL = 1e-13       # |KL_gpu - KL_numpy|: the terms differ by the last bit of log(), sums are in the same order


@pytest.fixture(scope="module")
def tr(gpu):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import tracker
    return tracker


def _mod(golden):
    d = golden["density_domain_eps"]
    return types.SimpleNamespace(domain=tuple(d[:4]), eps=float(d[4]))


def test_golden_reference_functions(tr, golden):
    mod = _mod(golden)
    Cc, Mb = golden["density_cloud_C"], golden["tci_fixed_boundary_sample_grid150"]
    for bins, sig in [(64, 1.0), (64, 0.0), (50, 2.5)]:
        tag = f"b{bins}_s{sig}"
        P_C = tr.mollified_histogram(mod, Cc, bins, sig)
        P_M = tr.mollified_histogram(mod, Mb, bins, sig)
        assert np.array_equal(P_C, golden[f"density_PC_{tag}"]) and np.array_equal(P_M, golden[f"density_PM_{tag}"])
        sc = golden[f"density_scalars_{tag}"]
        assert tr.tv_distance(P_C, P_M) == sc[0] and tr.overlap_mass(P_C, P_M) == sc[1]
        KL = tr.make_KL(mod.eps)
        assert abs(KL(P_M, P_C) - sc[2]) <= KL_ATOL * max(1.0, abs(sc[2]))
        X, T, kl0, klT = tr.gi_flow_to_threshold(KL, P_M, P_C, 0.1, 1e-6, 800, 5)
        assert T == int(sc[3]) and np.array_equal(X, golden[f"density_flow_XT_{tag}"])
        assert abs(kl0 - sc[4]) <= KL_ATOL * max(1.0, abs(sc[4])) and abs(klT - sc[5]) <= KL_ATOL
        assert tr.tv_distance(X, P_M) == sc[9]
        X, T, kl0, klT = tr.gi_flow_fixed_T(KL, P_M, P_C, 0.1, 25)
        assert T == 25 and np.array_equal(X, golden[f"density_flow_XF_{tag}"]) and abs(klT - sc[8]) <= KL_ATOL
        assert tr.fraction_outside_domain(Cc, mod.domain) == sc[10] and tr.fraction_outside_domain(Mb, mod.domain) == sc[11]


@pytest.mark.parametrize("n", [0, 1, 5, 7, 8, 9, 100, 127, 128, 129, 130, 136, 257, 1000, 4096, 7700, 12345, 64 * 64, 1024 * 1024, 1500 * 1100 + 3])
def test_sum_in_numpy_order(tr, n):
    rng = np.random.default_rng(n + 1)
    a = rng.standard_normal(n) * np.exp(rng.uniform(-8, 8, n))
    assert tr.sum_pairwise(a) == float(np.sum(a))


@pytest.mark.parametrize("bins", [1, 2, 7, 64, (33, 129), 1024])
def test_histogram2d_bit_exact(tr, bins):
    rng = np.random.default_rng(17)
    n = 150000                                                    # --mandelbrot-samples-max
    x = rng.uniform(-2.4, 1.4, n); y = rng.uniform(-1.8, 1.8, n)
    rg = [[-2.2, 1.2], [-1.6, 1.6]]
    bx, by = (bins, bins) if np.ndim(bins) == 0 else bins
    xe = np.linspace(*rg[0], bx + 1); ye = np.linspace(*rg[1], by + 1)
    # samples exactly on edges (first, interior, last), just outside, and NaN
    x[:bx + 1] = xe; y[:bx + 1] = rng.uniform(-1.6, 1.6, bx + 1)
    y[bx + 1:bx + by + 2] = ye
    x[-3:] = [np.nextafter(1.2, 2), np.nextafter(-2.2, -3), np.nan]
    H, xe2, ye2 = tr.histogram2d(x, y, (bx, by), rg)
    Href, xr, yr = np.histogram2d(x, y, bins=(bx, by), range=rg)
    assert np.array_equal(xe2, xr) and np.array_equal(ye2, yr)
    assert np.array_equal(H, Href)


@pytest.mark.parametrize("shape,sigma", [((64, 64), 1.0), ((1, 9), 1.0), ((9, 1), 2.0), ((3, 5), 3.0), ((50, 77), 2.5), ((128, 128), 0.3),
                                         ((1024, 1024), 1.0), ((257, 1023), 4.0)])
def test_gaussian_filter_bit_exact(tr, shape, sigma):
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    H = np.maximum(rng.poisson(0.4, shape).astype(np.float64), 1e-12)
    assert np.array_equal(tr.gaussian_filter(H, sigma, mode="nearest"), gaussian_filter(H, sigma=sigma, mode="nearest"))


def test_tracker_level_sizes(tr, oracle):
    """The tracker's last levels (bins 512 / 1024, 37 820 construct points, 150 000 boundary samples): P bit-exact against
    the numpy / scipy chain, the adaptive flow stops at the same sweep with the same X_T."""
    rng = np.random.default_rng(5)
    mod = types.SimpleNamespace(domain=(-2.2, 1.2, -1.6, 1.6), eps=1e-12)
    th = rng.uniform(0, 2 * np.pi, 150000)
    M = (0.5 * np.exp(1j * th) - 0.25 * np.exp(2j * th)) * (1 + 0.01 * rng.standard_normal(th.size))      # a noisy cardioid
    Cc = M[:37820] * (1 + 0.03 * rng.standard_normal(37820)) + 0.01
    KL = tr.make_KL(mod.eps)
    for bins in (512, 1024):
        P_M = tr.mollified_histogram(mod, M, bins, 1.0)
        P_C = tr.mollified_histogram(mod, Cc, bins, 1.0)
        assert np.array_equal(P_M, oracle.mollified_histogram(mod.domain, mod.eps, M, bins, 1.0))
        assert np.array_equal(P_C, oracle.mollified_histogram(mod.domain, mod.eps, Cc, bins, 1.0))
        assert tr.tv_distance(P_C, P_M) == oracle.tv_distance(P_C, P_M)
        assert tr.overlap_mass(P_C, P_M) == oracle.overlap_mass(P_C, P_M)
        ref = oracle.KL(P_M, P_C, mod.eps)
        assert abs(KL(P_M, P_C) - ref) <= KL_ATOL * max(1.0, abs(ref))
        X, T, kl0, klT = tr.gi_flow_to_threshold(KL, P_M, P_C, 0.1, 1e-6, 800, 5)
        Xr, Tr, kl0r, klTr = oracle.gi_flow(P_M, P_C, 0.1, 800, 5, 1e-6, mod.eps)
        assert T == Tr and np.array_equal(X, Xr)
        assert abs(kl0 - kl0r) <= KL_ATOL * max(1.0, abs(kl0r)) and abs(klT - klTr) <= KL_ATOL
        assert klT <= 1e-6 and 5 <= T < 800


def test_flow_edge_cases(tr, oracle):
    rng = np.random.default_rng(8)
    P = rng.random((16, 16)); P /= P.sum()
    X0 = rng.random((16, 16)); X0 /= X0.sum()
    KL = tr.make_KL(1e-12)
    # max_steps = 0: nothing happens
    X, T, kl0, klT = tr.gi_flow_to_threshold(KL, P, X0, 0.1, 1e-6, 0, 1)
    assert T == 0 and np.array_equal(X, X0) and kl0 == klT
    # threshold never met: runs max_steps sweeps
    X, T, kl0, klT = tr.gi_flow_to_threshold(KL, P, X0, 0.01, 0.0, 37, 1)
    Xr, Tr, _, klr = oracle.gi_flow(P, X0, 0.01, 37, 1, 0.0)
    assert T == Tr == 37 and np.array_equal(X, Xr) and abs(klT - klr) <= KL_ATOL
    # already below the threshold: the reference still performs min_steps sweeps (at least one)
    X, T, _, _ = tr.gi_flow_to_threshold(KL, P, P.copy(), 0.1, 1e-6, 800, 5)
    assert T == 5
    X, T, _, _ = tr.gi_flow_to_threshold(KL, P, P.copy(), 0.1, 1e-6, 800, 0)
    assert T == 1
    # stop sweeps that are not multiples of the host's polling interval
    for thr in (1e-2, 1e-3, 1e-5, 1e-8):
        X, T, _, klT = tr.gi_flow_to_threshold(KL, P, X0, 0.1, thr, 800, 1)
        Xr, Tr, _, klr = oracle.gi_flow(P, X0, 0.1, 800, 1, thr)
        assert T == Tr and np.array_equal(X, Xr) and abs(klT - klr) <= KL_ATOL
    # a KL_fn that is not ours is refused (no host loop behind the device flow)
    with pytest.raises(TypeError):
        tr.gi_flow_fixed_T(lambda p, x: 0.0, P, X0, 0.1, 3)
    with pytest.raises(ValueError):
        tr.gi_flow_fixed_T(KL, P, X0[:8], 0.1, 3)

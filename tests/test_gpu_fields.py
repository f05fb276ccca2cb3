"""GPU: K4 stencils (bit-exact), K4a log-potentials and K1b distance estimators (tolerance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 4), (4, 1), (2, 2), (3, 5), (23, 31), (200, 200), (301, 257), (400, 400), (1024, 2050), (33, 4096)])
def test_stencils_bit_exact(gpu, oracle, shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    U = rng.standard_normal(shape) * np.exp(rng.uniform(-20, 20, shape))
    h = 4.0 / 199
    assert np.array_equal(gpu.stencils.laplacian(U, h), oracle.laplacian(U, h))
    assert np.array_equal(gpu.stencils.smooth5(U), oracle.smooth5(U))


def test_stencils_golden(gpu, golden):
    U = golden["stencil_U"]; h = float(golden["stencil_h"][0])
    assert np.array_equal(gpu.stencils.laplacian(U, h), golden["stencil_laplacian"])
    assert np.array_equal(gpu.stencils.laplacian_fd(U, h), golden["stencil_laplacian_fd"])
    assert np.array_equal(gpu.stencils.smooth5(U), golden["stencil_smooth5"])


def test_stencil_linearity_full_size(gpu):
    """8192^2 field: the Laplacian of a constant is exactly 0 and of a linear-in-index ramp is 0 away
    from the periodic seam; shift-equivariance under np.roll (periodic wrap)."""
    n = 8192
    ramp = np.add.outer(np.arange(n, dtype=np.float64) * 3.0, np.arange(n, dtype=np.float64) * 5.0)
    L = gpu.stencils.laplacian(ramp, 1.0)
    assert not L[1:-1, 1:-1].any()
    U = np.random.default_rng(0).standard_normal((512, 768))
    A = gpu.stencils.laplacian(np.roll(U, (5, -9), axis=(0, 1)), 0.3)
    B = np.roll(gpu.stencils.laplacian(U, 0.3), (5, -9), axis=(0, 1))
    assert np.array_equal(A, B)


def test_log_potentials(gpu, oracle, golden):
    P = golden["logpot_points"]; gx, gy = golden["potgrid_x"], golden["potgrid_y"]
    rt = 1e-12           # relative, plus 1e-13 absolute where the sum of logs cancels to ~0
    np.testing.assert_allclose(gpu.potentials.log_potential(P, gx, gy), golden["logpot_potentials"], rtol=rt, atol=1e-13)
    X, Y = np.meshgrid(gx, gy)
    np.testing.assert_allclose(gpu.potentials.construct_potential(X, Y, P), golden["logpot_laplacian_cm"], rtol=rt, atol=1e-13)
    np.testing.assert_allclose(gpu.potentials.log_potential(P, gx, gy, use_hypot=True), golden["logpot_iterative"], rtol=rt, atol=1e-13)
    np.testing.assert_allclose(gpu.potentials.log_potential_from_points(golden["vario_grid_x"], golden["vario_grid_y"],
                                                                        P[:, 0] + 1j * P[:, 1], 1e-6),
                               golden["logpot_vario_eps1e-6"], rtol=rt, atol=1e-13)
    # bigger: 400^2 grid (Potentials.py:52-53) x 3000 points against the oracle
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1.2, 1.2, (3000, 2))
    g = np.linspace(-2, 2, 400)
    np.testing.assert_allclose(gpu.potentials.log_potential(pts, g, g), oracle.log_potential(pts, g, g, 1e-12, 0), rtol=rt, atol=1e-13)


def test_distance_estimators(gpu, oracle, golden):
    d, _ = gpu.potentials.distance_grid(golden["de_scalar_x"], golden["de_scalar_y"], 200, 1e6, 1e-16, 0)
    np.testing.assert_allclose(d, golden["de_scalar_dist"], rtol=1e-13, atol=0)
    xs = np.linspace(-2.25, 1.25, 300); ys = np.linspace(-1.75, 1.75, 280)
    for variant, R, eps in ((0, 1e6, 1e-16), (1, 4.0, 1e-14), (1, 250.0, 1e-12), (2, 250.0, 1e-12), (2, 4.0, 1e-12)):
        want, esc_o = oracle.distance_grid(xs, ys, 250, R, eps, variant)
        got, esc = gpu.potentials.distance_grid(xs, ys, 250, R, eps, variant)
        assert np.array_equal(esc, esc_o)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
    assert gpu.potentials.mandelbrot_distance_estimator(0.3 + 0.5j) == pytest.approx(
        float(oracle.distance_grid([0.3], [0.5], 200, 1e6, 1e-16, 0)[0][0, 0]), rel=1e-13)


def test_distance_estimator_tracker_module(gpu, oracle, golden):
    """LM_DE_FINAL_DZ against the reference's own output (tci_construct_mandelbrot_v002_fixed.py:35-47) and,
    on the zoom where dz stays finite, against the oracle to rounding."""
    for tag in ("tci_fixed", "tci_fixed_zoom"):
        got, esc = gpu.potentials.distance_grid(golden[tag + "_x"], golden[tag + "_y"], 250, 250.0, 1e-12, 2)
        want, want_esc = golden[tag + "_dist"], golden[tag + "_escaped"]
        assert (esc == want_esc).mean() > 0.999
        assert np.array_equal(got != 0, want != 0)
        m = want != 0
        np.testing.assert_allclose(got[m], want[m], rtol=1e-4)       # FMA-contaminated numpy arrays, see CPU test
        o, oesc = oracle.distance_grid(golden[tag + "_x"], golden[tag + "_y"], 250, 250.0, 1e-12, 2)
        assert np.array_equal(esc, oesc)
        np.testing.assert_allclose(got, o, rtol=1e-13, atol=0)
        # LM_DE_FINAL_DZ_NUMPY: numpy's FMA complex multiply restated -> the reference's masks exactly, values to rounding
        got3, esc3 = gpu.potentials.distance_grid(golden[tag + "_x"], golden[tag + "_y"], 250, 250.0, 1e-12, 3)
        assert np.array_equal(esc3, want_esc) and np.array_equal(got3 != 0, want != 0)
        np.testing.assert_allclose(got3, want, rtol=4e-15, atol=0)
        o3, oesc3 = oracle.distance_grid(golden[tag + "_x"], golden[tag + "_y"], 250, 250.0, 1e-12, 3)
        assert np.array_equal(esc3, oesc3)
        np.testing.assert_allclose(got3, o3, rtol=4e-15, atol=0)
    # the tracker's grid sizes: mask and zero pattern of the two variants against the oracle's numpy recipe
    for grid in (600, 912):
        xs = np.linspace(-2.2, 1.2, grid); ys = np.linspace(-1.6, 1.6, grid)
        got3, esc3 = gpu.potentials.distance_grid(xs, ys, 250, 250.0, 1e-12, 3)
        o3, oesc3 = oracle.distance_grid(xs, ys, 250, 250.0, 1e-12, 3)
        assert np.array_equal(esc3, oesc3) and np.array_equal(got3 != 0, o3 != 0)
        np.testing.assert_allclose(got3, o3, rtol=4e-15, atol=0)


def test_green_function_sums(gpu, oracle, golden):
    """g_real / dPhi (lucas_to_cardioid_v40_reference.py:201-257) with the O(M*N) sums on the GPU, against the
    reference's own outputs and the oracle; targets ON boundary nodes exercise the eps / dz_eps clamps."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import nystrom
    bdy, ds, sig, z = golden["green_bdy"], golden["green_ds"], golden["green_sigma"], golden["green_targets"]
    ar, ai, Cc, sh = golden["green_params"]
    a = complex(ar, ai)
    g = nystrom.g_real(z, bdy, sig, ds, a, Cc, sh)
    np.testing.assert_allclose(g, golden["green_g_real"], rtol=1e-12, atol=1e-13)
    dphi = nystrom.dPhi(z, bdy, sig, ds, a)
    np.testing.assert_allclose(dphi, golden["green_dPhi"], rtol=1e-12, atol=1e-12 * np.abs(golden["green_dPhi"]).max())
    # the reference's production size: 2*10^4 targets x 2000 nodes
    rng = np.random.default_rng(4)
    zz = rng.uniform(-2, 2, 20000) + 1j * rng.uniform(-2, 2, 20000)
    nodes = np.exp(2j * np.pi * np.arange(2000) / 2000) * (1 + 0.3 * np.cos(5 * 2 * np.pi * np.arange(2000) / 2000))
    w = rng.uniform(0.0, 0.01, 2000)
    np.testing.assert_allclose(nystrom.weighted_log_sum(zz, nodes, w), oracle.weighted_log_sum(zz, nodes, w), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(nystrom.weighted_cauchy_sum(zz, nodes, w), oracle.weighted_cauchy_sum(zz, nodes, w), rtol=1e-11, atol=1e-13)
    assert nystrom.weighted_log_sum(np.zeros(0, complex), nodes, w).shape == (0,)
    with pytest.raises(ValueError):
        nystrom.weighted_log_sum(zz, nodes, w[:-1])


def test_curvature_localpoly(gpu, oracle, golden):
    """the boundary consumer.  The fit is ill-conditioned in arclength at pixel spacing (cond ~ 1e6), so the
    GPU's scaled QR and LAPACK's SVD agree to ~1e-8 relative on the second derivatives; first derivatives and
    speed to 1e-11."""
    from helpers import assert_columns_close
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import curvature
    for tag, m, closed, stride in (("curv_closed", 7, True, 1), ("curv_open", 4, False, 3)):
        P, want = golden[tag + "_P"], golden[tag + "_out"]
        k, ks, sp, aux = curvature.compute_curvature_localpoly(P, neighbors=m, closed=closed, stride=stride)
        got = np.c_[k, ks, sp, aux["xprime"], aux["yprime"], aux["x2"], aux["y2"]]
        assert_columns_close(got, want, 1e-10, 1e-11, cols=(2, 3, 4))
        assert_columns_close(got, want, 1e-7, 1e-8, cols=(0, 1, 5, 6))
    # a circle of radius r has curvature 1/r; the real boundary of config 1 through the whole chain
    th = np.linspace(0, 2 * np.pi, 4000, endpoint=False)
    k, ks, sp, _ = curvature.compute_curvature_localpoly(np.c_[2.5 * np.cos(th), 2.5 * np.sin(th)], neighbors=7, closed=True)
    np.testing.assert_allclose(k, 0.4, rtol=1e-4)            # chord discretisation of the local fit
    assert (ks > 0).all()                                    # counter-clockwise: positive signed curvature
    xs = np.linspace(-2.1, 0.9, 600); ys = np.linspace(-1.5, 1.5, 600)
    lines, _ = gpu.contour.boundary_sample(xs, ys, 300, 288.0)
    B = gpu.contour.longest(lines)
    k, ks, sp, _ = curvature.compute_curvature_localpoly(B, neighbors=7, closed=True)
    want = oracle.curvature_localpoly(B[:400], 7, False)      # interior of an open window: same fits away from its ends
    np.testing.assert_allclose(ks[20:380], want[20:380, 1], rtol=1e-6, atol=1e-6 * np.abs(want[:, 1]).max())
    with pytest.raises(ValueError):
        curvature.compute_curvature_localpoly(B, neighbors=1)

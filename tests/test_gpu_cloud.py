"""GPU: the Lucas-Loci field stage (BASELINE.json config 5): K3 -> cloud -> K1d -> K4a -> K4 in one call
(lm_lucas_cloud_fields), the device-resident building blocks, and the point-sharded log-potential
(per-slice sums + sum + finish, what the multi-GPU path all-reduces)."""
import ctypes as C

import numpy as np
import pytest

from conftest import match_sorted_complex

pytestmark = pytest.mark.gpu


def _batch(npoly, maxdeg=25, seed=0):
    """config 5 generator."""
    rng = np.random.default_rng(seed)
    deg = rng.integers(2, maxdeg + 1, size=npoly).astype(np.int32)
    top = rng.integers(0, 3, size=(npoly, maxdeg)).astype(np.float64)
    top[np.arange(maxdeg)[None, :] >= deg[:, None]] = 0.0
    last = top[np.arange(npoly), deg - 1]
    top[np.arange(npoly), deg - 1] = np.where(last == 0, 1.0, last)
    return top, deg


def test_cloud_fields_vs_oracle(gpu, oracle):
    top, deg = _batch(1500)
    top[7] = 0.0; top[7, :3] = [1.0, 0.0, 0.0]; deg[7] = 3           # x^3 - x^2: two zero eigenvalues are filtered
    gx = np.linspace(-2, 2, 61); gy = np.linspace(-2, 2, 47)
    out = gpu.lucas.cloud_fields(top, deg, gx, gy, tol=1e-12, potential=(600, 2.0))
    cloud = out["cloud"]
    # the cloud is the concatenation, in polynomial order, of the kept 1/lambda of every polynomial
    counts = []
    for k in range(top.shape[0]):
        ref = oracle.inverse_eigenvalues_toprow(top[k, :deg[k]], 1e-12)
        counts.append(len(ref))
    assert out["n_points"] == sum(counts) == cloud.size
    assert counts[7] == 1
    off = 0
    loose = 0
    for k in range(top.shape[0]):
        if k % 10 == 0 or k == 7:
            ref = oracle.inverse_eigenvalues_toprow(top[k, :deg[k]], 1e-12)
            if match_sorted_complex(cloud[off:off + counts[k]], ref) > 1e-10:
                loose += 1                                           # clustered roots (see test_gpu_roots)
        off += counts[k]
    assert loose <= 3
    # K1d on that cloud: iteration counts bit-exact, g to rounding
    g, it, _ = oracle.batch_potential(cloud, 600, 2.0)
    assert np.array_equal(out["it"], it)
    np.testing.assert_allclose(out["g"], g, rtol=1e-14, atol=0)
    assert out["stats"]["potential_work"] == int(it.sum())
    # K4a on that cloud (Potentials.py:19-27) and K4 on that field (bit-exact stencil)
    pts = np.column_stack([cloud.real, cloud.imag])
    np.testing.assert_allclose(out["U"], oracle.log_potential(pts, gx, gy, 1e-12, 0), rtol=1e-12, atol=1e-13)
    assert np.array_equal(out["lapU"], oracle.laplacian(out["U"], gx[1] - gx[0]))
    assert out["stats"]["pairs"] == cloud.size * gx.size * gy.size


def test_cloud_fields_options(gpu, oracle):
    top, deg = _batch(300, maxdeg=12, seed=3)
    a = gpu.lucas.cloud_fields(top, deg)                                # cloud only
    assert a["U"] is None and a["g"] is None and a["cloud"].size == int(deg.sum())
    gx = np.linspace(-1.5, 1.5, 33)
    b = gpu.lucas.cloud_fields(top, deg, gx, gx, return_cloud=False, laplacian=False, eps=1e-6, variant=3)
    assert b["cloud"] is None and b["lapU"] is None
    pts = np.column_stack([a["cloud"].real, a["cloud"].imag])
    np.testing.assert_allclose(b["U"], oracle.log_potential(pts, gx, gx, 1e-6, 3), rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        gpu.lucas.cloud_fields(top, np.full(300, 13, dtype=np.int32))  # degree > maxdeg
    empty = gpu.lucas.cloud_fields(np.zeros((0, 5)), np.zeros(0, dtype=np.int32), gx, gx)
    assert empty["n_points"] == 0 and not empty["U"].any()


@pytest.mark.parametrize("variant,eps", [(0, 1e-12), (1, 1e-12), (2, 1e-12), (3, 1e-6), (0, 0.0), (0, 1e-3)])
def test_log_potential_edge_cases(gpu, oracle, variant, eps):
    """points ON grid nodes (|z-p| = 0: the eps term alone), clusters closer than eps*1e8, far-away points,
    point counts that are not a multiple of the product group, non-multiple-of-4 grid widths."""
    rng = np.random.default_rng(11)
    gx = np.linspace(-2, 2, 37); gy = np.linspace(-1, 1, 21)
    pts = rng.uniform(-2.2, 2.2, (1003, 2))
    pts[:5] = [[gx[3], gy[4]], [gx[10], gy[0]], [gx[36], gy[20]], [gx[7] + 1e-9, gy[7]], [gx[8], gy[8] - 3e-7]]
    pts[5] = [1e8, -3e7]
    if eps == 0.0:
        pts[:5] += 0.013                                      # log(0) = -inf in the reference too; keep it finite here
    got = gpu.potentials._logpot(pts[:, 0], pts[:, 1], gx, gy, eps, variant)
    want = oracle.log_potential(pts, gx, gy, eps, variant)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
    for n in (1, 7, 8, 9):
        got = gpu.potentials._logpot(pts[:n, 0], pts[:n, 1], gx, gy, eps, variant)
        np.testing.assert_allclose(got, oracle.log_potential(pts[:n], gx, gy, eps, variant), rtol=1e-12, atol=1e-13)


def test_point_sharded_sums(gpu, oracle):
    """The multi-GPU K4a path on one device: every 'rank' holds a slice of the cloud, produces raw per-cell sums,
    the sums are added (the all-reduce) and finished with the global point count."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200.device import DeviceBuffer
    rng = np.random.default_rng(2)
    pts = rng.standard_normal((50_001, 2))
    gx = np.linspace(-2, 2, 100); gy = np.linspace(-2, 2, 90)
    d_gx = DeviceBuffer(gx.nbytes); d_gx.upload(gx)
    d_gy = DeviceBuffer(gy.nbytes); d_gy.upload(gy)
    ncell = gx.size * gy.size
    total = np.zeros(ncell)
    cuts = sharding.item_slices(pts.shape[0], 3)
    assert cuts[0] == 0 and cuts[-1] == pts.shape[0]
    for a, b in zip(cuts[:-1], cuts[1:]):
        px = np.ascontiguousarray(pts[a:b, 0]); py = np.ascontiguousarray(pts[a:b, 1])
        d_px = DeviceBuffer(px.nbytes); d_px.upload(px)
        d_py = DeviceBuffer(py.nbytes); d_py.upload(py)
        d_s = DeviceBuffer(ncell * 8)
        gpu.shim.call("lm_log_potential_sums_dev", C.c_void_p(d_px.ptr), C.c_void_p(d_py.ptr), b - a, C.c_void_p(d_gx.ptr),
                      gx.size, C.c_void_p(d_gy.ptr), gy.size, 1e-12, 0, C.c_void_p(d_s.ptr), None)
        part = np.empty(ncell); d_s.download(part); gpu.shim.call("lm_stream_synchronize", None)
        total += part
    d_t = DeviceBuffer(ncell * 8); d_t.upload(total)
    d_U = DeviceBuffer(ncell * 8)
    gpu.shim.call("lm_log_potential_finish_dev", C.c_void_p(d_t.ptr), ncell, pts.shape[0], 0, C.c_void_p(d_U.ptr), None)
    U = np.empty((gy.size, gx.size)); d_U.download(U); gpu.shim.call("lm_stream_synchronize", None)
    np.testing.assert_allclose(U, oracle.log_potential(pts, gx, gy, 1e-12, 0), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(U, gpu.potentials.log_potential(pts, gx, gy), rtol=1e-13, atol=1e-14)


def test_degree_sort_mixed_classes(gpu, oracle):
    """degrees from all three solver classes (<= 32, <= 128, above) in one batch, in scrambled order."""
    degs = [2, 200, 33, 5, 128, 129, 32, 64, 1, 17]
    maxdeg = max(degs)
    top = np.zeros((len(degs), maxdeg)); deg = np.array(degs, dtype=np.int32)
    for k, d in enumerate(degs):
        top[k, :d] = 1.0
    vals, kept, iters = gpu.lucas.roots_batched(top, deg, invert=True, tol=1e-10)
    assert list(kept) == degs and (iters > 0).all()
    for k, d in enumerate(degs):
        assert match_sorted_complex(vals[k, :d], oracle.inverse_eigenvalues_toprow(np.ones(d), 1e-10)) < 1e-9


def test_sharded_cloud_fields_world1(gpu, oracle):
    """The multi-GPU cloud stage (sharding.sharded_cloud_fields) at world size 1: same fields as the fused single call."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    top, deg = _batch(800, seed=5)
    gx = np.linspace(-2, 2, 40); gy = np.linspace(-2, 2, 36)
    a = gpu.lucas.cloud_fields(top, deg, gx, gy, potential=(700, 2.0))
    b = sharding.sharded_cloud_fields(top, deg, gx, gy, potential=(700, 2.0))
    assert b["n_points_total"] == a["n_points"] and np.array_equal(b["cloud"], a["cloud"])
    assert np.array_equal(b["it"], a["it"]) and np.array_equal(b["g"], a["g"])
    np.testing.assert_allclose(b["U"], a["U"], rtol=1e-13, atol=1e-14)
    assert np.array_equal(b["lapU"], oracle.laplacian(b["U"], gx[1] - gx[0]))


def test_cloud_fields_int8_rows(gpu):
    """lm_lucas_cloud_fields_i8: small-integer first rows handed over as int8 give bit-identical outputs, across the
    2^20-polynomial chunk boundary of the streaming pipeline too."""
    rng = np.random.default_rng(21)
    npoly, maxdeg = (1 << 20) + 4097, 12
    deg = rng.integers(2, maxdeg + 1, npoly).astype(np.int32)
    top = rng.integers(0, 3, (npoly, maxdeg)).astype(np.float64)     # the config-5 coefficient set
    top[np.arange(maxdeg)[None, :] >= deg[:, None]] = 0.0
    top[np.arange(npoly), deg - 1] = rng.integers(1, 3, npoly)         # no zero eigenvalue
    gx = np.linspace(-2, 2, 24)
    a = gpu.lucas.cloud_fields(top, deg, gx, gx, potential=(200, 2.0))
    b = gpu.lucas.cloud_fields(top.astype(np.int8), deg, gx, gx, potential=(200, 2.0))
    assert a["n_points"] == b["n_points"] == int(deg.sum())
    assert np.array_equal(a["cloud"], b["cloud"]) and np.array_equal(a["U"], b["U"]) and np.array_equal(a["lapU"], b["lapU"])
    assert np.array_equal(a["g"], b["g"]) and np.array_equal(a["it"], b["it"])
    small = gpu.lucas.cloud_fields(np.array([[1, 1, 0], [2, 1, 1]], dtype=np.int8), np.array([2, 3], dtype=np.int32))
    ref = gpu.lucas.cloud_fields(np.array([[1.0, 1, 0], [2, 1, 1]]), np.array([2, 3], dtype=np.int32))
    assert np.array_equal(small["cloud"], ref["cloud"])
    empty = gpu.lucas.cloud_fields(np.zeros((0, 5), dtype=np.int8), np.zeros(0, dtype=np.int32))
    assert empty["n_points"] == 0
    # caller-provided (page-locked) cloud buffers: views of exactly n_points entries come back; too small -> LM_E_CAP
    n = int(deg[:5000].sum())
    cre = gpu.shim.pinned_empty(n + 7, np.float64); cim = gpu.shim.pinned_empty(n + 7, np.float64)
    c = gpu.lucas.cloud_fields(top[:5000].astype(np.int8), deg[:5000], cloud_out=(cre, cim))
    assert c["n_points"] == n and c["cloud"][0].size == n
    assert np.array_equal(c["cloud"][0] + 1j * c["cloud"][1], a["cloud"][:n])
    with pytest.raises(RuntimeError):
        gpu.lucas.cloud_fields(top[:5000], deg[:5000], cloud_out=(cre[:n - 1], cim[:n - 1]))


@pytest.mark.parametrize("tag,family", [("lucas", None), ("pell", "pell_like_all_twos")])
def test_per_n_and_cumulative_stats_golden(gpu, golden, tag, family):
    """per_n_stats / cumulative_stats (lucas_equipotential_test_v3.py:294-327) as ONE K3 batch + ONE K1d pass, against the
    rows produced by the reference's own functions (MAX_ITER 3000, n = 2..30): counts exact, order statistics to 1e-9
    (the roots agree with LAPACK's to 1e-10, the potential is a smooth function of the point)."""
    cols = ["count", "escaped", "escaped_frac", "g_median", "g_mean", "g_std", "g_p10", "g_p90"]
    for fn, key, xkey in ((gpu.lucas.per_n_stats, f"per_n_stats_{tag}_2_30_mi3000", "n"),
                          (gpu.lucas.cumulative_stats, f"cumulative_stats_{tag}_2_30_mi3000", "N")):
        rows = fn(2, 30, family=family, max_iter=3000, quiet=True)
        want = golden[key]
        assert [r[xkey] for r in rows] == list(range(2, 31))
        got = np.array([[r[c] for c in cols] for r in rows], dtype=np.float64)
        assert np.array_equal(got[:, :2], want[:, :2])                     # count, escaped
        np.testing.assert_allclose(got[:, 2:], want[:, 2:], rtol=1e-9, atol=1e-12, equal_nan=True)

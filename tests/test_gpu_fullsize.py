"""GPU: BASELINE.json configs 3, 4 and 5 at FULL size, checked through size-independent properties
(the oracle cannot finish these in seconds): exact work-count identity, agreement of independently
computed strips with the same rows of the full grid, conjugation symmetry, closed boundary loop,
known-answer polynomials embedded in the batch, backward error of a random sample of the roots."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _work_identity(d: np.ndarray, max_iter: int) -> int:
    total = 0
    for k in range(0, d.shape[0], 2048):                     # chunked: the full grid is 4.3 GB
        blk = d[k:k + 2048].astype(np.int64)
        total += int(np.minimum(blk + 1, max_iter).sum())
    return total


def test_config3_full(gpu, oracle):
    """32768^2, max_iter 10000 (1.8e12 pixel-iterations), fused K1 -> K2 with the dwell grid returned."""
    res, mi = 32768, 10000
    xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res)
    out = gpu.shim.pinned_empty((res, res), np.int32)
    lines, st = gpu.contour.boundary_sample(xs, ys, mi, 0.96 * mi, dwell_out=out)
    assert st["work_units"] == _work_identity(out, mi)
    assert 1.79e12 < st["work_units"] < 1.81e12                  # ~1.80e12
    # rows recomputed on their own (different tile schedule) and by the CPU oracle (a thin slice)
    for r0 in (0, 9000, 16380, 32760):
        strip, _, _ = gpu.escape.escape_grid(xs, ys[r0:r0 + 8], mi)
        assert np.array_equal(strip, out[r0:r0 + 8])
    cols = slice(20000, 20256)
    want, _ = oracle.dwell_grid(xs[cols], ys[16000:16002], mi)
    assert np.array_equal(out[16000:16002, cols], want)
    mirrored, _, _ = gpu.escape.escape_grid(xs, -ys[9000:9008], mi)
    assert np.array_equal(mirrored, out[9000:9008])
    best = gpu.contour.longest(lines)
    assert best.shape[0] > 300000 and np.array_equal(best[0], best[-1])
    # every vertex of the boundary sits on a grid line and separates a pixel > level from one <= level
    fx = (best[:, 0] - xs[0]) / (xs[1] - xs[0]); fy = (best[:, 1] - ys[0]) / (ys[1] - ys[0])
    assert ((np.abs(fx - np.rint(fx)) < 1e-6) | (np.abs(fy - np.rint(fy)) < 1e-6)).all()


def test_config4_seahorse_full(gpu, oracle):
    """Seahorse valley, 16384^2, max_iter 100000 (~9.2e12 pixel-iterations): divergence stress."""
    res, mi = 16384, 100000
    xs = np.linspace(-0.755, -0.735, res); ys = np.linspace(0.10, 0.12, res)
    with gpu.device.DeviceGrid(xs, ys) as g:
        g.escape(mi)
        work = g.work_units()
        d = g.dwell(pinned=True)
        lines = g.contour(0.96 * mi)
    assert work == _work_identity(d, mi)
    assert 8.0e12 < work < 1.0e13
    assert 0.30 < (d == mi).mean() < 0.38                       # 34 % interior
    for r0 in (0, 5000, 16376):
        strip, _, _ = gpu.escape.escape_grid(xs, ys[r0:r0 + 8], mi)
        assert np.array_equal(strip, d[r0:r0 + 8])
    want, _ = oracle.dwell_grid(xs[8000:8064], ys[8000:8001], mi)
    assert np.array_equal(d[8000:8001, 8000:8064], want)
    assert len(lines) > 100 and max(len(l) for l in lines) > 10000


def test_config5_lucas_cloud_full(gpu, oracle):
    """10^7 polynomials of degree <= 25 (config 5 generator): the first 96 are the four
    named families x n = 2..25 (known answers), a random sample of the rest is checked by backward
    error and against numpy.linalg.eigvals."""
    from conftest import match_sorted_complex
    from helpers import multiset_distance, true_roots_mp
    npoly, maxdeg = 10_000_000, 25
    rng = np.random.default_rng(0)
    deg = rng.integers(2, 26, size=npoly).astype(np.int32)
    top = rng.integers(0, 3, size=(npoly, maxdeg)).astype(np.float64)
    top[np.arange(maxdeg)[None, :] >= deg[:, None]] = 0.0
    last = top[np.arange(npoly), deg - 1]
    top[np.arange(npoly), deg - 1] = np.where(last == 0, 1.0, last)
    k = 0
    for fam in gpu.lucas.FAMILIES:
        for n in range(2, 26):
            top[k] = 0.0; top[k, :n] = gpu.lucas.family_toprow(fam, n); deg[k] = n; k += 1
    vals, kept, iters = gpu.lucas.roots_batched(top, deg, invert=True, tol=1e-12, sort=False)
    st = gpu.lucas.last_stats
    assert (iters > 0).all() and (kept[96:] == deg[96:]).all()     # a_d >= 1: no zero eigenvalue to filter
    assert st["work_units"] == int(deg.sum())
    k = 0
    for fam in gpu.lucas.FAMILIES:
        for n in range(2, 26):
            ref = oracle.inverse_eigenvalues_toprow(oracle.family_toprow(fam, n), 1e-12)
            assert kept[k] == len(ref)            # e.g. x^2 - x has the eigenvalue 0, which the reference filters too
            assert match_sorted_complex(vals[k, :kept[k]], ref) < 1e-10
            k += 1
    sample = rng.choice(npoly, size=400, replace=False)
    arbitrated = 0
    for s in sample:
        d = int(deg[s]); lam = 1.0 / vals[s, :d]
        c = np.concatenate([[1.0], -top[s, :d]])
        for r in lam:
            if abs(r) <= 1:
                be = abs(np.polyval(c, r)) / np.polyval(np.abs(c), abs(r))
            else:
                be = abs(np.polyval(c[::-1], 1 / r)) / np.polyval(np.abs(c[::-1]), abs(1 / r))
            assert be < 1e-13
        ref = oracle.eigvals_toprow(top[s, :d])
        if match_sorted_complex(lam, ref) > 1e-10:
            # clustered roots: a 60-digit solve arbitrates between the kernel and LAPACK (see test_gpu_roots)
            truth = true_roots_mp(top[s, :d])
            e_cuda, e_lapack = multiset_distance(lam, truth), multiset_distance(ref, truth)
            assert e_cuda <= max(16.0 * e_lapack, 1e-10), (int(s), e_cuda, e_lapack)
            arbitrated += 1
    assert arbitrated <= 8

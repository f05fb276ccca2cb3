"""CPU: the host logic of the sub-sampled semivariograms (pairstats.sample_semivariogram / sample_cross_semivariogram:
numpy global-stream draws, 4000-point chunks, per-bin caps, the single block per bin in which a random subset is drawn)
against outputs of the reference's own functions (variograms_construct_mandelbrot.py:178-315, tests/golden), with the
two device calls replaced by numpy restatements.  The GPU test (test_gpu_pairstats.py) runs the real kernels."""
import types

import numpy as np
import pytest


def _np_pair_histogram(locs, lo, hi, values=None, weight="none"):
    P = np.asarray(locs, dtype=np.float64)
    lo = np.asarray(lo, dtype=np.float64); hi = np.asarray(hi, dtype=np.float64)
    counts = np.zeros(lo.size, dtype=np.uint64); sums = np.zeros(lo.size)
    n = P.shape[0]
    v = None if values is None else np.asarray(values, dtype=np.float64)
    for a in range(0, n, 1000):
        dx = P[a:a + 1000, None, 0] - P[None, :, 0]; dy = P[a:a + 1000, None, 1] - P[None, :, 1]
        D = np.sqrt(dx * dx + dy * dy)
        upper = np.arange(n)[None, :] > np.arange(a, min(a + 1000, n))[:, None]
        w = (v[a:a + 1000, None] - v[None, :]) ** 2 if v is not None else None
        for k in range(lo.size):
            m = (D >= lo[k]) & (D < hi[k]) & upper
            counts[k] += np.uint64(m.sum())
            if w is not None:
                sums[k] += w[m].sum()
    return counts, sums


def _np_select(A, B, lo, hi, skip_diag, expect):
    xa, ya, va = A; xb, yb, vb = B
    dx = xa[:, None] - xb[None, :]; dy = ya[:, None] - yb[None, :]
    D = np.sqrt(dx * dx + dy * dy)
    m = (D >= lo) & (D < hi)
    if skip_diag:
        m &= ~np.eye(D.shape[0], D.shape[1], dtype=bool)
    out = ((va[:, None] - vb[None, :]) ** 2)[np.where(m)]
    assert out.size == expect
    return out


@pytest.mark.parametrize("tag", ["a", "b"])
def test_subsampled_semivariograms_host_logic(golden, monkeypatch, tag):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import pairstats
    monkeypatch.setattr(pairstats, "pair_histogram", _np_pair_histogram)
    monkeypatch.setattr(pairstats, "_select_sqdiff", _np_select)
    x0, x1, y0, y1, nx, ny = golden["semivario_grid_args"]
    X, Y = np.meshgrid(np.linspace(x0, x1, int(nx)), np.linspace(y0, y1, int(ny)), indexing="xy")
    grid = types.SimpleNamespace(X=X, Y=Y)
    bins = golden[f"semivario_{tag}_bins"]; cap = int(golden[f"semivario_{tag}_cap"][0])
    np.random.seed(777)
    rc, gam = pairstats.sample_semivariogram(golden["semivario_field1"], grid, bins, max_pairs_per_bin=cap)
    assert np.array_equal(rc, golden[f"semivario_{tag}_centers"])
    np.testing.assert_allclose(gam, golden[f"semivario_{tag}_gamma"], rtol=1e-12, atol=0)
    np.random.seed(778)
    rc, gam = pairstats.sample_cross_semivariogram(golden["semivario_field1"], golden["semivario_field2"], grid, bins, max_pairs_per_bin=cap)
    np.testing.assert_allclose(gam, golden[f"semivario_{tag}_cross_gamma"], rtol=1e-12, atol=0)

"""Pins the contour restatement to the real thing wherever matplotlib (+ contourpy) is installed.

The reference calls plt.contour(xs, ys, Z, levels=[level]) and keeps the longest line
(mandelbrot_boundary_sample.py:41-54; cs.allsegs[0] in mandelbrot_boundary_sample_spyder.py:35-43).  Neither
matplotlib nor contourpy exists in the build image, so oracle/lm_oracle_contour.c restates contourpy's mpl2014
algorithm from its published source and DESIGN.md marks K2 parity "unpinned".  These tests are the pin: they are
skipped where matplotlib is missing and compare, line for line and bit for bit, where it is present.
"""
import numpy as np
import pytest

mpl = pytest.importorskip("matplotlib")
mpl.use("Agg")
plt = pytest.importorskip("matplotlib.pyplot")


def _allsegs(xs, ys, Z, level):
    with mpl.rc_context({"contour.algorithm": "mpl2014"}):
        fig = plt.figure()
        try:
            cs = plt.contour(xs, ys, Z, levels=[level])
            return [np.asarray(seg, dtype=np.float64) for seg in cs.allsegs[0]]
        finally:
            plt.close(fig)


def _same(a, b):
    return len(a) == len(b) and all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("res,mi", [(96, 200), (257, 300)])
def test_oracle_matches_matplotlib_on_mandelbrot_windows(oracle, res, mi):
    xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res)
    d, _ = oracle.dwell_grid(xs, ys, mi)
    Z = d.astype(np.float64)
    want = _allsegs(xs, ys, Z, 0.96 * mi)
    got = oracle.contour_lines(xs, ys, Z, 0.96 * mi)
    assert _same(got, want)
    best = oracle.extract_contour(xs, ys, Z, mi, 0.96)
    assert np.array_equal(best, max(want, key=len))


def test_oracle_matches_matplotlib_on_random_fields(oracle):
    """integer fields full of saddles, lines cut by the border, values equal to the level."""
    rng = np.random.default_rng(0)
    for shape in ((7, 9), (20, 33), (64, 50)):
        Z = rng.integers(0, 6, size=shape).astype(np.float64)
        xs = np.linspace(0.0, 1.0, shape[1]); ys = np.linspace(-1.0, 2.0, shape[0])
        for level in (0.5, 2.0, 2.5, 4.0):
            assert _same(oracle.contour_lines(xs, ys, Z, level), _allsegs(xs, ys, Z, level)), (shape, level)


@pytest.mark.gpu
def test_gpu_matches_matplotlib(gpu):
    xs = np.linspace(-2.1, 0.9, 400); ys = np.linspace(-1.5, 1.5, 380)
    lines, _ = gpu.contour.boundary_sample(xs, ys, 300, 288.0)
    d, _, _ = gpu.escape.escape_grid(xs, ys, 300)
    want = _allsegs(xs, ys, d.astype(np.float64), 288.0)
    assert _same(list(lines), want)

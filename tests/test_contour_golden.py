"""Pins K2 to matplotlib's own output through a committed fixture (tests/golden/contour_mpl2014.npz, written by
oracle/gen_contour_golden.py on a machine that has matplotlib + contourpy).  The fixture does not exist yet -- the build
image has no matplotlib (parity unpinned) -- so these tests skip until somebody commits it; from then
on they run everywhere, CPU oracle and CUDA path alike, without matplotlib."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
FIXTURE = ROOT / "tests" / "golden" / "contour_mpl2014.npz"
pytestmark = pytest.mark.skipif(not FIXTURE.exists(), reason="contour_mpl2014.npz not generated yet (needs matplotlib once)")


def _cases():
    sys.path.insert(0, str(ROOT))
    from oracle.gen_contour_golden import cases
    return cases()


def _want(g, name):
    v, o = g[name + "_verts"], g[name + "_offsets"]
    return [v[o[k]:o[k + 1]] for k in range(len(o) - 1)]


def _same(a, b):
    return len(a) == len(b) and all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b))


def test_oracle_matches_the_matplotlib_fixture(oracle):
    with np.load(FIXTURE) as g:
        for name, xs, ys, Z, level in _cases():
            assert _same(oracle.contour_lines(xs, ys, Z, level), _want(g, name)), name


@pytest.mark.gpu
def test_gpu_matches_the_matplotlib_fixture(gpu):
    with np.load(FIXTURE) as g:
        for name, xs, ys, Z, level in _cases():
            got = gpu.contour.contour_lines(xs, ys, Z.astype(np.int32), level)
            assert _same(list(got), _want(g, name)), name

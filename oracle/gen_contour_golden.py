#!/usr/bin/env python3
"""Generate tests/golden/contour_mpl2014.npz with the REAL plt.contour (matplotlib + contourpy, algorithm mpl2014).

The build image and the GPU boxes have neither matplotlib nor contourpy, so K2 parity is marked "unpinned"
(DESIGN.md section 2).  Run this ONCE on any machine that has them,

    python oracle/gen_contour_golden.py          # writes tests/golden/contour_mpl2014.npz (a few hundred KB)

and commit the file: tests/test_contour_golden.py then pins the oracle restatement (CPU) and the CUDA path (GPU) to
matplotlib's own lines, bit for bit, with no skip -- windows of the Mandelbrot dwell field, integer levels, fields full
of saddles, lines cut by every border, values equal to the level.  The inputs are regenerated from seeds and formulas
by the test, only matplotlib's OUTPUT is stored (vertices and line offsets per case).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tests" / "golden" / "contour_mpl2014.npz"


def cases():
    """(name, xs, ys, Z float64, level) -- shared with tests/test_contour_golden.py."""
    from oracle import oracle
    out = []
    for res, mi, frac in ((96, 200, 0.96), (257, 300, 0.96), (180, 120, 0.5), (150, 64, 10.0 / 64)):
        xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res + 3)
        d, _ = oracle.dwell_grid(xs, ys, mi)
        out.append((f"mandel_{res}_{mi}_{frac:.3f}", xs, ys, d.astype(np.float64), frac * mi))
    xs = np.linspace(-0.755, -0.735, 200); ys = np.linspace(0.10, 0.12, 190)
    d, _ = oracle.dwell_grid(xs, ys, 400)
    out.append(("seahorse_200_400", xs, ys, d.astype(np.float64), 0.96 * 400))     # lines cut by the window border
    rng = np.random.default_rng(0)
    for shape in ((7, 9), (20, 33), (64, 50), (3, 2), (2, 40)):
        Z = rng.integers(0, 6, size=shape).astype(np.float64)
        xs = np.linspace(0.0, 1.0, shape[1]); ys = np.linspace(-1.0, 2.0, shape[0])
        for level in (0.5, 2.0, 2.5, 4.0):
            out.append((f"random_{shape[0]}x{shape[1]}_{level}", xs, ys, Z, level))
    return out


def main() -> None:
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    g = {}
    with matplotlib.rc_context({"contour.algorithm": "mpl2014"}):
        for name, xs, ys, Z, level in cases():
            fig = plt.figure()
            cs = plt.contour(xs, ys, Z, levels=[level])
            segs = [np.asarray(s, dtype=np.float64) for s in cs.allsegs[0]]
            plt.close(fig)
            g[name + "_verts"] = np.concatenate(segs) if segs else np.zeros((0, 2))
            g[name + "_offsets"] = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
    g["_versions"] = np.array([matplotlib.__version__, __import__("contourpy").__version__])
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB, {len(cases())} cases)")


if __name__ == "__main__":
    main()

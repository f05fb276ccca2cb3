"""Pure-Python restatement of the reference's escape-time loop, at the reference's own speed.

TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): it is never imported by the product.

    mandelbrot_dwell, compute_grid        mandelbrot_boundary_sample.py:22-39

The reference IS this double loop over CPython complex numbers; /root/reference does not exist on the GPU box, so
bench.py times this restatement there (cpu_baseline "python" entry, kind "port") and, where the reference checkout
is present (the build container), the reference's own `def`s loaded by path (kind "reference").
Pinned: tests/test_oracle_golden.py checks it against the dwell grids produced by the reference's own functions
(tests/golden/reference_vectors.npz, oracle/gen_golden.py).
"""
from __future__ import annotations

import ast
import time
from pathlib import Path

import numpy as np


def mandelbrot_dwell(x: float, y: float, max_iter: int = 300) -> int:
    """mandelbrot_boundary_sample.py:22-31 -- CPython complex arithmetic: z*z + c is
    (zr*zr - zi*zi) + cr, (zr*zi + zi*zr) + ci, unfused binary64."""
    c = complex(x, y)
    z = 0j
    for n in range(max_iter):
        z = z * z + c
        if (z.real * z.real + z.imag * z.imag) > 4.0:
            return n
    return max_iter


def dwell_rows(xs, ys_rows, max_iter: int) -> np.ndarray:
    """compute_grid's double loop (mandelbrot_boundary_sample.py:32-39) over the given rows only."""
    Z = np.zeros((len(ys_rows), len(xs)), dtype=float)
    for j, y in enumerate(ys_rows):
        for i, x in enumerate(xs):
            Z[j, i] = mandelbrot_dwell(x, y, max_iter=max_iter)
    return Z


def reference_defs(ref_dir: str = "/root/reference"):
    """The reference's own mandelbrot_dwell (its `def` block executed as is), or None when the checkout is absent."""
    f = Path(ref_dir) / "mandelbrot_boundary_sample.py"
    if not f.exists():
        return None
    tree = ast.parse(f.read_text().replace("\r\n", "\n"))
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("mandelbrot_dwell",)]
    ns = {"np": np}
    exec(compile(ast.Module(body=picked, type_ignores=[]), str(f), "exec"), ns)
    return ns["mandelbrot_dwell"]


def timed_sample(xs, ys, max_iter: int, budget_s: float = 10.0, cols_stride: int = 1, seed: int = 0):
    """Time the Python loop on random rows of the workload (every cols_stride-th column) until budget_s is used.
    -> dict(value [G pixel-iter/s], pixel_iters, seconds, rows, kind)"""
    fn = reference_defs()
    kind = "reference" if fn is not None else "port"
    if fn is None:
        fn = mandelbrot_dwell
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(ys))
    xcols = np.asarray(xs)[::cols_stride]
    work = 0
    rows = 0
    t0 = time.perf_counter()
    for j in order:
        y = float(ys[j])
        for x in xcols:
            d = fn(float(x), y, max_iter=max_iter)
            work += min(d + 1, max_iter)
        rows += 1
        if time.perf_counter() - t0 > budget_s or rows >= 64:
            break
    dt = time.perf_counter() - t0
    return {"value": work / dt / 1e9, "pixel_iters": int(work), "seconds": dt, "rows": rows, "kind": kind,
            "cols": int(len(xcols))}

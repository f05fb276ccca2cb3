"""Timing of the pair-statistics kernels at the tracker's cloud sizes; prints kernel ms and pairs/s."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, pairstats as ps

_shim.set_device(0)
rng = np.random.default_rng(2)
for n in (2400, 14820, 37820, 150000):
    th = rng.uniform(0, 2 * np.pi, n)
    P = np.c_[-0.5 + 1.2 * np.cos(th) * (1 - 0.5 * np.cos(th)), 1.2 * np.sin(th) * (1 - 0.5 * np.cos(th))] + 0.02 * rng.standard_normal((n, 2))
    v = np.hypot(P[:, 0], P[:, 1])
    for order in ("random", "sorted"):
        Q, w = (P, v) if order == "random" else (P[np.argsort(th)], v[np.argsort(th)])
        for weight, nb in (("none", 150), ("value", 60), ("dist2", 60), ("value", 1500)):
            e = np.linspace(0.0, 1.5, nb + 1)
            ps.pair_histogram(Q, e[:-1], e[1:], w, weight)
            t0 = time.perf_counter(); ps.pair_histogram(Q, e[:-1], e[1:], w, weight); t1 = time.perf_counter()
            st = ps.last_stats
            print(f"n={n} {order} weight={weight} nbins={nb}: kernel {st['kernel_ms']:.3f} ms, {st['work_units'] / st['kernel_ms'] / 1e6:.1f} G pairs/s, call {1e3 * (t1 - t0):.2f} ms")
    t0 = time.perf_counter(); d = ps.max_pair_distance(P); t1 = time.perf_counter()
    print(f"n={n} max distance {d:.6f}: call {1e3 * (t1 - t0):.2f} ms")

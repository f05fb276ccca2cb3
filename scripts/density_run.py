"""Timing of the tracker's density stage at bins 256/512/1024 against the numpy / scipy chain."""
import sys, time, types
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, tracker as tr
from oracle import oracle as orc

_shim.set_device(0)
rng = np.random.default_rng(5)
mod = types.SimpleNamespace(domain=(-2.2, 1.2, -1.6, 1.6), eps=1e-12)
th = rng.uniform(0, 2 * np.pi, 150000)
M = (0.5 * np.exp(1j * th) - 0.25 * np.exp(2j * th)) * (1 + 0.01 * rng.standard_normal(th.size))
Cc = M[:37820] * (1 + 0.03 * rng.standard_normal(37820)) + 0.01
KL = tr.make_KL(mod.eps)
for bins in (256, 512, 1024):
    tr.mollified_histogram(mod, M, bins, 1.0)
    t0 = time.perf_counter(); P_M = tr.mollified_histogram(mod, M, bins, 1.0); t1 = time.perf_counter()
    k_ms = tr.last_stats["kernel_ms"]
    P_C = tr.mollified_histogram(mod, Cc, bins, 1.0)
    t2 = time.perf_counter(); R_M = orc.mollified_histogram(mod.domain, mod.eps, M, bins, 1.0); t3 = time.perf_counter()
    print(f"bins={bins} mollified_histogram: device {k_ms:.3f} ms, call {1e3 * (t1 - t0):.2f} ms, numpy/scipy {1e3 * (t3 - t2):.2f} ms, equal={np.array_equal(P_M, R_M)}")
    tr.gi_flow_to_threshold(KL, P_M, P_C, 0.1, 1e-6, 800, 5)
    t0 = time.perf_counter(); X, T, kl0, klT = tr.gi_flow_to_threshold(KL, P_M, P_C, 0.1, 1e-6, 800, 5); t1 = time.perf_counter()
    st = tr.last_stats
    t2 = time.perf_counter(); Xr, Tr, kl0r, klTr = orc.gi_flow(P_M, P_C, 0.1, 800, 5, 1e-6, mod.eps); t3 = time.perf_counter()
    print(f"bins={bins} gi_flow_to_threshold: T={T} (numpy {Tr}) device {st['kernel_ms']:.3f} ms ({st['launches']} launches), call {1e3 * (t1 - t0):.2f} ms, "
          f"numpy {1e3 * (t3 - t2):.1f} ms, X equal={np.array_equal(X, Xr)}, |dKL|={abs(klT - klTr):.2e}")
    t0 = time.perf_counter(); a = tr.density_compare(P_C, P_M); t1 = time.perf_counter()
    t2 = time.perf_counter(); b = (2 * orc.tv_distance(P_C, P_M), orc.overlap_mass(P_C, P_M), orc.KL(P_C, P_M)); t3 = time.perf_counter()
    print(f"bins={bins} tv/overlap/KL: call {1e3 * (t1 - t0):.2f} ms, numpy {1e3 * (t3 - t2):.2f} ms, {a[0] == b[0]} {a[1] == b[1]} {abs(a[2] - b[2]):.1e}")

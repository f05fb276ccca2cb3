"""One invocation of the pair-statistics and density-flow kernels for ncu (see profiles/INDEX.md)."""
import sys, types
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, pairstats as ps, tracker as tr

_shim.set_device(0)
rng = np.random.default_rng(2)
n = 37820
th = rng.uniform(0, 2 * np.pi, n)
P = np.c_[-0.5 + 1.2 * np.cos(th) * (1 - 0.5 * np.cos(th)), 1.2 * np.sin(th) * (1 - 0.5 * np.cos(th))] + 0.02 * rng.standard_normal((n, 2))
e = np.linspace(0.0, 1.5, 61)
ps.pair_histogram(P, e[:-1], e[1:], np.hypot(P[:, 0], P[:, 1]), "value")
mod = types.SimpleNamespace(domain=(-2.2, 1.2, -1.6, 1.6), eps=1e-12)
t2 = rng.uniform(0, 2 * np.pi, 150000)
M = (0.5 * np.exp(1j * t2) - 0.25 * np.exp(2j * t2)) * (1 + 0.01 * rng.standard_normal(t2.size))
P_M = tr.mollified_histogram(mod, M, 1024, 1.0)
P_C = tr.mollified_histogram(mod, M[:37820] * 1.01 + 0.01, 1024, 1.0)
print(tr.gi_flow_fixed_T(tr.KL, P_M, P_C, 0.1, 8)[1:])

"""lm_nearest_match at the tracker's largest size (25 000 x 25 000)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import potentials
rng = np.random.default_rng(0)
X = rng.standard_normal(25000) + 1j * rng.standard_normal(25000); Y = rng.standard_normal(25000) + 1j * rng.standard_normal(25000)
for _ in range(2):
    t0 = time.perf_counter(); i, d = potentials.nearest_match(X, Y); dt = time.perf_counter() - t0
    st = potentials.last_stats
    print(f"nearest_match 25000x25000: kernel {st['kernel_ms']:.2f} ms = {st['work_units']/st['kernel_ms']/1e6:.0f} G pairs/s, host {dt*1e3:.1f} ms")

"""K4a alone: 400^2 grid x N points (ncu target)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import potentials
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
pts = np.random.default_rng(1).uniform(-1.8, 1.8, (npts, 2)); g = np.linspace(-2, 2, 400)
for variant, eps in ((0, 1e-12), (3, 1e-6)):
    for rep in range(2):
        potentials._logpot(pts[:, 0], pts[:, 1], g, g, eps, variant); st = potentials.last_stats
        print(f"K4a variant {variant} eps {eps:g}: {npts} pts kernel {st['kernel_ms']:.2f} ms = {st['work_units']/st['kernel_ms']/1e6:.0f} G pairs/s", flush=True)

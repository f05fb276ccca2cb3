"""K3 alone on a config-5 shaped batch (ncu target)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import lucas
npoly = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
deg = rng.integers(2, 26, size=npoly).astype(np.int32)
top = rng.integers(0, 3, size=(npoly, 25)).astype(np.float64)
top[np.arange(25)[None, :] >= deg[:, None]] = 0.0
last = top[np.arange(npoly), deg - 1]
top[np.arange(npoly), deg - 1] = np.where(last == 0, 1.0, last)
for rep in range(2):
    vals, kept, iters = lucas.roots_batched(top, deg, invert=True, tol=1e-12, sort=False)
    st = lucas.last_stats
    print(f"K3: {npoly} polys {st['work_units']} roots kernel {st['kernel_ms']:.2f} ms = {st['work_units']/st['kernel_ms']/1e3:.0f} M roots/s, "
          f"sweeps mean {iters.mean():.2f}", flush=True)

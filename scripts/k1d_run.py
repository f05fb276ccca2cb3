"""K1d (batch_potential) on a config-5 cloud: kernel time single-pass vs two-pass."""
import sys, time, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import escape, lucas
npoly = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
deg = rng.integers(2, 26, size=npoly).astype(np.int32)
top = rng.integers(0, 3, size=(npoly, 25)).astype(np.float64)
top[np.arange(25)[None, :] >= deg[:, None]] = 0.0
last = top[np.arange(npoly), deg - 1]
top[np.arange(npoly), deg - 1] = np.where(last == 0, 1.0, last)
cloud = lucas.cloud_fields(top, deg)["cloud"]
for mode in (sys.argv[2:] or ["0", "1"]):
    os.environ["LM_K1D_TWO_PASS"] = mode
    for rep in range(2):
        g, it, phi = escape.batch_potential(cloud, 20000, 2.0)
        st = escape.last_stats
        print(f"two_pass={mode}: {cloud.size} pts, kernel {st['kernel_ms']:.2f} ms, {st['work_units']/st['kernel_ms']/1e6:.0f} Gpi/s, launches {st['launches']}", flush=True)

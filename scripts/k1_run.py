"""Run K1 (dwell-only) on one configuration a few times; used as the ncu target."""
import sys, argparse
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import escape
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=8192)
ap.add_argument("--max_iter", type=int, default=2000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--xlim", nargs=2, type=float, default=[-2.1, 0.9])
ap.add_argument("--ylim", nargs=2, type=float, default=[-1.5, 1.5])
ap.add_argument("--field", type=int, default=0)
a = ap.parse_args()
xs = np.linspace(a.xlim[0], a.xlim[1], a.res); ys = np.linspace(a.ylim[0], a.ylim[1], a.res)
for r in range(a.reps):
    d, f, st = escape.escape_grid(xs, ys, a.max_iter, 2.0, a.field)
    print(f"res={a.res} mi={a.max_iter} kernel_ms={st['kernel_ms']:.3f} work={st['work_units']} Gpi/s={st['work_units']/st['kernel_ms']/1e6:.1f}", flush=True)

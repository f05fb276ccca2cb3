"""K1 once (resident), then K2 a few times from the device; ncu target for the contour kernels."""
import sys, time, argparse
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import contour, device
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=16384)
ap.add_argument("--max_iter", type=int, default=1000)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
xs = np.linspace(-2.1, 0.9, a.res); ys = np.linspace(-1.5, 1.5, a.res)
with device.DeviceGrid(xs, ys) as g:
    g.escape(a.max_iter)
    for r in range(a.reps):
        t0 = time.perf_counter()
        lines = g.contour(0.96 * a.max_iter)
        dt = time.perf_counter() - t0
        st = contour.last_stats
        print(f"K2 res={a.res}: host {1e3*dt:.1f} ms, kernels {st['kernel_ms']:.3f} ms = {st['work_units']/st['kernel_ms']/1e6:.0f} GB/s algorithmic, "
              f"{len(lines)} lines, {sum(len(l) for l in lines)} vertices", flush=True)

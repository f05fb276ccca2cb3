"""K1 over a whole grid in ONE device-resident launch (lm_escape_grid_f64_dev), a few times; the ncu target for the
roofline `traffic` figure of bench.py (the host-buffer API splits the grid into chunks)."""
import argparse, ctypes as C, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=32768); ap.add_argument("--max_iter", type=int, default=10000)
ap.add_argument("--reps", type=int, default=2); ap.add_argument("--f32", action="store_true")
a = ap.parse_args()
build.build(); _shim.set_device(0); torch.cuda.set_device(0)
dev = torch.device("cuda", 0); stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
xs = torch.from_numpy(np.linspace(-2.1, 0.9, a.res)).to(dev); ys = torch.from_numpy(np.linspace(-1.5, 1.5, a.res)).to(dev)
d = torch.empty((a.res, a.res), dtype=torch.int32, device=dev); w = torch.zeros(1, dtype=torch.int64, device=dev)
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if a.f32:
        _shim.call("lm_escape_grid_f32_dev", P(xs), a.res, P(ys), a.res, a.max_iter, 2.0, P(d), P(w), stream)
    else:
        _shim.call("lm_escape_grid_f64_dev", P(xs), a.res, P(ys), a.res, a.max_iter, 2.0, 0, P(d), None, None, P(w), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"res={a.res} mi={a.max_iter} {'f32' if a.f32 else 'f64'} one launch: {ms:.3f} ms, {int(w.item())/ms/1e6:.1f} Gpi/s", flush=True)

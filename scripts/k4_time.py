"""K4 stencils on device buffers, CUDA-event timed (8192^2 and 16384^2)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, device
for n in (8192, 16384):
    U = np.random.default_rng(0).standard_normal((n, n))
    a = device.DeviceBuffer(U.nbytes); b = device.DeviceBuffer(U.nbytes); a.upload(U)
    for name, args in (("lm_laplacian5_periodic_dev", (0.01,)), ("lm_smooth5_interior_dev", ())):
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3): _shim.call(name, C.c_void_p(a.ptr), n, n, *args, C.c_void_p(b.ptr), None)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(10): _shim.call(name, C.c_void_p(a.ptr), n, n, *args, C.c_void_p(b.ptr), None)
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        print(f"{name} {n}^2: {ms:.3f} ms = {n*n*16/ms/1e6:.0f} GB/s algorithmic (16 B/pixel)", flush=True)
    a.free(); b.free()

"""Stencil kernels on device buffers (ncu target)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
U = np.random.default_rng(0).standard_normal((n, n))
a = device.DeviceBuffer(U.nbytes); b = device.DeviceBuffer(U.nbytes); a.upload(U)
for _ in range(4):
    _shim.call("lm_laplacian5_periodic_dev", C.c_void_p(a.ptr), n, n, 0.01, C.c_void_p(b.ptr), None)
    _shim.call("lm_smooth5_interior_dev", C.c_void_p(a.ptr), n, n, C.c_void_p(b.ptr), None)
_shim.call("lm_device_synchronize")
print("ok")

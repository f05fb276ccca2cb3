import ctypes as C, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim
for variant in (0, 1):
    for w in (4, 8, 12, 16, 24, 32, 48, 64):
        g = C.c_double()
        _shim.call("lm_probe_k1_loop", variant, w, 20000, C.byref(g))
        print(f"variant={variant} warps/SM={w} Gpi/s={g.value:.1f}  (6-instr ceiling at 18.28 Tinstr/s = 3047)", flush=True)

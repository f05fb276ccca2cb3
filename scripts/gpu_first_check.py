"""Ad-hoc first GPU check: K1 dwell/potentials, K4 stencils, probes vs the CPU oracle."""
import sys, time, json, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, escape
from oracle import oracle

print("device", _shim.device_info(), flush=True)
print("missing", _shim.missing_exports())
ok = True
for res, mi in [(64, 100), (257, 500), (1000, 500), (2000, 500)]:
    xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res)
    t = time.time(); d_o, w_o = oracle.dwell_grid(xs, ys, mi); t_o = time.time() - t
    d, _, st = escape.escape_grid(xs, ys, mi)
    mism = int((d != d_o).sum())
    print(f"dwell res={res} mi={mi} mismatches={mism} work={st['work_units']} oracle_work={w_o} "
          f"kernel_ms={st['kernel_ms']:.3f} Gpi/s={st['work_units']/st['kernel_ms']/1e6:.1f} oracle_s={t_o:.2f}", flush=True)
    ok &= mism == 0 and st['work_units'] == w_o
    d64, _, _ = escape.escape_grid(xs, ys, mi, want_dwell="f64")
    ok &= bool((d64 == d_o).all())
# potentials
xs = np.linspace(-2, 2, 200); ys = np.linspace(-2, 2, 200)
for mode, R, mi in [(1, 2.0, 300), (2, 10.0, 300), (3, 2.0, 200), (4, 4.0, 500)]:
    d_o, f_o = oracle.potential_grid(xs, ys, mi, R, mode)
    d, f, st = escape.escape_grid(xs, ys, mi, R, mode)
    mism = int((d != d_o).sum())
    err = float(np.max(np.abs(f - f_o) / np.maximum(np.abs(f_o), 1e-300)))
    print(f"potential mode={mode} dwell mismatches={mism} max rel err={err:.3e}", flush=True)
    ok &= mism == 0 and err < 1e-13
# points
rng = np.random.default_rng(0)
pts = (rng.uniform(-2, 1, 5000) + 1j * rng.uniform(-1.5, 1.5, 5000))
g_o, it_o, phi_o = oracle.batch_potential(pts, 2000, 2.0)
g, it, phi = escape.batch_potential(pts, 2000, 2.0)
print("points it mismatches", int((it != it_o).sum()), "g err", float(np.max(np.abs(g - g_o))),
      "phi err", float(np.nanmax(np.abs(phi - phi_o))), "nan agree", bool((np.isnan(phi.real) == np.isnan(phi_o.real)).all()), flush=True)
ok &= bool((it == it_o).all())
# stencils
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import stencils
for shape in [(200, 200), (301, 257), (1024, 2048)]:
    U = rng.standard_normal(shape)
    L = stencils.laplacian(U, 0.02); L_o = oracle.laplacian(U, 0.02)
    S = stencils.smooth5(U); S_o = oracle.smooth5(U)
    print("stencil", shape, "lap equal", bool((L == L_o).all()), "smooth equal", bool((S == S_o).all()), flush=True)
    ok &= bool((L == L_o).all()) and bool((S == S_o).all())
# probes
a = C.c_double(); b = C.c_double()
_shim.call("lm_probe_fp64_peak", 4000, C.byref(a), C.byref(b))
print("fp64 peak TFLOP/s (DFMA)", a.value, "DMUL/DADD Tinstr/s", b.value, flush=True)
g = C.c_double()
_shim.call("lm_probe_hbm_copy", 1 << 30, 5, C.byref(g))
print("hbm copy GB/s", g.value, flush=True)
# bigger perf runs
for res, mi in [(4096, 2000), (8192, 2000)]:
    xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res)
    for rep in range(2):
        d, _, st = escape.escape_grid(xs, ys, mi)
        print(f"perf res={res} mi={mi} kernel_ms={st['kernel_ms']:.2f} work={st['work_units']:.4g} "
              f"Gpi/s={st['work_units']/st['kernel_ms']/1e6:.1f}", flush=True)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)

"""Host / device timeline of bench.py's device-resident step (K1, records, link) -- debugging aid."""
import ctypes as C, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np, torch
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build
build.build(); _shim.set_device(0); torch.cuda.set_device(0)
dev = torch.device("cuda", 0); lib = _shim.load()
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
res, mi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 10000)
xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res); level = 0.96 * mi
xs_d = torch.from_numpy(xs).to(dev); ys_d = torch.from_numpy(ys).to(dev)
dwell = torch.empty((res + 1, res), dtype=torch.int32, device=dev)
work = torch.zeros(1, dtype=torch.int64, device=dev)
rec = torch.empty((max(int(0.002 * res * res) + 4096, 1 << 16), 8), dtype=torch.int64, device=dev)
n = C.c_int64(0); nv = C.c_int64(0); nl = C.c_int64(0); first = np.zeros(1, dtype=np.int64)
P = lambda t: C.c_void_p(t.data_ptr())
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = T(); e[0].record()
    _shim.call("lm_escape_grid_f64_dev", P(xs_d), res, P(ys_d), res, mi, 2.0, 0, P(dwell), None, None, P(work), stream)
    h1 = time.perf_counter(); e[1].record(); t1 = T()
    rc = lib.lm_contour_records_dev(P(dwell), _shim.ptr(xs), res, _shim.ptr(ys), res, 0, float(level), P(rec), rec.shape[0], C.byref(n), stream)
    h2 = time.perf_counter(); e[2].record(); t2 = T()
    rc2 = lib.lm_contour_link_dev(P(rec), n.value, _shim.ptr(xs), res, _shim.ptr(ys), res, float(level), None, 0, C.byref(nv), _shim.ptr(first), 0, C.byref(nl), stream)
    h3 = time.perf_counter(); e[3].record(); t3 = T()
    w = int(work.item()); t4 = T()
    print(f"rep {rep}: K1 host-call {1e3*(h1-t0):.2f} ms, synced {1e3*(t1-t0):.2f} (events {e[0].elapsed_time(e[1]):.2f}) | records call {1e3*(h2-t1):.2f}, synced {1e3*(t2-t1):.2f} "
          f"(events {e[1].elapsed_time(e[2]):.2f}) rc={rc} n={n.value} | link call {1e3*(h3-t2):.2f}, synced {1e3*(t3-t2):.2f} (events {e[2].elapsed_time(e[3]):.2f}) rc={rc2} | item {1e3*(t4-t3):.2f}", flush=True)
# the same back to back without the intermediate syncs
for rep in range(3):
    t0 = T()
    _shim.call("lm_escape_grid_f64_dev", P(xs_d), res, P(ys_d), res, mi, 2.0, 0, P(dwell), None, None, P(work), stream)
    a = time.perf_counter()
    lib.lm_contour_records_dev(P(dwell), _shim.ptr(xs), res, _shim.ptr(ys), res, 0, float(level), P(rec), rec.shape[0], C.byref(n), stream)
    b = time.perf_counter()
    lib.lm_contour_link_dev(P(rec), n.value, _shim.ptr(xs), res, _shim.ptr(ys), res, float(level), None, 0, C.byref(nv), _shim.ptr(first), 0, C.byref(nl), stream)
    c = time.perf_counter(); w = int(work.item()); t1 = T()
    print(f"back to back: total {1e3*(t1-t0):.2f} ms; K1 call returned after {1e3*(a-t0):.2f}, records after {1e3*(b-t0):.2f}, link after {1e3*(c-t0):.2f}", flush=True)

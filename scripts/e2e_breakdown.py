"""Where does the host-buffer path spend its time?  (ad-hoc profiling aid)"""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, contour, escape
res, mi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 10000)
xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res)
out = _shim.pinned_empty((res, res), np.int32)
st = _shim.Stats()
for rep in range(3):
    t0 = time.perf_counter()
    _shim.call("lm_escape_grid_f64", _shim.ptr(xs), res, _shim.ptr(ys), res, mi, 2.0, 0, _shim.ptr(out), None, None, C.byref(st))
    t1 = time.perf_counter()
    lines = contour.contour_lines(xs, ys, out, 0.96 * mi)
    t2 = time.perf_counter()
    print(f"K1 host call {1e3*(t1-t0):.1f} ms (kernel {st.kernel_ms:.1f} ms)  K2 host call {1e3*(t2-t1):.1f} ms (kernels {contour.last_stats['kernel_ms']:.2f} ms) lines={len(lines)}", flush=True)
# K2 pieces
d_dev = _shim.load().lm_dev_alloc(out.nbytes)
t0 = time.perf_counter(); _shim.call("lm_memcpy_h2d", C.c_void_p(d_dev), _shim.ptr(out), out.nbytes, None); _shim.call("lm_stream_synchronize", None); t1 = time.perf_counter()
print(f"H2D {out.nbytes/1e9:.2f} GB pinned: {1e3*(t1-t0):.1f} ms = {out.nbytes/1e9/(t1-t0):.1f} GB/s")
t0 = time.perf_counter(); lines = contour.contour_lines_dev(d_dev, xs, ys, 0.96 * mi); t1 = time.perf_counter()
print(f"K2 from device: {1e3*(t1-t0):.1f} ms (kernels {contour.last_stats['kernel_ms']:.2f} ms), {sum(len(l) for l in lines)} vertices in {len(lines)} lines")
recs = np.empty((2_000_000, 8), dtype=np.int64); n = C.c_int64(0)
t0 = time.perf_counter(); _shim.call("lm_contour_classify_dev", C.c_void_p(d_dev), _shim.ptr(xs), res, _shim.ptr(ys), res, 0, 0.96 * mi, _shim.ptr(recs), recs.shape[0], C.byref(n), None); t1 = time.perf_counter()
print(f"classify (kernels + record D2H): {1e3*(t1-t0):.1f} ms, {n.value} records")
t0 = time.perf_counter(); l2 = contour.link_records(recs[:n.value], xs, ys, 0.96 * mi); t1 = time.perf_counter()
print(f"link (upload 64 B/record + device link + lines D2H): {1e3*(t1-t0):.1f} ms")
out2 = _shim.pinned_empty((res, res), np.int32)
t0 = time.perf_counter(); _shim.call("lm_memcpy_d2h", _shim.ptr(out2), C.c_void_p(d_dev), out.nbytes, None); _shim.call("lm_stream_synchronize", None); t1 = time.perf_counter()
print(f"D2H {out.nbytes/1e9:.2f} GB pinned: {1e3*(t1-t0):.1f} ms = {out.nbytes/1e9/(t1-t0):.1f} GB/s")
# the device linker alone: records resident, lines into page-locked buffers
drec = _shim.load().lm_dev_alloc(n.value * 64)
_shim.call("lm_memcpy_h2d", C.c_void_p(drec), _shim.ptr(recs), n.value * 64, None); _shim.call("lm_stream_synchronize", None)
verts = _shim.pinned_empty((4 * n.value + 16, 2), np.float64); offs = _shim.pinned_empty(2 * n.value + 17, np.int64); nv = C.c_int64(0); nl = C.c_int64(0)
for rep in range(3):
    t0 = time.perf_counter()
    _shim.call("lm_contour_link_dev", C.c_void_p(drec), n.value, _shim.ptr(xs), res, _shim.ptr(ys), res, 0.96 * mi, _shim.ptr(verts), verts.shape[0], C.byref(nv),
               _shim.ptr(offs), offs.size - 1, C.byref(nl), None)
    t1 = time.perf_counter()
    print(f"lm_contour_link_dev (records resident, {nv.value} vertices / {nl.value} lines to pinned host): {1e3*(t1-t0):.2f} ms")

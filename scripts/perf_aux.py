"""Throughput of the auxiliary kernels (K3 roots, K4 stencils, K4a log-potential, K1d points)."""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, device, escape, lucas, potentials, stencils

def cfg5_batch(npoly, seed=0, maxdeg=25):
    rng = np.random.default_rng(seed)
    deg = rng.integers(2, maxdeg + 1, size=npoly).astype(np.int32)
    top = rng.integers(0, 3, size=(npoly, maxdeg)).astype(np.float64)
    col = np.arange(maxdeg)[None, :]
    top[col >= deg[:, None]] = 0.0
    last = top[np.arange(npoly), deg - 1]
    top[np.arange(npoly), deg - 1] = np.where(last == 0, 1.0, last)
    return top, deg

npoly = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
top, deg = cfg5_batch(npoly)
for rep in range(2):
    t0 = time.perf_counter()
    vals, kept, iters = lucas.roots_batched(top, deg, sort=False)
    dt = time.perf_counter() - t0
    st = lucas.last_stats
    print(f"K3 roots: {npoly} polys, {st['work_units']} roots, kernel {st['kernel_ms']:.1f} ms = {st['work_units']/st['kernel_ms']/1e3:.1f} M roots/s "
          f"({npoly/st['kernel_ms']/1e3:.2f} M polys/s), host call {dt*1e3:.0f} ms, mean sweeps {iters.mean():.2f}, max {iters.max()}, failed {(iters<0).sum()}", flush=True)
# numpy baseline on a subsample (1 core)
sub = 20000
t0 = time.perf_counter()
for d in range(2, 26):
    sel = np.where(deg[:sub] == d)[0]
    if sel.size == 0: continue
    M = np.zeros((sel.size, d, d)); M[:, 0, :] = top[sel, :d]
    idx = np.arange(1, d); M[:, idx, idx - 1] = 1.0
    np.linalg.eigvals(M)
dt = time.perf_counter() - t0
print(f"numpy eigvals (stacked, 1 process): {deg[:sub].sum()/dt/1e6:.3f} M roots/s ({sub/dt/1e3:.1f} k polys/s)")
# lucas family (all ones) high degree
for n in (300, 1220):
    t0 = time.perf_counter(); pts = lucas.construct_points(range(20, n + 1, 20)); dt = time.perf_counter() - t0
    print(f"construct_points(range(20,{n+1},20)): {len(pts)} pts in {dt*1e3:.1f} ms (kernel {lucas.last_stats['kernel_ms']:.1f} ms)")
# stencils on device buffers
n = 8192
U = np.random.default_rng(0).standard_normal((n, n))
a = device.DeviceBuffer(U.nbytes); b = device.DeviceBuffer(U.nbytes); a.upload(U)
import torch
for name, args in (("lm_laplacian5_periodic_dev", (0.01,)), ("lm_smooth5_interior_dev", ())):
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3): _shim.call(name, C.c_void_p(a.ptr), n, n, *args, C.c_void_p(b.ptr), None)
    torch.cuda.synchronize(); ev0.record()
    for _ in range(10): _shim.call(name, C.c_void_p(a.ptr), n, n, *args, C.c_void_p(b.ptr), None)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f"{name} 8192^2: {ms:.3f} ms = {n*n*16/ms/1e6:.0f} GB/s algorithmic (16 B/pixel)")
# log potential
pts = np.random.default_rng(1).uniform(-1.2, 1.2, (20000, 2)); g = np.linspace(-2, 2, 400)
potentials.log_potential(pts, g, g); st = potentials.last_stats
print(f"log_potential 400^2 x 20000 pts: kernel {st['kernel_ms']:.1f} ms = {st['work_units']/st['kernel_ms']/1e6:.1f} G pairs/s")
# points potential
c = lucas.compute_inverse_eigenvalues(2, 200)
t0 = time.perf_counter(); gg, it, phi = escape.batch_potential(c, 20000, 2.0); dt = time.perf_counter() - t0
st = escape.last_stats
print(f"batch_potential {c.size} pts max_iter 20000: kernel {st['kernel_ms']:.2f} ms = {st['work_units']/st['kernel_ms']/1e6:.1f} Gpi/s, host {dt*1e3:.1f} ms")
# log potential at larger point counts, all variants
for npts, variant, eps in ((2_000_000, 0, 1e-12), (2_000_000, 2, 1e-12), (2_000_000, 3, 1e-6)):
    pts = np.random.default_rng(1).uniform(-1.8, 1.8, (npts, 2))
    potentials._logpot(pts[:, 0], pts[:, 1], g, g, eps, variant); st = potentials.last_stats
    print(f"log_potential 400^2 x {npts} pts variant {variant} eps {eps:g}: kernel {st['kernel_ms']:.1f} ms = "
          f"{st['work_units']/st['kernel_ms']/1e6:.1f} G pairs/s")
# the fused cloud stage (config 5) on the same batch
for rep in range(2):
    t0 = time.perf_counter()
    out = lucas.cloud_fields(top, deg, g, g, potential=(20000, 2.0))
    dt = time.perf_counter() - t0
    st = out["stats"]
    print(f"cloud_fields {npoly} polys: {st['n_points']} pts, host {dt*1e3:.0f} ms; roots {st['roots_ms']:.1f} ms "
          f"({st['n_roots']/st['roots_ms']/1e3:.0f} M roots/s), compact {st['compact_ms']:.2f} ms, "
          f"potential {st['potential_ms']:.1f} ms ({st['potential_work']/max(st['potential_ms'],1e-9)/1e6:.0f} Gpi/s, "
          f"inside {(out['it']==20000).mean():.3f}), logpot {st['logpot_ms']:.1f} ms ({st['pairs']/st['logpot_ms']/1e6:.0f} G pairs/s), "
          f"stencil {st['stencil_ms']:.3f} ms", flush=True)

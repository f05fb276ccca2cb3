"""K3 alone on a config-5 shaped batch (ncu target)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import lucas
npoly = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
deg = rng.integers(2, 26, size=npoly).astype(np.int32)
top = rng.integers(0, 3, size=(npoly, 25)).astype(np.float64)
top[np.arange(25)[None, :] >= deg[:, None]] = 0.0
last = top[np.arange(npoly), deg - 1]
top[np.arange(npoly), deg - 1] = np.where(last == 0, 1.0, last)
for rep in range(2):
    vals, kept, iters = lucas.roots_batched(top, deg, invert=True, tol=1e-12, sort=False)
    st = lucas.last_stats
    print(f"K3: {npoly} polys {st['work_units']} roots kernel {st['kernel_ms']:.2f} ms = {st['work_units']/st['kernel_ms']/1e3:.0f} M roots/s, "
          f"sweeps mean {iters.mean():.2f}", flush=True)
# accuracy of the last run: relative residual of every root (in 1/lambda convention -> back to lambda), and the
# multiset distance to np.linalg.eigvals on a sample
lam = 1.0 / vals
worst = 0.0
for k in range(0, npoly, max(npoly // 20000, 1)):
    d = int(deg[k]); z = lam[k, :kept[k]]
    c = np.concatenate([[1.0], -top[k, :d]])
    res = np.abs(np.polyval(c, z)) / np.polyval(np.abs(c), np.abs(z))
    worst = max(worst, float(res.max()))
print(f"max relative residual |p(z)| / p~(|z|) over the sample: {worst:.3e}")
import time
t0 = time.time(); bad = 0; worst_d = 0.0
for k in range(0, npoly, max(npoly // 3000, 1)):
    d = int(deg[k])
    M = np.zeros((d, d)); M[0, :] = top[k, :d]; M[np.arange(1, d), np.arange(d - 1)] = 1.0
    ev = np.linalg.eigvals(M); ev = ev[np.abs(ev) > 1e-12]
    a = np.sort_complex(1.0 / ev); b = np.sort_complex(vals[k, :kept[k]])
    if a.size != b.size:
        bad += 1; continue
    # greedy matching
    used = np.zeros(b.size, bool); w = 0.0
    for v in a:
        dist = np.abs(b - v); dist[used] = np.inf; j = int(np.argmin(dist)); used[j] = True
        w = max(w, dist[j] / abs(v))
    worst_d = max(worst_d, w)
    if w > 1e-10: bad += 1
print(f"vs numpy eigvals on a sample: worst relative distance {worst_d:.3e}, polynomials beyond 1e-10: {bad}")

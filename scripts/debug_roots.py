import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import lucas
npoly, maxdeg = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 25
rng = np.random.default_rng(0)
deg = rng.integers(2, 26, size=npoly).astype(np.int32)
top = rng.integers(0, 3, size=(npoly, maxdeg)).astype(np.float64)
top[np.arange(maxdeg)[None, :] >= deg[:, None]] = 0.0
last = top[np.arange(npoly), deg - 1]
top[np.arange(npoly), deg - 1] = np.where(last == 0, 1.0, last)
vals, kept, iters = lucas.roots_batched(top, deg, invert=True, tol=1e-12, sort=False)
bad = np.where(kept != deg)[0]
print("bad", bad.size, "of", npoly)
for b in bad[:6]:
    d = deg[b]
    print("poly", b, "deg", d, "top", top[b, :d], "kept", kept[b], "iters", iters[b])
    v2, k2, i2 = lucas.roots_batched(top[b:b+1], deg[b:b+1], invert=False, sort=False)
    print("  roots", v2[0, :d])
    print("  numpy", np.linalg.eigvals(np.vstack([top[b, :d], np.eye(d)[:-1]])))

"""CPU-only timing of the host linker (lm_contour_link) on records built from the oracle's dwell grid."""
import sys, time, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from oracle import oracle
from helpers import records_from_dwell
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import contour

res = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mi = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cache = Path(f"/tmp/link_records_{res}_{mi}.npz")
xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res)
lvl = 0.96 * mi
if cache.exists():
    recs = np.load(cache)["recs"]
else:
    t0 = time.time(); d, _ = oracle.dwell_grid(xs, ys, mi); print("dwell", time.time() - t0)
    above = d > lvl
    cross = ~((above[:-1, :-1] == above[:-1, 1:]) & (above[:-1, :-1] == above[1:, :-1]) & (above[:-1, :-1] == above[1:, 1:]))
    jj, ii = np.nonzero(cross)
    print("crossing quads", jj.size)
    t0 = time.time()
    parts = []
    for j, i in zip(jj, ii):                      # one 2x2 block per crossing quad through the reference restatement
        r = records_from_dwell(d[j:j + 2, i:i + 2], xs[i:i + 2], ys[j:j + 2], lvl)
        r[:, 0] = j * res + i
        parts.append(r)
    recs = np.concatenate(parts)
    print("records", recs.shape, time.time() - t0)
    np.savez(cache, recs=recs)
os.environ["LM_LINK_DEBUG"] = "1"
for _ in range(4):
    t0 = time.perf_counter(); lines = contour.link_records(recs, xs, ys, lvl); dt = time.perf_counter() - t0
    print(f"link {recs.shape[0]} records -> {len(lines)} lines, {int(lines.lengths().sum())} vertices: {dt*1e3:.1f} ms")

"""Load-balance study on ONE GPU (K1 blocks are independent, so the per-rank times of an N-GPU run are just the times
of the N row blocks run one after another).

  1. full K1 pass -> exact per-row features of the workload: pixel-iterations, pixels, careful-phase iterations
     sum(min(it, 64)), late escapers count(escaped and it > 64)
  2. K1 timed (CUDA events) on B bands of rows cut at equal iterations -> least-squares cost model
        t = a*iters + b*pixels + c*min(it,64) + d*late
  3. for N in 2, 4, 8: cuts from (i) the iteration-only coarse profile (round 1), (ii) the coarse profile weighted
     by a candidate model; per-block K1 times -> efficiency mean/max.
Writes gpurun_out/balance_<cfg>.json.   usage: balance_probe.py cfg3|cfg4|cfg2
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from bench import WORKLOADS  # noqa: E402
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build, sharding  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w = WORKLOADS[cfg]
build.build(); _shim.set_device(0); torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
res, mi = w["res"], w["max_iter"]
xs = np.linspace(*w["xlim"], res); ys = np.linspace(*w["ylim"], res)
xs_d = torch.from_numpy(xs).to(dev); ys_d = torch.from_numpy(ys).to(dev)
dwell = torch.empty((res, res), dtype=torch.int32, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())


def k1(r0, r1, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _shim.call("lm_escape_grid_f64_dev", P(xs_d), res, C.c_void_p(ys_d.data_ptr() + 8 * r0), r1 - r0, mi, 2.0, 0,
                   C.c_void_p(dwell.data_ptr() + 4 * res * r0), None, None, None, stream)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


t_full = k1(0, res, reps=2)
it = torch.clamp(dwell.to(torch.int64) + 1, max=mi)
esc = dwell < mi
F = torch.stack([it.sum(1), torch.full((res,), res, device=dev), torch.clamp(it, max=64).sum(1),
                 (esc & (it > 64)).sum(1)], 1).double().cpu().numpy()            # [rows, 4]
del it, esc
out = {"cfg": cfg, "t_full_ms": t_full, "total_iters": float(F[:, 0].sum())}

B = 48
cuts = sharding.balanced_row_cuts(F[:, 0], B)
tb = np.array([k1(a, b) for a, b in zip(cuts[:-1], cuts[1:])])
Fb = np.array([F[a:b].sum(0) for a, b in zip(cuts[:-1], cuts[1:])])
out["bands"] = {"cuts": cuts, "ms": tb.tolist(), "features": Fb.tolist()}
models = {}
for name, cols in (("iters", [0]), ("iters+pixels", [0, 1]), ("iters+pixels+careful", [0, 1, 2]), ("all4", [0, 1, 2, 3])):
    coef, *_ = np.linalg.lstsq(Fb[:, cols], tb, rcond=None)
    pred = Fb[:, cols] @ coef
    models[name] = {"cols": cols, "coef": coef.tolist(), "max_rel_err": float(np.abs(pred / tb - 1).max()),
                    "rms_rel_err": float(np.sqrt(np.mean((pred / tb - 1) ** 2)))}
out["models"] = models
print(json.dumps(models, indent=1), flush=True)

# the product's one-shot plan (coarse pre-pass + the committed cost model) against round 1's iteration-only profile
effs = {}
import time
for N in (2, 4, 8):
    row = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan = sharding.plan_row_cuts(xs, ys, mi, N, device=dev)
    torch.cuda.synchronize(); setup_ms = 1e3 * (time.perf_counter() - t0)
    prof_it = sharding.coarse_row_profile(xs, ys, mi, iterations_only=True)
    for name, c in (("iterations only (round 1)", sharding.balanced_row_cuts(prof_it, N)), ("cost model (one shot)", plan["cuts"])):
        t = np.array([k1(a, b, reps=2) for a, b in zip(c[:-1], c[1:])])
        row[name] = {"cuts": c, "ms": t.tolist(), "efficiency": float(t.mean() / t.max()), "sum_over_full": float(t.sum() / t_full)}
        print(N, name, f"eff={t.mean() / t.max():.4f}", [round(x, 2) for x in t], flush=True)
    row["setup_ms"] = setup_ms
    row["model"] = plan["model"]
    effs[str(N)] = row
out["efficiency"] = effs
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"balance_{cfg}.json").write_text(json.dumps(out))

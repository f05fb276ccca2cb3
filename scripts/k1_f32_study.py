"""fp32 variant of K1 against the fp64 kernel: dwell mismatch / interior-mask agreement on the windows of BASELINE.json
configs 1-4, and device-resident throughput at full size (CUDA events)."""
import ctypes as C, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import numpy as np, torch
from bench import WORKLOADS
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build
build.build(); _shim.set_device(0); torch.cuda.set_device(0)
dev = torch.device("cuda", 0); stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
out = {}
for name, w in WORKLOADS.items():
    for res in ((1024, w["res"]) if name != "cfg4" else (1024, 4096)):
        mi = w["max_iter"]
        xs = torch.from_numpy(np.linspace(*w["xlim"], res)).to(dev); ys = torch.from_numpy(np.linspace(*w["ylim"], res)).to(dev)
        d64 = torch.empty((res, res), dtype=torch.int32, device=dev); d32 = torch.empty_like(d64)
        w64 = torch.zeros(1, dtype=torch.int64, device=dev); w32 = torch.zeros(1, dtype=torch.int64, device=dev)
        t = {}
        for tag, fn, dst, wk in (("f64", "lm_escape_grid_f64_dev", d64, w64), ("f32", "lm_escape_grid_f32_dev", d32, w32)):
            best = 1e30
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if tag == "f64":
                    _shim.call(fn, P(xs), res, P(ys), res, mi, 2.0, 0, P(dst), None, None, P(wk), stream)
                else:
                    _shim.call(fn, P(xs), res, P(ys), res, mi, 2.0, P(dst), P(wk), stream)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            t[tag] = best
        mism = float((d64 != d32).double().mean())
        inside = float(((d64 == mi) == (d32 == mi)).double().mean())
        esc = (d64 < mi) & (d32 < mi)
        rel = float(((d64[esc] - d32[esc]).abs().double() / (d64[esc].double() + 1)).mean()) if esc.any() else 0.0
        big = float((((d64 - d32).abs() > 1) ).double().mean())
        row = {"res": res, "max_iter": mi, "dwell_mismatch_frac": mism, "mismatch_by_more_than_1_frac": big, "interior_mask_agreement": inside,
               "mean_rel_dwell_diff_of_escaped": rel, "f64_ms": t["f64"], "f32_ms": t["f32"],
               "f64_gpi": int(w64.item()) / t["f64"] / 1e6, "f32_gpi": int(w32.item()) / t["f32"] / 1e6}
        out[f"{name}@{res}"] = row
        print(name, json.dumps(row), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "k1_f32_study.json").write_text(json.dumps(out, indent=1))

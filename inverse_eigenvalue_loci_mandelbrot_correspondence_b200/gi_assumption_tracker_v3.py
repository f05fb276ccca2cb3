#!/usr/bin/env python3
"""Drop-in for the reference's `gi_assumption_tracker_v3.py` (the script behind Table A.1): same command line, same
`<out-prefix>.csv` / `<out-prefix>.json`, with every stage of a level on the B200.

    python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.gi_assumption_tracker_v3 \
        --T-fixed 25 --sigma-bins 3 --construct-max-growth 1.6 --mandelbrot-samples-growth 1.6 \
        --mandelbrot-samples-max 300000 --bins-max 512 --out-prefix v3_T25_sigma3_dense

reproduces the reference's published `v3_T25_sigma3_dense.csv` (all columns but `runtime_sec`).

What the reference does per level (gi_assumption_tracker_v3.py:205-299) and where it runs here:
  mod.construct_points(ns)            Lucas loci, n = step..construct_max_n          K3   (lm_roots_batched)
  mod.sample_mandelbrot_boundary()    DE grid + quantile mask + np.random.choice     K1b  (lm_distance_grid_f64, numpy-FMA recipe)
  mod.entropic_ot_alignment(C, M)     nearest-neighbour matching                          (lm_nearest_match)
  mod.procrustes_align_no_scale       2x2 SVD                                         host, as in the stock module
  mollified_histogram x 2, KL, TV, overlap, GI flow                                       (lm_mollified_histogram, lm_density_compare, lm_gi_flow)
then the level schedule (:296-299).  `--module` defaults to the package's GPU module; any module with the stock contract
works (its generators then run wherever that module runs them), the density stage always runs on the device.
"""
from __future__ import annotations

import argparse
import csv
import importlib.util
import json
import math
import sys
import time
from pathlib import Path

import numpy as np

_ROOT = Path(__file__).resolve().parents[1]
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

DEFAULT_MODULE = Path(__file__).resolve().parent / "tci_construct_mandelbrot_b200.py"

# column order of the reference's Row dataclass (gi_assumption_tracker_v3.py:48-80) = header of the CSV
COLUMNS = ("bins", "mesh_proxy", "construct_max_n", "construct_step", "n_construct_pts", "mandelbrot_grid", "mandelbrot_samples",
           "n_mandel_pts", "alpha", "sigma_bins", "mode", "T_n", "kl_initial", "delta_n", "kl_PM_PC", "pinsker_tv_bound_XT_PM",
           "tv_XT_PM", "tv_PC_PM", "overlap_mass_PC_PM", "mass_outside_domain_C", "mass_outside_domain_M", "tv_bound_PC_PM",
           "compound", "compound_with_pinsker", "stop_reason", "runtime_sec")


def load_module(module_path, module_name: str = "tci_fixed_import"):
    spec = importlib.util.spec_from_file_location(module_name, str(module_path))
    if spec is None or spec.loader is None:
        raise RuntimeError(f"Unable to load module at {module_path}")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# (flag, type, default) of the reference's command line (gi_assumption_tracker_v3.py:154-186); dest = flag with "-" -> "_"
CLI_OPTIONS = (
    ("seed", int, 7), ("domain", str, "-2.2:1.2:-1.6:1.6"), ("alpha", float, 0.1),
    ("bins-start", int, 64), ("bins-max", int, 1024),
    ("construct-step", int, 20), ("construct-max-start", int, 300), ("construct-max-growth", float, 1.35),
    ("mandelbrot-grid-start", int, 600), ("mandelbrot-grid-growth", float, 1.15),
    ("mandelbrot-samples-start", int, 25000), ("mandelbrot-samples-growth", float, 1.35), ("mandelbrot-samples-max", int, 150000),
    ("sigma-bins", float, 1.0), ("T-fixed", int, -1), ("kl-threshold", float, 1e-6), ("max-steps", int, 800), ("min-steps", int, 5),
    ("compound-threshold", float, 1e-3), ("tv-threshold", float, 0.05), ("out-prefix", str, "gi_assumptions_v3"),
)
CLI_HELP = {"domain": "xmin:xmax:ymin:ymax (negative first value: write --domain=...)",
            "sigma-bins": "blur width in histogram cells; 0 keeps the raw histogram",
            "T-fixed": "a positive value runs exactly that many GI sweeps instead of stopping on --kl-threshold"}


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--module", default=str(DEFAULT_MODULE), help="script with the tci_construct_mandelbrot_v002_fixed.py contract")
    for flag, kind, default in CLI_OPTIONS:
        ap.add_argument("--" + flag, type=kind, default=default, help=CLI_HELP.get(flag))
    return ap


def _is_stock_KL(mod) -> bool:
    """Does mod.KL compute the stock module's formula (tci_construct_mandelbrot_v002_fixed.py:84-86),
    sum(P_ * (log P_ - log X_)) with P_, X_ clipped at mod.eps?  Checked on a probe pair that exercises the clip."""
    eps = float(getattr(mod, "eps", 1e-12))
    rng = np.random.default_rng(12345)
    P = rng.random((7, 5)); X = rng.random((7, 5))
    P[0, 0] = 0.0; X[1, 1] = 0.0; P[2, 2] = eps / 3; X[3, 3] = eps / 7
    P /= P.sum(); X /= X.sum()
    P_ = np.clip(P, eps, None); X_ = np.clip(X, eps, None)
    want = float(np.sum(P_ * (np.log(P_) - np.log(X_))))
    try:
        got = float(mod.KL(P, X))
    except Exception:
        return False
    return bool(np.isfinite(got) and abs(got - want) <= 1e-12 * max(1.0, abs(want)))


def _device_ops(mod):
    """the density-stage functions, bound to the device implementation; KL_fn carries the module's eps.
    The device GI flow evaluates the STOCK KL itself, so a plug-in module with a different KL (other clipping, base or
    symmetrisation) is refused instead of being silently replaced (the reference calls mod.KL everywhere, :214-224)."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import tracker
    if getattr(mod.KL, "lm_eps", None) is not None:
        return tracker, mod.KL
    if not _is_stock_KL(mod):
        raise SystemExit(f"--module {getattr(mod, '__file__', mod)}: its KL() is not the stock formula "
                         "sum(P_*(log P_ - log X_)) with P_, X_ = clip(., eps); the device GI flow cannot stand in for it. "
                         "Use a module whose KL comes from tracker.make_KL(eps) or the stock definition.")
    return tracker, tracker.make_KL(float(getattr(mod, "eps", 1e-12)))


def run_level(mod, ops, KL_fn, args, domain, bins: int, construct_max_n: int, mandel_grid: int, mandel_samples: int) -> dict:
    """One row of the table: generators -> alignment -> densities -> GI flow -> bounds (:205-291)."""
    t0 = time.time()
    mod.mandelbrot_grid = int(mandel_grid)
    mod.mandelbrot_samples = int(mandel_samples)
    step = int(args.construct_step)
    C = mod.construct_points(list(range(step, int(construct_max_n) + 1, step)))
    M = mod.sample_mandelbrot_boundary()
    M_aligned, C_sub = mod.entropic_ot_alignment(C, M)
    C_aligned = mod.procrustes_align_no_scale(C_sub, M_aligned)

    P_M = ops.mollified_histogram(mod, M_aligned, bins=bins, sigma_bins=float(args.sigma_bins))
    P_C = ops.mollified_histogram(mod, C_aligned, bins=bins, sigma_bins=float(args.sigma_bins))
    kl_PM_PC = float(KL_fn(P_M, P_C))
    alpha = float(args.alpha)
    if int(args.T_fixed) > 0:
        mode = f"fixedT={int(args.T_fixed)}"
        X_T, T_n, kl0, delta = ops.gi_flow_fixed_T(KL_fn, P_M, P_C, alpha, int(args.T_fixed))
        stop_reason = "fixed_T"
    else:
        mode = "adaptive"
        X_T, T_n, kl0, delta = ops.gi_flow_to_threshold(KL_fn, P_M, P_C, alpha, float(args.kl_threshold), int(args.max_steps),
                                                        int(args.min_steps))
        stop_reason = "kl_threshold_met" if delta <= float(args.kl_threshold) else "max_steps_reached"
    pinsker = math.sqrt(0.5 * float(delta))
    factor = (1.0 - alpha) ** (-int(T_n)) if int(T_n) > 0 else float("inf")
    values = (int(bins), 1.0 / float(bins), int(construct_max_n), step, int(C_aligned.size), int(mandel_grid), int(mandel_samples),
              int(M_aligned.size), alpha, float(args.sigma_bins), mode, int(T_n), float(kl0), float(delta), kl_PM_PC, float(pinsker),
              float(ops.tv_distance(X_T, P_M)), float(ops.tv_distance(P_C, P_M)), float(ops.overlap_mass(P_C, P_M)),
              float(ops.fraction_outside_domain(C_aligned, domain)), float(ops.fraction_outside_domain(M_aligned, domain)),
              float(factor * pinsker), float(factor * math.sqrt(float(delta))), float(factor * pinsker), stop_reason,
              float(time.time() - t0))
    return dict(zip(COLUMNS, values))


def run(args, mod=None, ops=None, KL_fn=None, log=print) -> tuple[list[dict], str]:
    """The level loop of main() (:193-299).  mod / ops / KL_fn default to the module named by args and the device ops."""
    np.random.seed(int(args.seed))
    domain = tuple(float(v) for v in args.domain.split(":"))
    if mod is None:
        mod = load_module(args.module)
    mod.domain = domain
    if ops is None:
        ops, KL_dev = _device_ops(mod)
        KL_fn = KL_fn or KL_dev
    elif KL_fn is None:
        KL_fn = mod.KL
    rows: list[dict] = []
    bins = int(args.bins_start)
    construct_max_n = int(args.construct_max_start)
    mandel_grid = int(args.mandelbrot_grid_start)
    mandel_samples = int(args.mandelbrot_samples_start)
    reason = ""
    step = int(args.construct_step)
    while bins <= int(args.bins_max):
        row = run_level(mod, ops, KL_fn, args, domain, bins, construct_max_n, mandel_grid, mandel_samples)
        rows.append(row)
        log(f"[{row['mode']} bins={bins}] delta_n={row['delta_n']:.3e}  Tn={row['T_n']}  TV(PC,PM)={row['tv_PC_PM']:.3e}  "
            f"overlap={row['overlap_mass_PC_PM']:.3e}  KL(PM||PC)={row['kl_PM_PC']:.3e}  compound={row['compound']:.3e}  "
            f"({row['runtime_sec']:.2f} s)")
        if (row["delta_n"] <= float(args.kl_threshold) and row["compound"] <= float(args.compound_threshold)
                and row["tv_PC_PM"] <= float(args.tv_threshold)):
            reason = "global_stop: kl<=threshold AND compound<=threshold AND TV(P_C,P_M)<=tv_threshold"
            break
        bins *= 2
        construct_max_n = int(round((construct_max_n * float(args.construct_max_growth)) / step)) * step
        mandel_grid = int(round(mandel_grid * float(args.mandelbrot_grid_growth)))
        mandel_samples = min(int(args.mandelbrot_samples_max), int(round(mandel_samples * float(args.mandelbrot_samples_growth))))
    return rows, reason


def write_outputs(args, rows: list[dict], reason: str) -> tuple[str, str]:
    csv_path, json_path = f"{args.out_prefix}.csv", f"{args.out_prefix}.json"
    with open(csv_path, "w", newline="", encoding="utf-8") as f:
        if rows:
            w = csv.DictWriter(f, fieldnames=list(COLUMNS))
            w.writeheader()
            w.writerows(rows)
    # the JSON header of the reference: its settings in its order (:305-330), then the rows
    order = ("module", "seed", "domain", "alpha", "sigma_bins", "bins_start", "bins_max", "T_fixed", "kl_threshold", "max_steps",
             "min_steps", "compound_threshold", "tv_threshold", "construct_step", "construct_max_start", "construct_max_growth",
             "mandelbrot_grid_start", "mandelbrot_grid_growth", "mandelbrot_samples_start", "mandelbrot_samples_growth",
             "mandelbrot_samples_max")
    kinds = {flag.replace("-", "_"): kind for flag, kind, _ in CLI_OPTIONS}
    meta = {}
    for key in order:
        value = getattr(args, key)
        meta[key] = tuple(float(v) for v in value.split(":")) if key == "domain" else kinds.get(key, str)(value)
    meta["global_stop_reason"] = reason
    meta["rows"] = rows
    with open(json_path, "w", encoding="utf-8") as f:
        json.dump(meta, f, indent=2)
    return csv_path, json_path


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    rows, reason = run(args)
    csv_path, json_path = write_outputs(args, rows, reason)
    print(f"\nWrote:\n  {csv_path}\n  {json_path}")
    if reason:
        print(f"Stopped early: {reason}")


if __name__ == "__main__":
    main()

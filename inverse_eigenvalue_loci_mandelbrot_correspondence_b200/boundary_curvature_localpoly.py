#!/usr/bin/env python3
"""GPU drop-in for the reference script boundary_curvature_localpoly.py (README Step 3; same CLI, same outputs).

    python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.boundary_curvature_localpoly \\
        --input_csv outputs/mandel_boundary.csv --output_prefix outputs/mandel_curv_localpoly --neighbors 7 --closed True

Outputs (boundary_curvature_localpoly.py:186-241):
  <prefix>_curvature.csv          idx,x,y,curvature,kappa_signed,speed,xprime,yprime,x2,y2 ("%.10g")
  <prefix>_curvature_hist.png     histogram of kappa (64 bins)
  <prefix>_curvature_overlay.png  the curve coloured by kappa
  <prefix>_summary.txt            n / mean / median / std / q05 / q95 / max

The per-point quadratic fits run on the device (curvature.compute_curvature_localpoly -> lm_curvature_localpoly).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .curvature import compute_curvature_localpoly, load_points  # noqa: F401  (reference-compatible names)


def save_csv(prefix: str, P, kappa, kappa_s, speed, aux) -> str:
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    out_csv = f"{prefix}_curvature.csv"
    cols = [np.arange(P.shape[0]), P[:, 0], P[:, 1], kappa, kappa_s, speed, aux["xprime"], aux["yprime"], aux["x2"], aux["y2"]]
    np.savetxt(out_csv, np.column_stack(cols), delimiter=",", comments="", fmt="%.10g",
               header="idx,x,y,curvature,kappa_signed,speed,xprime,yprime,x2,y2")
    return out_csv


def plot_outputs(prefix: str, P, kappa):
    hist_png, overlay_png = f"{prefix}_curvature_hist.png", f"{prefix}_curvature_overlay.png"
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.figure(figsize=(6, 4))
        plt.hist(kappa, bins=64)
        plt.xlabel(r"Curvature $\kappa$"); plt.ylabel("Count"); plt.title("Local-Polynomial Curvature Histogram")
        plt.tight_layout(); plt.savefig(hist_png, dpi=200); plt.close()
        plt.figure(figsize=(5, 5))
        sc = plt.scatter(P[:, 0], P[:, 1], c=kappa, s=8)
        plt.axis("equal"); plt.axis("off")
        plt.colorbar(sc, fraction=0.046, pad=0.04).set_label(r"$\kappa$")
        plt.title("Curvature Overlay (Local-Polynomial)")
        plt.tight_layout(); plt.savefig(overlay_png, dpi=220); plt.close()
    except ImportError:
        from .png import colored_scatter_png, histogram_png
        histogram_png(hist_png, kappa, bins=64)
        colored_scatter_png(overlay_png, P[:, 0], P[:, 1], kappa)
    return hist_png, overlay_png


def write_summary(prefix: str, kappa) -> str:
    out_txt = f"{prefix}_summary.txt"
    stats = (("n", len(kappa)), ("mean", float(np.mean(kappa))), ("median", float(np.median(kappa))), ("std", float(np.std(kappa))),
             ("q05", float(np.quantile(kappa, 0.05))), ("q95", float(np.quantile(kappa, 0.95))), ("max", float(np.max(kappa))))
    with open(out_txt, "w") as f:
        f.write("Local-Polynomial Curvature Summary\n")
        f.write("\n".join(f"{k}: {v:.10g}" for k, v in stats) + "\n")
    return out_txt


def _truthy(s: str) -> bool:
    return s.lower() in ("true", "1", "yes", "y")


def main(argv=None):
    ap = argparse.ArgumentParser(description="Local-polynomial curvature on 2D boundary points.")
    ap.add_argument("--input_csv", required=True, help="CSV with ordered boundary points (columns: x,y or header with x,y).")
    ap.add_argument("--output_prefix", required=True, help="Prefix for outputs (CSV/PNG/TXT).")
    ap.add_argument("--neighbors", type=int, default=7, help="Use +-neighbors points for quadratic fits (window size=2*neighbors+1).")
    ap.add_argument("--closed", type=_truthy, default=True, help="Treat boundary as closed (wrap indices).")
    ap.add_argument("--stride", type=int, default=1, help="Evaluate every 'stride' points (others interpolated).")
    args = ap.parse_args(argv)

    P = load_points(args.input_csv)
    if P.shape[0] < 2 * args.neighbors + 1:
        print(f"ERROR: Need at least {2 * args.neighbors + 1} points; got {P.shape[0]}.", file=sys.stderr)
        sys.exit(2)
    kappa, kappa_s, speed, aux = compute_curvature_localpoly(P, neighbors=args.neighbors, closed=args.closed, stride=args.stride)
    out_csv = save_csv(args.output_prefix, P, kappa, kappa_s, speed, aux)
    hist_png, overlay_png = plot_outputs(args.output_prefix, P, kappa)
    out_txt = write_summary(args.output_prefix, kappa)
    print("Wrote:")
    for p in (out_csv, hist_png, overlay_png, out_txt):
        print("  ", p)


if __name__ == "__main__":
    main()

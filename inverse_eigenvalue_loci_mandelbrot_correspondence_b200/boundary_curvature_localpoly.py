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


CSV_COLUMNS = ("idx", "x", "y", "curvature", "kappa_signed", "speed", "xprime", "yprime", "x2", "y2")


def write_curvature_csv(prefix: str, P, kappa, kappa_s, speed, aux) -> str:
    """<prefix>_curvature.csv: one row per boundary point, 10 significant digits (:186-193)."""
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    path = f"{prefix}_curvature.csv"
    table = np.column_stack([np.arange(P.shape[0]), P[:, 0], P[:, 1], kappa, kappa_s, speed] + [aux[k] for k in CSV_COLUMNS[6:]])
    np.savetxt(path, table, delimiter=",", comments="", fmt="%.10g", header=",".join(CSV_COLUMNS))
    return path


def write_pictures(prefix: str, P, kappa):
    """<prefix>_curvature_hist.png (64-bin histogram) and <prefix>_curvature_overlay.png (curve coloured by kappa), :195-220."""
    hist_png, overlay_png = f"{prefix}_curvature_hist.png", f"{prefix}_curvature_overlay.png"
    try:
        import matplotlib
        matplotlib.use("Agg")
        from matplotlib import pyplot
    except ImportError:
        from .png import colored_scatter_png, histogram_png
        histogram_png(hist_png, kappa, bins=64)
        colored_scatter_png(overlay_png, P[:, 0], P[:, 1], kappa)
        return hist_png, overlay_png
    fig, ax = pyplot.subplots(figsize=(6, 4))
    ax.hist(kappa, bins=64)
    ax.set(xlabel=r"Curvature $\kappa$", ylabel="Count", title="Local-Polynomial Curvature Histogram")
    fig.tight_layout(); fig.savefig(hist_png, dpi=200); pyplot.close(fig)
    fig, ax = pyplot.subplots(figsize=(5, 5))
    dots = ax.scatter(P[:, 0], P[:, 1], c=kappa, s=8)
    ax.axis("equal"); ax.axis("off"); ax.set_title("Curvature Overlay (Local-Polynomial)")
    fig.colorbar(dots, fraction=0.046, pad=0.04).set_label(r"$\kappa$")
    fig.tight_layout(); fig.savefig(overlay_png, dpi=220); pyplot.close(fig)
    return hist_png, overlay_png


def write_summary(prefix: str, kappa) -> str:
    """<prefix>_summary.txt: n, mean, median, std, 5 % / 95 % quantiles, max (:222-240)."""
    path = f"{prefix}_summary.txt"
    k = np.asarray(kappa, dtype=float)
    q05, q95 = (float(np.quantile(k, q)) for q in (0.05, 0.95))
    entries = {"n": len(k), "mean": float(k.mean()), "median": float(np.median(k)), "std": float(k.std()), "q05": q05, "q95": q95,
               "max": float(k.max())}
    with open(path, "w") as f:
        f.write("Local-Polynomial Curvature Summary\n" + "".join(f"{name}: {value:.10g}\n" for name, value in entries.items()))
    return path


def _truthy(s: str) -> bool:
    return s.lower() in ("true", "1", "yes", "y")


def main(argv=None):
    ap = argparse.ArgumentParser(description="Local-polynomial curvature on 2D boundary points.")
    ap.add_argument("--input_csv", required=True, help="ordered boundary points, two columns with or without an x,y header")
    ap.add_argument("--output_prefix", required=True, help="where <prefix>_curvature.csv, the two PNGs and <prefix>_summary.txt go")
    ap.add_argument("--neighbors", type=int, default=7, help="half width m of the fitting window (2m+1 points)")
    ap.add_argument("--closed", type=_truthy, default=True, help="True / 1 / yes: the curve is closed and the window wraps around")
    ap.add_argument("--stride", type=int, default=1, help="fit every stride-th point, interpolate the rest")
    args = ap.parse_args(argv)

    P = load_points(args.input_csv)
    if P.shape[0] < 2 * args.neighbors + 1:
        print(f"ERROR: Need at least {2 * args.neighbors + 1} points; got {P.shape[0]}.", file=sys.stderr)
        sys.exit(2)
    kappa, kappa_s, speed, aux = compute_curvature_localpoly(P, neighbors=args.neighbors, closed=args.closed, stride=args.stride)
    written = [write_curvature_csv(args.output_prefix, P, kappa, kappa_s, speed, aux), *write_pictures(args.output_prefix, P, kappa),
               write_summary(args.output_prefix, kappa)]
    print("Wrote:")
    for path in written:
        print("  ", path)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""GPU drop-in for the reference script construct_boundary_alpha.py (README Step 2; same CLI, same outputs).

    python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.construct_boundary_alpha \\
        --input_csv 0_data/construct_points.csv --alpha 30.0 --output_prefix outputs/construct

Outputs (construct_boundary_alpha.py:127-160):
  <prefix>_boundary.csv   ordered boundary points, header "x,y", "%.18e"
  <prefix>_edges.csv      boundary edges (i, j) as the alpha filter reports them, header "i,j", "%d"
  <prefix>_boundary.png   cloud + traced boundary
  <prefix>_meta.txt       alpha / N / ordered_points

`--main_loop` switches to what construct_boundary_alpha_spyder_v2.py does (README step 2): every boundary component is traced,
the longest closed loop (else the longest open chain) is kept and resampled to `--target_n` points by arclength;
<prefix>_edges.csv then lists consecutive vertex pairs and <prefix>_meta.txt gains a `closed=` line (:180-199 there).

The triangulation is scipy's Delaunay as in the reference; the radius test and edge-multiplicity count run on the device
(alpha_shape.alpha_shape_edges -> lm_alpha_shape_edges); the walk over the boundary edges is the reference's.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from .alpha_shape import (alpha_shape_edges, circumradius, connected_components, densify, main_boundary,  # noqa: F401
                          order_boundary, trace_loop_or_chain)                                          # reference-compatible names
from .curvature import load_points  # noqa: F401

NO_EDGES_MSG = "Alpha-shape produced no boundary edges. Try smaller alpha (tighter) or larger (looser)."


def save_outputs(P: np.ndarray, edges, ordered_idx, alpha: float, output_prefix: str, resample_to: int = 0, closed=None):
    """resample_to > 0: the v2 script's outputs (densified boundary, consecutive edge list, closed= in the meta file)."""
    B = P[ordered_idx, :]
    if resample_to:
        B = densify(B, resample_to)
        edges = np.c_[np.arange(len(ordered_idx) - 1), np.arange(1, len(ordered_idx))]
    os.makedirs(os.path.dirname(output_prefix), exist_ok=True)
    b_csv = f"{output_prefix}_boundary.csv"
    np.savetxt(b_csv, B, delimiter=",", header="x,y", comments="")
    e_csv = f"{output_prefix}_edges.csv"
    np.savetxt(e_csv, np.asarray(edges, dtype=int), fmt="%d", delimiter=",", header="i,j", comments="")
    out_png = f"{output_prefix}_boundary.png"
    try:
        import matplotlib
        matplotlib.use("Agg")
        from matplotlib import pyplot
    except ImportError:
        from .png import cloud_with_polyline_png
        cloud_with_polyline_png(out_png, P, B)
    else:                                     # the reference's picture (:145-151): faint cloud, boundary line on top
        fig, ax = pyplot.subplots(figsize=(6, 6))
        ax.scatter(P[:, 0], P[:, 1], s=2, alpha=0.25)
        ax.plot(B[:, 0], B[:, 1], lw=1.0)
        ax.axis("equal"); ax.axis("off")
        fig.tight_layout(); fig.savefig(out_png, dpi=220); pyplot.close(fig)
    meta = f"{output_prefix}_meta.txt"
    with open(meta, "w") as f:
        f.write(f"alpha={alpha}\nN={len(P)}\nordered_points={len(B)}\n" + (f"closed={closed}\n" if closed is not None else ""))
    return b_csv, e_csv, out_png, meta


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input_csv", required=True)
    ap.add_argument("--alpha", type=float, default=25.0, help="triangles with circumradius >= 1/alpha are dropped (larger = tighter)")
    ap.add_argument("--output_prefix", required=True)
    ap.add_argument("--main_loop", action="store_true", help="keep the longest closed boundary component and resample it (the v2 script)")
    ap.add_argument("--target_n", type=int, default=1500, help="points of the resampled boundary with --main_loop")
    args = ap.parse_args(argv)

    P = load_points(args.input_csv)
    edges = alpha_shape_edges(P, alpha=args.alpha)
    if len(edges) == 0:
        raise SystemExit(NO_EDGES_MSG)
    if args.main_loop:
        try:
            ordered_idx, was_closed = main_boundary(edges)
        except ValueError as e:
            raise SystemExit(str(e))
        paths = save_outputs(P, edges, ordered_idx, args.alpha, args.output_prefix, resample_to=args.target_n, closed=was_closed)
    else:
        ordered_idx = order_boundary(P, edges)
        paths = save_outputs(P, edges, ordered_idx, args.alpha, args.output_prefix)
    print("Wrote:")
    for p in paths:
        print(" ", p)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""GPU drop-in for the reference script mandelbrot_boundary_sample.py (same CLI, same outputs).

    python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.mandelbrot_boundary_sample \\
        --xlim -2.1 0.9 --ylim -1.5 1.5 --res 2000 --max_iter 500 --level 0.96 --output_prefix outputs/mandel

Outputs (mandelbrot_boundary_sample.py:71-90):
  <prefix>_boundary.csv   header "x,y", "%.18e" (np.savetxt defaults)
  <prefix>_boundary.png   scatter of the boundary sample
  <prefix>_meta.txt       xlim / ylim / res / max_iter / level

compute_grid and extract_contour keep the reference signatures; the dwell grid stays on the
GPU between the two stages when main() drives them (only the boundary polyline comes back).

Several GPUs: launch the same module with torchrun, one process per GPU --
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.mandelbrot_boundary_sample --res 32768 ...
the rows are sharded over the ranks (sharding.ShardedBoundary), rank 0 writes the three files.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import contour as _contour
from .contour import extract_contour  # noqa: F401  (reference-compatible name)
from .escape import compute_grid, mandelbrot_dwell  # noqa: F401

FAIL_MSG = "Failed to extract a usable contour. Try different --level or higher --res."


def boundary_from_window(xlim, ylim, res: int, max_iter: int, level_frac: float):
    """compute_grid + extract_contour with the dwell grid kept resident on the device.  Under torchrun
    (WORLD_SIZE > 1) the rows are sharded over the ranks; the contour comes back on rank 0 (None elsewhere)."""
    xs = np.linspace(xlim[0], xlim[1], res)
    ys = np.linspace(ylim[0], ylim[1], res)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        from . import _shim, sharding
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local); _shim.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        job = sharding.ShardedBoundary(xs, ys, max_iter, level_frac * max_iter)
        lines = job.run()
        dist.barrier()
        dist.destroy_process_group()
        return xs, ys, (_contour.longest(lines) if lines is not None else None)
    lines, _ = _contour.boundary_sample(xs, ys, max_iter, level_frac * max_iter)
    return xs, ys, _contour.longest(lines)


def save_outputs(contour: np.ndarray, output_prefix: str, xlim, ylim, res: int, max_iter: int, level: float):
    out_csv = f"{output_prefix}_boundary.csv"
    os.makedirs(os.path.dirname(output_prefix), exist_ok=True)
    np.savetxt(out_csv, contour, delimiter=",", header="x,y", comments="")
    out_png = f"{output_prefix}_boundary.png"
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.figure(figsize=(6, 6))
        plt.scatter(contour[:, 0], contour[:, 1], s=1)
        plt.axis("equal"); plt.axis("off")
        plt.tight_layout()
        plt.savefig(out_png, dpi=220)
        plt.close()
    except ImportError:
        from .png import scatter_png
        scatter_png(out_png, contour[:, 0], contour[:, 1])
    out_meta = f"{output_prefix}_meta.txt"
    with open(out_meta, "w") as f:
        f.write(f"xlim={xlim}\nylim={ylim}\nres={res}\nmax_iter={max_iter}\nlevel={level}\n")
    return out_csv, out_png, out_meta


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--xlim", nargs=2, type=float, default=[-2.1, 0.9])
    ap.add_argument("--ylim", nargs=2, type=float, default=[-1.5, 1.5])
    ap.add_argument("--res", type=int, default=1500)
    ap.add_argument("--max_iter", type=int, default=400)
    ap.add_argument("--level", type=float, default=0.96, help="Fraction of max_iter for isocontour")
    ap.add_argument("--output_prefix", required=True)
    args = ap.parse_args(argv)

    _, _, contour = boundary_from_window(args.xlim, args.ylim, args.res, args.max_iter, args.level)
    if int(os.environ.get("RANK", "0")) != 0:
        return                                   # sharded run: rank 0 holds the contour and writes the files
    if contour is None or contour.shape[0] < 50:
        raise SystemExit(FAIL_MSG)
    out_csv, out_png, out_meta = save_outputs(contour, args.output_prefix, args.xlim, args.ylim, args.res,
                                              args.max_iter, args.level)
    print("Wrote:")
    print(" ", out_csv)
    print(" ", out_png)
    print(" ", out_meta)


if __name__ == "__main__":
    main()

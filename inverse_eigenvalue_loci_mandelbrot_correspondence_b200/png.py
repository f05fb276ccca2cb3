"""Minimal PNG writers for the scripts' pictures when matplotlib is not installed.

The reference (mandelbrot_boundary_sample.py:76-82) saves plt.scatter(x, y, s=1) on a 6x6 in
figure at dpi 220 (1320x1320 px), equal axes, axes off.  With matplotlib present the host
script uses it unchanged; otherwise this module rasterises the same picture (one dot per
vertex in matplotlib's default blue on white) with numpy + zlib.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


def scatter_png(path: str, x, y, size_px: int = 1320, margin: float = 0.04, color=(31, 119, 180)) -> None:
    x = np.asarray(x, dtype=float); y = np.asarray(y, dtype=float)
    img = np.full((size_px, size_px, 3), 255, dtype=np.uint8)
    if x.size:
        xmin, xmax, ymin, ymax = x.min(), x.max(), y.min(), y.max()
        span = max(xmax - xmin, ymax - ymin, 1e-300)          # equal aspect
        cx, cy = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
        scale = (1.0 - 2.0 * margin) * (size_px - 1) / span
        px = np.clip(np.rint((x - cx) * scale + 0.5 * (size_px - 1)).astype(int), 0, size_px - 1)
        py = np.clip(np.rint((cy - y) * scale + 0.5 * (size_px - 1)).astype(int), 0, size_px - 1)
        for dx in (0, 1):                                       # s=1 pt^2 at 220 dpi is about 3 px across
            for dy in (0, 1):
                img[np.clip(py + dy, 0, size_px - 1), np.clip(px + dx, 0, size_px - 1)] = color
    write_png(path, img)


def write_png(path: str, img: np.ndarray) -> None:
    """uint8 [h, w, 3] -> 8-bit RGB PNG."""
    h, w = img.shape[:2]
    raw = b"".join(b"\x00" + np.ascontiguousarray(img[r]).tobytes() for r in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def _to_pixels(x, y, size_px: int, margin: float, box=None):
    x = np.asarray(x, dtype=float); y = np.asarray(y, dtype=float)
    xmin, xmax, ymin, ymax = box if box is not None else (x.min(), x.max(), y.min(), y.max())
    span = max(xmax - xmin, ymax - ymin, 1e-300)              # equal aspect
    cx, cy = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
    scale = (1.0 - 2.0 * margin) * (size_px - 1) / span
    px = np.clip(np.rint((x - cx) * scale + 0.5 * (size_px - 1)).astype(int), 0, size_px - 1)
    py = np.clip(np.rint((cy - y) * scale + 0.5 * (size_px - 1)).astype(int), 0, size_px - 1)
    return px, py


def cloud_with_polyline_png(path: str, P, B, size_px: int = 1320, margin: float = 0.04) -> None:
    """construct_boundary_alpha.py:145-151: the cloud as faint dots, the ordered boundary as a line on top."""
    P = np.asarray(P, dtype=float); B = np.asarray(B, dtype=float)
    img = np.full((size_px, size_px, 3), 255, dtype=np.uint8)
    box = (P[:, 0].min(), P[:, 0].max(), P[:, 1].min(), P[:, 1].max())
    px, py = _to_pixels(P[:, 0], P[:, 1], size_px, margin, box)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            img[np.clip(py + dy, 0, size_px - 1), np.clip(px + dx, 0, size_px - 1)] = (199, 221, 236)     # blue at alpha 0.25 on white
    if len(B) >= 2:
        seg = np.hypot(np.diff(B[:, 0]), np.diff(B[:, 1]))
        steps = np.maximum(2, np.ceil(seg / max(seg.sum(), 1e-300) * 8 * size_px).astype(int))
        xs = np.concatenate([np.linspace(B[k, 0], B[k + 1, 0], steps[k]) for k in range(len(B) - 1)])
        ys = np.concatenate([np.linspace(B[k, 1], B[k + 1, 1], steps[k]) for k in range(len(B) - 1)])
        lx, ly = _to_pixels(xs, ys, size_px, margin, box)
        for d in (0, 1):
            img[np.clip(ly + d, 0, size_px - 1), lx] = (31, 119, 180)
            img[ly, np.clip(lx + d, 0, size_px - 1)] = (31, 119, 180)
    write_png(path, img)


def histogram_png(path: str, values, bins: int = 64, size=(1200, 800)) -> None:
    """plt.hist(values, bins=64) as plain bars (boundary_curvature_localpoly.py:196-205)."""
    w, h = size
    img = np.full((h, w, 3), 255, dtype=np.uint8)
    counts, _ = np.histogram(np.asarray(values, dtype=float), bins=bins)
    top = max(int(counts.max()), 1)
    x0, x1, y0, y1 = int(0.08 * w), int(0.97 * w), int(0.06 * h), int(0.90 * h)
    img[y1, x0:x1] = 0; img[y0:y1, x0] = 0
    bw = (x1 - x0) / bins
    for k, c in enumerate(counts):
        a, b = x0 + int(round(k * bw)) + 1, x0 + int(round((k + 1) * bw))
        t = y1 - int(round((y1 - y0) * c / top))
        img[t:y1, a:max(b, a + 1)] = (31, 119, 180)
    write_png(path, img)


def colored_scatter_png(path: str, x, y, c, size_px: int = 1100, margin: float = 0.06) -> None:
    """plt.scatter(x, y, c=kappa, s=8) with a viridis-like ramp (boundary_curvature_localpoly.py:207-218)."""
    c = np.asarray(c, dtype=float)
    img = np.full((size_px, size_px, 3), 255, dtype=np.uint8)
    lo, hi = (np.nanmin(c), np.nanmax(c)) if c.size else (0.0, 1.0)
    t = np.clip((c - lo) / max(hi - lo, 1e-300), 0.0, 1.0)
    stops = np.array([[68, 1, 84], [59, 82, 139], [33, 145, 140], [94, 201, 98], [253, 231, 37]], dtype=float)
    pos = t * (len(stops) - 1)
    i0 = np.clip(pos.astype(int), 0, len(stops) - 2)
    rgb = (stops[i0] + (stops[i0 + 1] - stops[i0]) * (pos - i0)[:, None]).astype(np.uint8)
    px, py = _to_pixels(x, y, size_px, margin)
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            img[np.clip(py + dy, 0, size_px - 1), np.clip(px + dx, 0, size_px - 1)] = rgb
    write_png(path, img)

"""Minimal PNG writer for <prefix>_boundary.png when matplotlib is not installed.

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
    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(size_px))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", size_px, size_px, 8, 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))

"""Shared test helpers (CPU side)."""
from __future__ import annotations

import numpy as np

E, N, W, S = 0, 1, 2, 3


def _exit_edge(edge: int, d: int) -> int:
    if d == 0:
        return (edge + 2) & 3
    return (edge + 3) & 3 if d > 0 else (edge + 1) & 3


def records_from_dwell(dwell: np.ndarray, xs: np.ndarray, ys: np.ndarray, level: float, row_offset: int = 0,
                       nx_global: int | None = None) -> np.ndarray:
    """numpy restatement of the record format lm_contour_classify_dev emits (csrc/lm_contour.cu),
    used to exercise the host-only linker (lm_contour_link) without a GPU."""
    ny, nx = dwell.shape
    recs = []
    for j in range(ny - 1):
        for i in range(nx - 1):
            z = [[int(dwell[j, i]), int(dwell[j, i + 1])], [int(dwell[j + 1, i]), int(dwell[j + 1, i + 1])]]
            sw, se, nw, ne = (float(z[0][0]) > level, float(z[0][1]) > level, float(z[1][0]) > level, float(z[1][1]) > level)
            s = sw + se + nw + ne
            if s == 0 or s == 4:
                continue
            zmid = 0.25 * (((float(z[0][0]) + float(z[0][1])) + float(z[1][0])) + float(z[1][1]))
            right = zmid > level
            ent = {E: se and not ne, N: ne and not nw, W: nw and not sw, S: sw and not se}
            segs = []
            for edge in (S, W, N, E):
                if not ent[edge]:
                    continue
                pl, pr = {E: (sw, nw), N: (se, sw), W: (ne, se), S: (nw, ne)}[edge]
                if (not pl) and pr:
                    d = -1 if right else 1
                elif (not pl) and (not pr):
                    d = 1
                elif pl and pr:
                    d = -1
                else:
                    d = 0
                segs.append((edge, _exit_edge(edge, d)))
            v = np.zeros(4)
            for k, (_, ex) in enumerate(segs):
                (dj1, di1, dj2, di2) = {E: (0, 1, 1, 1), N: (1, 1, 1, 0), W: (1, 0, 0, 0), S: (0, 0, 0, 1)}[ex]
                z1, z2 = np.float64(z[dj1][di1]), np.float64(z[dj2][di2])
                f = (z2 - np.float64(level)) / (z2 - z1)
                g = np.float64(1.0) - f
                v[2 * k] = xs[i + di1] * f + xs[i + di2] * g
                v[2 * k + 1] = ys[j + dj1] * f + ys[j + dj2] * g
            config = (8 if nw else 0) | (4 if ne else 0) | (2 if sw else 0) | (1 if se else 0)
            meta = config | (16 if right else 0) | (len(segs) << 16)
            for k, (en, ex) in enumerate(segs):
                meta |= (en << (8 + 4 * k)) | (ex << (10 + 4 * k))
            quad = (row_offset + j) * (nx_global or nx) + i
            w1 = (z[0][0] & 0xffffffff) | ((z[0][1] & 0xffffffff) << 32)
            w2 = (z[1][0] & 0xffffffff) | ((z[1][1] & 0xffffffff) << 32)
            rec = np.zeros(8, dtype=np.int64)
            rec[0] = quad
            rec[1] = np.array([w1], dtype=np.uint64).view(np.int64)[0]
            rec[2] = np.array([w2], dtype=np.uint64).view(np.int64)[0]
            rec[3] = meta
            rec[4:8] = v.view(np.int64)
            recs.append(rec)
    if not recs:
        return np.zeros((0, 8), dtype=np.int64)
    return np.stack(recs)


def lines_equal(a, b) -> bool:
    if len(a) != len(b):
        return False
    return all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b))


def assert_columns_close(got, want, rtol: float, afrac: float, cols=None) -> None:
    """Column-wise allclose with an absolute floor proportional to each column's largest magnitude."""
    got = np.asarray(got); want = np.asarray(want)
    for j in (range(want.shape[1]) if cols is None else cols):
        np.testing.assert_allclose(got[:, j], want[:, j], rtol=rtol, atol=afrac * float(np.abs(want[:, j]).max()),
                                   err_msg=f"column {j}")

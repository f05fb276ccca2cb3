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


def link_records_numpy(records: np.ndarray, xs: np.ndarray, ys: np.ndarray, level: float):
    """numpy restatement of the DEVICE linker (csrc/lm_contour_link.cu): successor table, pointer jumping along the
    predecessors (heads of the open chains / smallest node id of every cycle), list ranking from the leaders, scan
    over the leaders (open lines first, then loops), scatter of the vertices.  Used to exercise the sharding
    logic on the CPU (gloo tests) and, against the sequential mpl2014 restatement of the oracle, to pin the
    ordering rule the kernels implement.  -> list of (N, 2) arrays."""
    records = np.ascontiguousarray(records, dtype=np.int64).reshape(-1, 8)
    nx, ny = xs.size, ys.size
    n = records.shape[0]
    if n == 0:
        return []
    quad = records[:, 0]
    meta = records[:, 3].astype(np.uint64)
    nseg = ((meta >> np.uint64(16)) & np.uint64(3)).astype(np.int64)
    NS = 2 * n
    ids = np.arange(NS)
    k_of = ids >> 1; s_of = ids & 1
    valid = s_of < nseg[k_of]
    entry = ((meta[k_of] >> (8 + 4 * s_of).astype(np.uint64)) & np.uint64(3)).astype(np.int64)
    exit_ = ((meta[k_of] >> (10 + 4 * s_of).astype(np.uint64)) & np.uint64(3)).astype(np.int64)
    qi = (quad % nx)[k_of]; qj = (quad // nx)[k_of]

    def on_border(e):
        return ((e == E) & (qi == nx - 2)) | ((e == N) & (qj == ny - 2)) | ((e == W) & (qi == 0)) | ((e == S) & (qj == 0))

    head = valid & on_border(entry)
    leaves = on_border(exit_)
    step = np.select([exit_ == E, exit_ == W, exit_ == N], [1, -1, nx], default=-nx)
    q2 = quad[k_of] + step
    k2 = np.searchsorted(quad, q2)
    k2c = np.minimum(k2, n - 1)
    found = valid & ~leaves & (quad[k2c] == q2)
    want = (exit_ + 2) & 3
    succ = np.full(NS, -1)
    for s2 in range(2):
        t = 2 * k2c + s2
        ok = found & valid[t] & (entry[t] == want)
        succ[ok] = t[ok]
    if (valid & ~leaves & (succ < 0)).any():
        raise ValueError("inconsistent crossing records")
    pred = np.full(NS, -1)
    src = ids[succ >= 0]
    pred[succ[src]] = src
    if (valid & ~head & (pred < 0)).any():
        raise ValueError("inconsistent crossing records")
    live = ids[valid]
    rounds = max(int(np.ceil(np.log2(max(live.size, 2)))), 1)
    p = np.where(head | ~valid, ids, pred); m = ids.copy()
    for _ in range(rounds):
        m = np.minimum(m, m[p]); p = p[p]
    leader = np.where(head[p], p, m)
    isl = valid & (leader == ids)
    p = np.where(isl | ~valid, ids, pred); d = np.where(isl | ~valid, 0, 1)
    for _ in range(rounds):
        d = d + d[p]; p = p[p]
    assert (p[live] == leader[live]).all()
    nvl = np.zeros(NS, dtype=np.int64)
    last = valid & ((succ == -1) | (succ == leader))
    nvl[leader[last]] = d[last] + 2
    order = np.concatenate([ids[isl & head], ids[isl & ~head]])
    offs = np.concatenate([[0], np.cumsum(nvl[order])]).astype(np.int64)
    base = np.zeros(NS, dtype=np.int64); base[order] = offs[:-1]
    verts = np.full((offs[-1], 2), np.nan)
    xy = records[:, 4:8].copy().view(np.float64).reshape(n, 2, 2)
    nstart = np.zeros(NS, dtype=bool); nstart[order] = (~head[order]) & (entry[order] == N)
    L = leader[live]
    verts[base[L] + d[live] + np.where(nstart[L], 0, 1)] = xy[live >> 1, live & 1]
    corners = records[:, 1:3].copy().view(np.uint32).view(np.int32).reshape(n, 2, 2)     # [k][dj][di]
    for v in order:
        if nstart[v]:
            verts[base[v] + nvl[v] - 1] = xy[v >> 1, v & 1]
        else:
            k = v >> 1
            dj1, di1, dj2, di2 = {E: (0, 1, 1, 1), N: (1, 1, 1, 0), W: (1, 0, 0, 0), S: (0, 0, 0, 1)}[int(entry[v])]
            z1, z2 = np.float64(corners[k, dj1, di1]), np.float64(corners[k, dj2, di2])
            f = (z2 - np.float64(level)) / (z2 - z1); g = np.float64(1.0) - f
            verts[base[v]] = (xs[qi[v] + di1] * f + xs[qi[v] + di2] * g, ys[qj[v] + dj1] * f + ys[qj[v] + dj2] * g)
    return [verts[offs[t]:offs[t + 1]] for t in range(offs.size - 1)]


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


def true_roots_mp(top_row, digits: int = 60) -> np.ndarray:
    """Roots of x^d - a_1 x^(d-1) - ... - a_d to ~`digits` digits (mpmath Durand-Kerner in extended precision), rounded to
    complex128: the arbiter between two double-precision solvers on ill-conditioned (clustered) roots."""
    import mpmath as mp
    top_row = [float(v) for v in top_row]
    d = len(top_row)
    while d > 0 and top_row[d - 1] == 0.0:           # trailing zeros are roots at 0
        d -= 1
    nzero = len(top_row) - d
    with mp.workdps(digits):
        coeffs = [mp.mpf(1)] + [-mp.mpf(v) for v in top_row[:d]]
        roots = mp.polyroots(coeffs, maxsteps=500, extraprec=4 * digits) if d > 0 else []
        out = [complex(r) for r in roots]
    return np.array(out + [0j] * nzero, dtype=np.complex128)


def multiset_distance(a, b) -> float:
    """Max relative distance after greedy nearest matching of two complex multisets (relative to the b values)."""
    a = np.asarray(a, dtype=np.complex128).ravel(); b = np.asarray(b, dtype=np.complex128).ravel()
    assert a.shape == b.shape
    used = np.zeros(a.size, dtype=bool)
    worst = 0.0
    for v in b:
        dist = np.abs(a - v); dist[used] = np.inf
        k = int(np.argmin(dist)); used[k] = True
        worst = max(worst, float(dist[k] / max(abs(v), 1e-300)))
    return worst

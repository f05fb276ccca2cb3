"""numpy prototype of the device contour linker (csrc/lm_contour_link.cu), checked against the mpl2014
restatement (oracle.contour_lines) on random integer fields full of saddles and on Mandelbrot windows.

Development tool: it pins the *ordering rule* the kernels implement --
  node            = (record k, segment s), id 2k+s, monotone in mpl2014's scan order (raster quad, edges S,W,N,E)
  open lines      = chains whose first node enters through a grid-border edge, in order of that node's id
  closed lines    = cycles, in order of their smallest node id, started at that node
  vertices        = [entry vertex of the first node] + exit vertex of every node; a closed line whose first node
                    is entered through its N edge drops the entry vertex and repeats its first exit vertex
-- before they were written in CUDA.  CPU only (uses the oracle, like tests/)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import records_from_dwell  # noqa: E402
from oracle import oracle  # noqa: E402

E, N, W, S = 0, 1, 2, 3


def link(records, xs, ys, level):
    nx, ny = xs.size, ys.size
    n = records.shape[0]
    quad = records[:, 0]
    meta = records[:, 3].astype(np.uint64)
    nseg = ((meta >> 16) & 3).astype(int)
    NS = 2 * n
    valid = np.zeros(NS, bool); entry = np.zeros(NS, int); exit_ = np.zeros(NS, int)
    for s in range(2):
        valid[s::2] = nseg > s
        entry[s::2] = (meta >> (8 + 4 * s)) & 3
        exit_[s::2] = (meta >> (10 + 4 * s)) & 3
    qi = quad % nx; qj = quad // nx
    index = {int(q): k for k, q in enumerate(quad)}
    succ = np.full(NS, -2); pred = np.full(NS, -1); head = np.zeros(NS, bool)

    def on_border(k, e):
        return (e == E and qi[k] == nx - 2) or (e == N and qj[k] == ny - 2) or (e == W and qi[k] == 0) or (e == S and qj[k] == 0)

    for v in range(NS):
        if not valid[v]:
            continue
        k = v >> 1
        head[v] = on_border(k, entry[v])
        if on_border(k, exit_[v]):
            succ[v] = -1
            continue
        q2 = int(quad[k]) + {E: 1, W: -1, N: nx, S: -nx}[exit_[v]]
        k2 = index[q2]
        want = (exit_[v] + 2) & 3
        t = [2 * k2 + s for s in range(nseg[k2]) if entry[2 * k2 + s] == want]
        assert len(t) == 1
        succ[v] = t[0]; pred[t[0]] = v
    ids = np.nonzero(valid)[0]
    rounds = max(int(np.ceil(np.log2(max(ids.size, 2)))), 1)
    # phase A: (p, m) jumping along pred; heads are terminals
    p = np.where(head, np.arange(NS), pred); m = np.arange(NS)
    for _ in range(rounds):
        p2 = p.copy(); m2 = m.copy()
        for v in ids:
            if head[v]:
                continue
            m2[v] = min(m[v], m[p[v]]); p2[v] = p[p[v]]
        p, m = p2, m2
    leader = np.where(head[np.clip(p, 0, NS - 1)], p, m)
    # phase B: rank from the leader
    isl = leader == np.arange(NS)
    p = np.where(isl, np.arange(NS), pred); d = np.where(isl, 0, 1)
    for _ in range(rounds):
        p2 = p.copy(); d2 = d.copy()
        for v in ids:
            d2[v] = d[v] + d[p[v]]; p2[v] = p[p[v]]
        p, d = p2, d2
    assert all(p[v] == leader[v] for v in ids)
    # line sizes: written by the last node of each line
    nvl = np.zeros(NS, int)
    for v in ids:
        if succ[v] == -1 or succ[v] == leader[v]:
            nvl[leader[v]] = d[v] + 2
    openl = [v for v in ids if isl[v] and head[v]]
    cycl = [v for v in ids if isl[v] and not head[v]]
    order = openl + cycl
    offs = np.concatenate([[0], np.cumsum([nvl[v] for v in order])]).astype(np.int64)
    verts = np.full((offs[-1], 2), np.nan)
    base = {v: offs[t] for t, v in enumerate(order)}
    xy = records[:, 4:8].copy().view(np.float64).reshape(n, 2, 2)

    def entry_vertex(v):
        k = v >> 1; e = entry[v]
        dj1, di1, dj2, di2 = {E: (0, 1, 1, 1), N: (1, 1, 1, 0), W: (1, 0, 0, 0), S: (0, 0, 0, 1)}[e]
        zc = np.array([records[k, 1], records[k, 2]]).view(np.uint32).view(np.int32).reshape(2, 2)
        z1, z2 = np.float64(zc[dj1][di1]), np.float64(zc[dj2][di2])
        f = (z2 - np.float64(level)) / (z2 - z1); g = np.float64(1.0) - f
        return xs[qi[k] + di1] * f + xs[qi[k] + di2] * g, ys[qj[k] + dj1] * f + ys[qj[k] + dj2] * g

    for v in ids:
        L = leader[v]; b = base[L]
        nstart = (not head[L]) and entry[L] == N
        pos = b + d[v] + (0 if nstart else 1)
        verts[pos] = xy[v >> 1, v & 1]
        if v == L:
            if nstart:
                verts[b + nvl[L] - 1] = xy[v >> 1, v & 1]
            else:
                verts[b] = entry_vertex(v)
    return verts, offs


def check(dwell, xs, ys, level):
    rec = records_from_dwell(dwell, xs, ys, level)
    want = oracle.contour_lines(xs, ys, dwell.astype(np.float64), level)
    verts, offs = link(rec, xs, ys, level)
    got = [verts[offs[k]:offs[k + 1]] for k in range(offs.size - 1)]
    assert len(got) == len(want), (len(got), len(want))
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.array_equal(a, b)
    return len(got)


if __name__ == "__main__":
    oracle.build()
    rng = np.random.default_rng(5)
    tot = 0
    for trial in range(60):
        ny, nx = rng.integers(2, 40, size=2)
        dwell = rng.integers(0, 4, size=(ny, nx)).astype(np.int32)
        xs = np.linspace(-1.0, 2.0, nx); ys = np.linspace(0.5, 1.7, ny)
        for level in (0.5, 1.0, 1.5, 2.5):
            tot += check(dwell, xs, ys, level)
    xs = np.linspace(-2.1, 0.9, 150); ys = np.linspace(-1.5, 1.5, 150)
    d, _ = oracle.dwell_grid(xs, ys, 60)
    for level in (0.96 * 60, 3.0, 10.5):
        tot += check(d.astype(np.int32), xs, ys, level)
    print("ok", tot, "lines")

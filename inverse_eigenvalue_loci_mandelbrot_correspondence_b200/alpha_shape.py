"""Host side of the alpha-shape boundary extraction: same names and return values as

  circumradius(p, q, r), alpha_shape_edges(P, alpha), order_boundary(P, edges)     construct_boundary_alpha.py:45-125

The per-triangle radius test and the edge-multiplicity count (a Python loop over tri.simplices and a dict in the
reference) run in liblm_b200.so:lm_alpha_shape_edges; the boundary edges come back in the reference's own order.  The
triangulation is the reference's third-party call (scipy.spatial.Delaunay, Qhull) and stays what it is: pass
`simplices=` to supply your own.  order_boundary is the reference's sequential walk over the (few) boundary edges.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from ._shim import Stats

last_stats: dict = {}


def _filter(P, simplices, alpha, want_radius=False):
    P = np.asarray(P, dtype=np.float64)
    if P.ndim != 2 or P.shape[1] != 2:
        raise ValueError("P must have shape (N, 2)")
    x = np.ascontiguousarray(P[:, 0]); y = np.ascontiguousarray(P[:, 1])
    tri = np.ascontiguousarray(simplices, dtype=np.int32).reshape(-1, 3)
    ntri = tri.shape[0]
    keep = np.zeros(ntri, dtype=np.uint8)
    radius = np.empty(ntri, dtype=np.float64) if want_radius else None
    cap = 3 * ntri
    edges = np.empty((max(cap, 1), 2), dtype=np.int32)
    n = C.c_int64(0)
    st = Stats()
    _shim.call("lm_alpha_shape_edges", _shim.ptr(x), _shim.ptr(y), x.size, _shim.ptr(tri), ntri, float(alpha), _shim.ptr(keep),
               _shim.ptr(radius), _shim.ptr(edges), cap, C.byref(n), C.byref(st))
    global last_stats
    last_stats = st.as_dict()
    return keep.astype(bool), radius, edges[:int(n.value)]


def circumradii(P, simplices) -> np.ndarray:
    """circumradius of every triangle (inf for degenerate ones)."""
    return _filter(P, simplices, 1.0, want_radius=True)[1]


def circumradius(p, q, r) -> float:
    return float(circumradii(np.array([p, q, r], dtype=np.float64), [[0, 1, 2]])[0])


def delaunay_simplices(P) -> np.ndarray:
    from scipy.spatial import Delaunay          # the reference's own call (construct_boundary_alpha.py:59)
    return Delaunay(np.asarray(P, dtype=np.float64)).simplices


def alpha_shape_edges(P, alpha, simplices=None):
    """Boundary edges [(i, j), ...] (i < j) of the alpha shape, in the reference's order; [] when nothing is kept."""
    if simplices is None:
        simplices = delaunay_simplices(P)
    _, _, edges = _filter(P, simplices, alpha)
    return [(int(i), int(j)) for i, j in edges]


def order_boundary(P, edges):
    """Trace the boundary edges into an ordered vertex list (construct_boundary_alpha.py:84-125)."""
    from collections import defaultdict
    adj = defaultdict(list)
    for i, j in edges:
        adj[i].append(j); adj[j].append(i)
    start = None
    for k, v in adj.items():
        if len(v) == 1:
            start = k
            break
    if start is None:
        start = edges[0][0]
    ordered = []
    curr, prev = start, None
    while True:
        ordered.append(curr)
        nxt = None
        for n in adj[curr]:
            if n != prev:
                nxt = n
                break
        if nxt is None:
            break
        prev, curr = curr, nxt
        if curr == start:
            ordered.append(curr)
            break
        if len(ordered) > len(P) + 5:
            break
    return ordered

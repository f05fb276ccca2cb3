"""Host side of the alpha-shape boundary extraction: same names and return values as

  circumradius(p, q, r), alpha_shape_edges(P, alpha), order_boundary(P, edges)     construct_boundary_alpha.py:45-125

and, for the variant README step 2 runs (construct_boundary_alpha_spyder_v2.py:63-176: every boundary component traced, the
longest closed loop kept, the curve resampled to a fixed number of points by arclength),

  connected_components(edges), trace_loop_or_chain(adj, comp_nodes), main_boundary(edges), densify(B, target_n)

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


# ---- construct_boundary_alpha_spyder_v2.py: all components, longest closed loop, arclength resampling -------------------
# The traversal orders below are those of Python sets of ints / int tuples, as in the reference: they decide which vertex a
# loop starts at, so the same containers are used on purpose.

def _edge_key(u, w):
    return (u, w) if u < w else (w, u)


def connected_components(edges):
    """[(node set, edge list)] per connected component of the boundary-edge graph, and the adjacency lists (:63-85)."""
    from collections import defaultdict, deque
    adj = defaultdict(list)
    nodes = set()
    for i, j in edges:
        adj[i].append(j); adj[j].append(i)
        nodes.update((i, j))
    seen, comps = set(), []
    for root in nodes:
        if root in seen:
            continue
        seen.add(root)
        members, found, queue = {root}, [], deque([root])
        while queue:
            u = queue.popleft()
            for w in adj[u]:
                found.append(_edge_key(u, w))
                if w not in seen:
                    seen.add(w); members.add(w); queue.append(w)
        comps.append((members, list(set(found))))
    return comps, adj


def trace_loop_or_chain(adj, comp_nodes):
    """(ordered vertex list, is_closed) of one component (:87-118): a cycle when every vertex has degree 2, otherwise the
    longest simple chain found from the degree-1 vertices (or, failing those, from the irregular ones)."""
    irregular = [v for v in comp_nodes if len(adj[v]) != 2]
    budget = len(comp_nodes) + 5
    if not irregular and len(comp_nodes) > 2:
        first = next(iter(comp_nodes))
        path, prev, curr = [first], None, first
        for _ in range(budget):
            around = adj[curr]
            step = around[0] if around[0] != prev else (around[1] if len(around) > 1 else None)
            if step is None:
                break
            path.append(step)
            prev, curr = curr, step
            if curr == first:
                break
        return path, True
    starts = [v for v in irregular if len(adj[v]) == 1] or irregular or list(comp_nodes)
    longest = []
    for s0 in starts:
        path, visited, prev, curr = [s0], {s0}, None, s0
        for _ in range(budget):
            onward = [x for x in adj[curr] if x != prev]
            if not onward or onward[0] in visited:
                break
            prev, curr = curr, onward[0]
            path.append(curr); visited.add(curr)
        if len(path) > len(longest):
            longest = path
    return longest, False


def main_boundary(edges, min_vertices: int = 5):
    """(ordered vertex indices, was_closed) of the component the v2 script keeps (:127-152): the longest closed loop, or the
    longest open chain when no component closes; components traced to fewer than min_vertices entries are ignored.
    Raises ValueError when nothing usable is left."""
    from collections import defaultdict
    comps, _ = connected_components(edges)
    closed, opened = [], []
    for members, comp_edges in comps:
        local = defaultdict(list)
        for i, j in comp_edges:
            local[i].append(j); local[j].append(i)
        path, is_closed = trace_loop_or_chain(local, members)
        if len(path) >= min_vertices:
            (closed if is_closed else opened).append(path)
    if closed:
        return max(closed, key=len), True
    if opened:
        return max(opened, key=len), False
    raise ValueError("No usable boundary component found. Adjust alpha.")


def densify(B, target_n: int = 1500) -> np.ndarray:
    """Arclength resampling of the ordered boundary (:156-178): duplicates dropped (first occurrence kept, order preserved),
    the curve closed if it is not, then target_n points at uniform arclength by linear interpolation."""
    B = np.asarray(B, dtype=np.float64)
    _, first = np.unique(B, axis=0, return_index=True)
    B = B[np.sort(first)]
    if not np.allclose(B[0], B[-1]):
        B = np.vstack([B, B[0]])
    s = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(B, axis=0), axis=1))])
    if s[-1] < 1e-12:
        raise ValueError("Boundary arclength too small after cleaning; adjust alpha or input.")
    s_new = np.linspace(0.0, s[-1], int(target_n))
    return np.c_[np.interp(s_new, s, B[:, 0]), np.interp(s_new, s, B[:, 1])]


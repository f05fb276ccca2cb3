"""CPU: the C-ABI library loads, exports every symbol include/lm_b200.h declares, fails loudly
without a device, and the host-side logic (PNG/CSV writers) is correct; the ordering rule of the device
contour linker is pinned against the sequential mpl2014 restatement through its numpy restatement."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from helpers import lines_equal, link_records_numpy, records_from_dwell

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "lm_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(shim):
    lib = shim.load()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(names) == set(shim.EXPORTS), set(names) ^ set(shim.EXPORTS)
    assert shim.missing_exports() == []
    assert lib.lm_abi_version() == 2


def test_no_cpu_fallback(shim):
    """Without a GPU every compute entry point must fail with LM_E_NODEV, never compute on the CPU."""
    if shim.device_count() > 0:
        pytest.skip("a CUDA device is present")
    xs = np.linspace(-2, 1, 8); ys = np.linspace(-1, 1, 8)
    out = np.zeros((8, 8), dtype=np.int32)
    st = shim.Stats()
    rc = shim.load().lm_escape_grid_f64(shim.ptr(xs), 8, shim.ptr(ys), 8, 50, 2.0, 0, shim.ptr(out), None, None, C.byref(st))
    assert rc == shim.LM_E_NODEV
    assert "device" in shim.last_error().lower()
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import escape, lucas, stencils
    with pytest.raises(shim.LmError):
        escape.compute_grid((-2, 1), (-1, 1), 8, 50)
    with pytest.raises(shim.LmError):
        lucas.compute_inverse_eigenvalues(2, 5)
    with pytest.raises(shim.LmError):
        stencils.laplacian(np.zeros((4, 4)), 0.1)


def test_product_does_not_import_oracle():
    pkg = ROOT / "inverse_eigenvalue_loci_mandelbrot_correspondence_b200"
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        txt = f.read_text()
        assert "oracle" not in txt.replace("lm_oracle_contour.c / mpl2014", ""), f


@pytest.mark.parametrize("case", [(96, 500, (-2.1, 0.9), (-1.5, 1.5), 480.0), (120, 300, (-0.8, -0.7), (0.05, 0.15), 150.0),
                                  (75, 100, (-1.0, 0.5), (-0.3, 1.2), 3.0), (64, 60, (-0.755, -0.735), (0.10, 0.12), 57.6)])
def test_link_rule_matches_sequential_oracle(oracle, case):
    """The ordering rule of the device linker (helpers.link_records_numpy restates csrc/lm_contour_link.cu: open
    chains by head id, loops by smallest node id, list ranking) against the sequential mpl2014 restatement."""
    res, mi, xl, yl, lvl = case
    xs = np.linspace(*xl, res); ys = np.linspace(*yl, res + 7)
    d, _ = oracle.dwell_grid(xs, ys, mi)
    ref = oracle.contour_lines(xs, ys, d.astype(float), lvl)
    assert lines_equal(ref, link_records_numpy(records_from_dwell(d, xs, ys, lvl), xs, ys, lvl))


def test_link_rule_random_fields(oracle):
    rng = np.random.default_rng(1)
    nlines = 0
    for _ in range(150):
        ny, nx = rng.integers(2, 14), rng.integers(2, 14)
        d = rng.integers(0, 6, size=(ny, nx)).astype(np.int32)
        xs = np.sort(rng.uniform(-1, 1, nx)); ys = np.sort(rng.uniform(-1, 1, ny))
        lvl = float(rng.choice([1.5, 2.0, 2.5, 3.0]))
        ref = oracle.contour_lines(xs, ys, d.astype(float), lvl)
        assert lines_equal(ref, link_records_numpy(records_from_dwell(d, xs, ys, lvl), xs, ys, lvl))
        nlines += len(ref)
    assert nlines > 500


def test_oracle_uses_every_segment_once(oracle):
    """mpl2014 visits a saddle quad twice, once per segment.  On fields full of saddles the oracle's lines must use
    every segment of every crossed quad exactly once: no duplicated and no missing pieces (this is what pins the
    start edge of an already visited saddle: S/W when the first visit came in through N/E, else N/E)."""
    rng = np.random.default_rng(11)
    for _ in range(60):
        ny, nx = rng.integers(3, 20), rng.integers(3, 20)
        d = rng.integers(0, 3, size=(ny, nx)).astype(np.int32)
        xs = np.arange(nx, dtype=float); ys = np.arange(ny, dtype=float)
        lvl = 0.5
        lines = oracle.contour_lines(xs, ys, d.astype(float), lvl)
        nseg = int(((records_from_dwell(d, xs, ys, lvl)[:, 3] >> 16) & 3).sum())
        assert sum(len(ln) - 1 for ln in lines) == nseg
        pieces = set()
        for ln in lines:
            for a, b in zip(ln[:-1], ln[1:]):
                key = (tuple(a), tuple(b))
                assert key not in pieces
                pieces.add(key)


def test_contour_oracle_invariants(oracle):
    """Internal consistency of the (unpinned) mpl2014 restatement: vertices lie on grid edges that
    straddle the level, interior lines are closed, the longest line of the README window is closed."""
    xs = np.linspace(-2.1, 0.9, 150); ys = np.linspace(-1.5, 1.5, 150)
    d, _ = oracle.dwell_grid(xs, ys, 300)
    lvl = 0.96 * 300
    lines = oracle.contour_lines(xs, ys, d.astype(float), lvl)
    assert lines
    best = max(lines, key=len)
    assert np.array_equal(best[0], best[-1]) or np.allclose(best[0], best[-1], rtol=0, atol=1e-12)
    dx, dy = xs[1] - xs[0], ys[1] - ys[0]
    for ln in lines:
        fx = (ln[:, 0] - xs[0]) / dx; fy = (ln[:, 1] - ys[0]) / dy
        on_v = np.abs(fx - np.rint(fx)) < 1e-9
        on_h = np.abs(fy - np.rint(fy)) < 1e-9
        assert (on_v | on_h).all()
        step = np.hypot(np.diff(ln[:, 0]) / dx, np.diff(ln[:, 1]) / dy)
        assert (step <= np.sqrt(2) + 1e-9).all()


def test_png_and_csv_writers(tmp_path, shim):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import mandelbrot_boundary_sample as mbs
    t = np.linspace(0, 2 * np.pi, 200)
    contour = np.column_stack([np.cos(t), np.sin(t)])
    prefix = str(tmp_path / "out" / "mandel")
    csv, png, meta = mbs.save_outputs(contour, prefix, [-2.1, 0.9], [-1.5, 1.5], 2000, 500, 0.96)
    lines = Path(csv).read_text().splitlines()
    assert lines[0] == "x,y" and len(lines) == 201
    assert re.fullmatch(r"-?\d\.\d{18}e[+-]\d\d,-?\d\.\d{18}e[+-]\d\d", lines[1])
    back = np.loadtxt(csv, delimiter=",", skiprows=1)
    assert np.array_equal(back, contour)                       # %.18e round-trips binary64
    assert Path(png).read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"
    assert Path(meta).read_text() == "xlim=[-2.1, 0.9]\nylim=[-1.5, 1.5]\nres=2000\nmax_iter=500\nlevel=0.96\n"


def test_polylines_container():
    """contour.Polylines: the lines of a level as a lazy sequence of views (what cs.allsegs[0] is to the reference)."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200.contour import Polylines, longest
    verts = np.arange(20, dtype=np.float64).reshape(10, 2)
    L = Polylines(verts, np.array([0, 3, 3, 7, 10], dtype=np.int64))
    assert len(L) == 4 and [len(a) for a in L] == [3, 0, 4, 3]
    assert np.array_equal(L[2], verts[3:7]) and np.array_equal(L[-1], verts[7:10])
    assert [a.shape for a in L[1:3]] == [(0, 2), (4, 2)]
    with pytest.raises(IndexError):
        L[4]
    assert np.array_equal(L.longest(), verts[3:7]) and np.array_equal(longest(L), verts[3:7])
    assert np.array_equal(longest([verts[:2], verts[2:8], verts[8:]]), verts[2:8])     # plain lists still work
    E = Polylines(np.empty((0, 2)), np.zeros(1, dtype=np.int64))
    assert len(E) == 0 and not E and E.longest() is None and longest(E) is None and longest([]) is None
    # ties: the FIRST of the longest lines, like max(paths, key=len)
    T = Polylines(verts, np.array([0, 5, 10], dtype=np.int64))
    assert np.array_equal(T.longest(), verts[:5])


def test_new_row_host_logic_cpu(shim, golden):
    """Host-side pieces of the section-8f rows that need no device: scipy's gaussian kernel, the boundary walk of the alpha
    shape (against the reference's traced loops), argument checks; and that nothing computes without a GPU."""
    from scipy.ndimage import _filters
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import alpha_shape, pairstats, tracker
    for sigma in (0.3, 1.0, 2.5, 4.0):
        w, radius = tracker.gaussian_kernel1d(sigma)
        assert radius == int(4.0 * sigma + 0.5)
        assert np.array_equal(w, _filters._gaussian_kernel1d(sigma, 0, radius))
    P = golden["alpha_points"]
    for tag in ("a6", "a12"):
        edges = [tuple(e) for e in golden[f"alpha_{tag}_edges"].tolist()]
        assert np.array_equal(np.asarray(alpha_shape.order_boundary(P, edges), dtype=np.int32), golden[f"alpha_{tag}_ordered"])
    with pytest.raises(ValueError):
        pairstats.pair_histogram(np.zeros((4, 3)), [0.0], [1.0])
    with pytest.raises(ValueError):
        pairstats.pair_histogram(np.zeros((4, 2)), [0.0, 1.0], [1.0])
    with pytest.raises(ValueError):
        tracker.gaussian_filter(np.zeros((4, 4)), 1.0, mode="reflect")
    with pytest.raises(TypeError):
        tracker.gi_flow_fixed_T(lambda p, x: 0.0, np.ones((2, 2)), np.ones((2, 2)), 0.1, 2)
    assert tracker.fraction_outside_domain(np.array([0j, 3 + 0j, 1 + 1j, -5j]), (-2.2, 1.2, -1.6, 1.6)) == 0.5
    if shim.device_count() < 1:
        mod = type("M", (), {"domain": (-2.2, 1.2, -1.6, 1.6), "eps": 1e-12})
        for call in (lambda: pairstats.pair_histogram(np.zeros((4, 2)), [0.0], [1.0]),
                     lambda: pairstats.max_pair_distance(np.zeros((4, 2))),
                     lambda: tracker.mollified_histogram(mod, np.zeros(4, complex), 8, 1.0),
                     lambda: tracker.KL(np.ones(4) / 4, np.ones(4) / 4),
                     lambda: tracker.gi_flow_fixed_T(tracker.KL, np.ones(4) / 4, np.ones(4) / 4, 0.1, 2),
                     lambda: tracker.sum_pairwise(np.ones(9)),
                     lambda: alpha_shape.alpha_shape_edges(np.zeros((3, 2)), 1.0, simplices=[[0, 1, 2]])):
            with pytest.raises(RuntimeError):
                call()

"""GPU: K2 (level set) through the C ABI against the mpl2014 restatement (bit-exact vertices, same
line order) and the full drop-in script."""
from pathlib import Path

import numpy as np
import pytest

from helpers import lines_equal, records_from_dwell

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("res,mi,xlim,ylim,frac", [
    (400, 500, (-2.1, 0.9), (-1.5, 1.5), 0.96),
    (513, 300, (-2.1, 0.9), (-1.5, 1.5), 0.5),
    (300, 400, (-0.755, -0.735), (0.10, 0.12), 0.96),      # lines cut by the window border, many saddles
    (257, 200, (-1.0, 0.5), (-0.3, 1.2), 0.02),
    (131, 64, (-2.0, 1.0), (-1.5, 1.5), 0.96),
])
def test_contour_vs_oracle(gpu, oracle, res, mi, xlim, ylim, frac):
    xs = np.linspace(*xlim, res); ys = np.linspace(*ylim, res + 5)
    d, _ = oracle.dwell_grid(xs, ys, mi)
    lvl = frac * mi
    want = oracle.contour_lines(xs, ys, d.astype(float), lvl)
    got = gpu.contour.contour_lines(xs, ys, d, lvl)
    assert lines_equal(want, got)
    best = gpu.contour.extract_contour(xs, ys, d.astype(float), mi, frac)
    ref_best = oracle.extract_contour(xs, ys, d.astype(float), mi, frac)
    assert (best is None and ref_best is None) or np.array_equal(best, ref_best)


def test_records_match_numpy_restatement(gpu, oracle):
    import ctypes as C
    xs = np.linspace(-2.1, 0.9, 300); ys = np.linspace(-1.5, 1.5, 200)
    d, _ = oracle.dwell_grid(xs, ys, 300)
    lvl = 288.0
    want = records_from_dwell(d, xs, ys, lvl)
    with gpu.device.DeviceGrid(xs, ys) as g:
        g.escape(300)
        recs = np.empty((len(want) + 8, 8), dtype=np.int64)
        n = C.c_int64(0)
        gpu.shim.call("lm_contour_classify_dev", C.c_void_p(g.d_dwell.ptr), gpu.shim.ptr(xs), xs.size, gpu.shim.ptr(ys), ys.size,
                      0, lvl, gpu.shim.ptr(recs), recs.shape[0], C.byref(n), None)
    assert n.value == len(want)
    assert np.array_equal(recs[: n.value], want)


@pytest.mark.parametrize("case", [(96, 500, (-2.1, 0.9), (-1.5, 1.5), 480.0), (120, 300, (-0.8, -0.7), (0.05, 0.15), 150.0),
                                  (75, 100, (-1.0, 0.5), (-0.3, 1.2), 3.0), (64, 60, (-0.755, -0.735), (0.10, 0.12), 57.6)])
def test_contour_linker_matches_oracle(gpu, oracle, case):
    """lm_contour_link (records uploaded, chained on the device) against the dense mpl2014 restatement."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import contour
    res, mi, xl, yl, lvl = case
    xs = np.linspace(*xl, res); ys = np.linspace(*yl, res + 7)
    d, _ = oracle.dwell_grid(xs, ys, mi)
    ref = oracle.contour_lines(xs, ys, d.astype(float), lvl)
    got = contour.link_records(records_from_dwell(d, xs, ys, lvl), xs, ys, lvl)
    assert lines_equal(ref, got)


def test_contour_linker_random_fields(gpu, oracle):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import contour
    rng = np.random.default_rng(1)
    for _ in range(150):
        ny, nx = rng.integers(2, 14), rng.integers(2, 14)
        d = rng.integers(0, 6, size=(ny, nx)).astype(np.int32)
        xs = np.sort(rng.uniform(-1, 1, nx)); ys = np.sort(rng.uniform(-1, 1, ny))
        lvl = float(rng.choice([1.5, 2.0, 2.5, 3.0]))
        ref = oracle.contour_lines(xs, ys, d.astype(float), lvl)
        got = contour.link_records(records_from_dwell(d, xs, ys, lvl), xs, ys, lvl)
        assert lines_equal(ref, got)


def test_contour_linker_split_blocks(gpu, oracle):
    """Records of two row blocks (one halo row each side of the seam) concatenate to the single-block result."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import contour
    xs = np.linspace(-2.1, 0.9, 90); ys = np.linspace(-1.5, 1.5, 80)
    d, _ = oracle.dwell_grid(xs, ys, 200)
    lvl = 192.0
    whole = records_from_dwell(d, xs, ys, lvl)
    cut = 37
    top = records_from_dwell(d[: cut + 1], xs, ys[: cut + 1], lvl, row_offset=0)
    bot = records_from_dwell(d[cut:], xs, ys[cut:], lvl, row_offset=cut)
    both = np.concatenate([top, bot])
    assert np.array_equal(whole, both)
    assert lines_equal(contour.link_records(both, xs, ys, lvl), oracle.contour_lines(xs, ys, d.astype(float), lvl))




def test_contour_linker_rejects_inconsistent_records(gpu, oracle):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import contour
    xs = np.linspace(-2.1, 0.9, 90); ys = np.linspace(-1.5, 1.5, 80)
    d, _ = oracle.dwell_grid(xs, ys, 200)
    recs = records_from_dwell(d, xs, ys, 192.0)
    assert len(recs) > 50
    with pytest.raises(ValueError, match="inconsistent"):
        contour.link_records(np.delete(recs, len(recs) // 2, axis=0), xs, ys, 192.0)      # a neighbour quad is missing
    with pytest.raises(ValueError, match="inconsistent"):
        contour.link_records(recs[::-1], xs, ys, 192.0)                                     # not in raster order
    assert len(contour.link_records(recs[:0], xs, ys, 192.0)) == 0


def test_contour_link_dev_many_small_loops(gpu, oracle):
    """A checkerboard-like field: thousands of 4-node loops and saddles, lines cut by every border."""
    rng = np.random.default_rng(7)
    d = rng.integers(0, 2, size=(257, 300)).astype(np.int32)
    xs = np.linspace(0, 1, 300); ys = np.linspace(0, 1, 257)
    want = oracle.contour_lines(xs, ys, d.astype(float), 0.5)
    got = gpu.contour.contour_lines(xs, ys, d, 0.5)
    assert len(want) > 3000
    assert lines_equal(want, got)


@pytest.mark.parametrize("shape", [(2, 2), (2, 300), (300, 2), (3, 129), (5, 130),
                                   # widths that are multiples of 4 take the bulk-copy mark kernel: tile / strip / row-group edges
                                   (2, 4), (3, 8), (6, 1024), (9, 1028), (5, 1032), (7, 2052), (4, 132), (11, 128), (13, 2048)])
def test_contour_small_and_ragged(gpu, oracle, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    ny, nx = shape
    d = rng.integers(0, 5, size=(ny, nx)).astype(np.int32)
    xs = np.linspace(0, 1, nx); ys = np.linspace(0, 1, ny)
    for lvl in (0.5, 2.0, 3.5, 10.0):
        assert lines_equal(oracle.contour_lines(xs, ys, d.astype(float), lvl), gpu.contour.contour_lines(xs, ys, d, lvl))


def test_no_contour_returns_none(gpu):
    xs = np.linspace(2.5, 3.0, 50); ys = np.linspace(2.5, 3.0, 50)     # everything escapes at once
    _, _, Z = gpu.escape.compute_grid((2.5, 3.0), (2.5, 3.0), 50, 100)
    assert gpu.contour.extract_contour(xs, ys, Z, 100, 0.96) is None
    with pytest.raises(ValueError):
        gpu.contour.contour_lines(xs, ys, Z + 0.5, 1.0)


def test_device_resident_pipeline_config1(gpu, oracle):
    """BASELINE config 1 at full size: K1 -> K2 without the dwell grid leaving the GPU."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import mandelbrot_boundary_sample as mbs
    xs, ys, best = mbs.boundary_from_window((-2.1, 0.9), (-1.5, 1.5), 2000, 500, 0.96)
    d, _ = oracle.dwell_grid(xs, ys, 500)
    want = oracle.extract_contour(xs, ys, d.astype(float), 500, 0.96)
    assert np.array_equal(best, want)
    assert best.shape[0] > 10000 and np.array_equal(best[0], best[-1])


def test_script_cli_outputs(gpu, oracle, tmp_path):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import mandelbrot_boundary_sample as mbs
    prefix = str(tmp_path / "outputs" / "mandel")
    mbs.main(["--xlim", "-2.1", "0.9", "--ylim", "-1.5", "1.5", "--res", "600", "--max_iter", "300",
              "--level", "0.96", "--output_prefix", prefix])
    csv = Path(prefix + "_boundary.csv").read_text().splitlines()
    assert csv[0] == "x,y"
    got = np.loadtxt(prefix + "_boundary.csv", delimiter=",", skiprows=1)
    xs = np.linspace(-2.1, 0.9, 600); ys = np.linspace(-1.5, 1.5, 600)
    d, _ = oracle.dwell_grid(xs, ys, 300)
    want = oracle.extract_contour(xs, ys, d.astype(float), 300, 0.96)
    assert np.array_equal(got, want)
    assert Path(prefix + "_boundary.png").read_bytes()[:4] == b"\x89PNG"
    assert Path(prefix + "_meta.txt").read_text() == "xlim=[-2.1, 0.9]\nylim=[-1.5, 1.5]\nres=600\nmax_iter=300\nlevel=0.96\n"
    with pytest.raises(SystemExit) as e:
        mbs.main(["--xlim", "2.5", "3.0", "--ylim", "2.5", "3.0", "--res", "64", "--max_iter", "50", "--output_prefix", prefix])
    assert "Failed to extract a usable contour" in str(e.value)


def test_fused_boundary_sample(gpu, oracle):
    """lm_boundary_sample == compute_grid followed by extract_contour, dwell optionally returned."""
    xs = np.linspace(-2.1, 0.9, 700); ys = np.linspace(-1.5, 1.5, 500)
    d_ref, work = oracle.dwell_grid(xs, ys, 400)
    want = oracle.contour_lines(xs, ys, d_ref.astype(float), 384.0)
    lines, st = gpu.contour.boundary_sample(xs, ys, 400, 384.0)
    assert lines_equal(want, lines) and st["work_units"] == work
    for dt in (np.int32, np.float64):
        out = np.full((500, 700), -7, dtype=dt)
        lines, _ = gpu.contour.boundary_sample(xs, ys, 400, 384.0, dwell_out=out)
        assert lines_equal(want, lines) and np.array_equal(out, d_ref)
    # large enough for several row chunks (> 256 MB of output): pinned int32 output, checked by symmetry + strip
    res = 9000
    xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res)
    out = gpu.shim.pinned_empty((res, res), np.int32)
    lines, st = gpu.contour.boundary_sample(xs, ys, 200, 192.0, dwell_out=out)
    strip, _, _ = gpu.escape.escape_grid(xs, ys[4500:4508], 200)
    assert np.array_equal(out[4500:4508], strip)
    assert st["work_units"] == int(np.minimum(out.astype(np.int64) + 1, 200).sum())
    best = gpu.contour.longest(lines)
    assert np.array_equal(best[0], best[-1]) and len(best) > 20000


def test_shard_escape_keeps_block_for_halo_exchange(gpu):
    """The N > 1 e2e path on one device: two row shards computed through lm_shard_escape (host copy + HBM-resident
    block with a halo slot), the lower shard's halo filled with the upper shard's first row, K2 per shard from the
    resident block, records concatenated and linked -> identical to the single-call result."""
    import ctypes as C
    from helpers import lines_equal
    shim = gpu.shim
    xs = np.linspace(-2.1, 0.9, 700); ys = np.linspace(-1.5, 1.5, 640)
    mi, lvl = 300, 0.96 * 300
    want_lines, _ = gpu.contour.boundary_sample(xs, ys, mi, lvl)
    full, _, _ = gpu.escape.escape_grid(xs, ys, mi)
    cut = 290
    recs = []
    upper_first = None
    for (r0, r1, has_halo) in ((cut, ys.size, False), (0, cut, True)):          # upper shard first: it owns the halo row
        rows = r1 - r0
        out = np.empty((rows, xs.size), dtype=np.int32)
        blk = C.c_void_p()
        st = shim.Stats()
        ys_rows = np.ascontiguousarray(ys[r0:r1])
        pot = np.empty((rows, xs.size)); pblk = C.c_void_p()
        cost = np.linspace(1.0, 3.0, rows) if has_halo else None        # a row-cost hint only reorders the chunks
        shim.call("lm_shard_escape", shim.ptr(xs), xs.size, shim.ptr(ys_rows), rows, mi, shim.ptr(out), shim.ptr(pot), 1,
                  shim.ptr(cost), C.byref(blk), C.byref(pblk), C.byref(st))
        assert np.array_equal(out, full[r0:r1]) and blk.value and pblk.value
        back = np.empty_like(pot)
        shim.call("lm_memcpy_d2h", shim.ptr(back), pblk, back.nbytes, None); shim.call("lm_stream_synchronize", None)
        assert np.array_equal(back, pot) and np.array_equal(pot, gpu.escape.escape_grid(xs, ys_rows, mi, field_mode=1)[1])
        assert st.work_units == int(np.minimum(out.astype(np.int64) + 1, mi).sum())
        if not has_halo:
            upper_first = out[0].copy()
        else:
            shim.call("lm_memcpy_h2d", C.c_void_p(blk.value + rows * xs.size * 4), shim.ptr(upper_first), upper_first.nbytes, None)
        ys_blk = np.ascontiguousarray(ys[r0:r1 + (1 if has_halo else 0)])
        cap = 1 << 16
        buf = np.empty((cap, 8), dtype=np.int64); n = C.c_int64(0)
        shim.call("lm_contour_classify_dev", blk, shim.ptr(xs), xs.size, shim.ptr(ys_blk), ys_blk.size, r0, float(lvl),
                  shim.ptr(buf), cap, C.byref(n), None)
        recs.append((r0, buf[: n.value].copy()))
    allrec = np.concatenate([r for _, r in sorted(recs, key=lambda t: t[0])])
    got = gpu.contour.link_records(allrec, xs, ys, lvl)
    assert lines_equal(got, want_lines)


def test_boundary_sample_with_potential(gpu, oracle):
    """config 2's stage in one call: dwell grid + smooth potential + ordered boundary from one K1 pass."""
    from helpers import lines_equal
    xs = np.linspace(-2.1, 0.9, 333); ys = np.linspace(-1.5, 1.5, 301)
    mi = 400
    d = np.empty((ys.size, xs.size), dtype=np.int32); g = np.empty((ys.size, xs.size))
    lines, st = gpu.contour.boundary_sample(xs, ys, mi, 0.96 * mi, dwell_out=d, potential_out=g)
    d_o, g_o = oracle.potential_grid(xs, ys, mi, 2.0, oracle.FIELD_GREEN)
    assert np.array_equal(d, d_o)
    np.testing.assert_allclose(g, g_o, rtol=1e-14, atol=0)
    assert lines_equal(lines, oracle.contour_lines(xs, ys, d_o.astype(float), 0.96 * mi))
    plain, _ = gpu.contour.boundary_sample(xs, ys, mi, 0.96 * mi)
    assert lines_equal(lines, plain)
    assert gpu.contour.longest(lines).shape == oracle.extract_contour(xs, ys, d_o.astype(float), mi, 0.96).shape


def test_cost_profile_device_equals_host(gpu):
    """sharding.coarse_row_profile: the device-resident path (exact int64 sums) and the host-buffer / numpy path give the
    same per-row cost estimate, so every rank derives the same cuts whichever it uses."""
    import torch
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    xs = np.linspace(-2.1, 0.9, 3000); ys = np.linspace(-1.5, 1.5, 2500)
    dev = torch.device("cuda", 0)
    a = sharding.coarse_row_profile(xs, ys, 700, rows=300, cols=400)
    b = sharding.coarse_row_profile(xs, ys, 700, rows=300, cols=400, device=dev)
    assert np.array_equal(a, b)
    plan = sharding.plan_row_cuts(xs, ys, 700, 4, device=dev)
    assert plan["cuts"][0] == 0 and plan["cuts"][-1] == ys.size and len(plan["cuts"]) == 5
    assert plan["balance_estimate"] > 0.99

"""CPU: the N > 1 host logic with torch.distributed (gloo, world_size 2): work-balanced row cuts,
first-row all-gather for the K2 halo, record gather + global linking.  The per-rank compute is
stood in by the oracle (dwell) and the numpy record restatement; the collectives and the linker
are the product's."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_balanced_row_cuts_properties():
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    rng = np.random.default_rng(0)
    for _ in range(50):
        ny = int(rng.integers(8, 400)); parts = int(rng.integers(1, 9))
        w = rng.random(ny) ** 8 * 1000 + (rng.random(ny) < 0.3) * 5000
        cuts = sharding.balanced_row_cuts(w, parts)
        assert cuts[0] == 0 and cuts[-1] == ny and len(cuts) == parts + 1
        assert all(b > a for a, b in zip(cuts[:-1], cuts[1:]))
    with pytest.raises(ValueError):
        sharding.balanced_row_cuts(np.ones(3), 4)
    assert sharding.balanced_row_cuts(np.ones(8), 4) == [0, 2, 4, 6, 8]


def test_work_balanced_beats_equal_rows(oracle):
    """equal-rows sharding is ~50 % / ~37 % efficient at 4 / 8 shards, work-balanced >= 95 %."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    xs = np.linspace(-2.1, 0.9, 256); ys = np.linspace(-1.5, 1.5, 512)
    d, _ = oracle.dwell_grid(xs, ys, 2000)
    w = np.minimum(d.astype(np.int64) + 1, 2000).sum(axis=1)
    for parts, eq_max in ((4, 0.60), (8, 0.45)):
        equal = [round(k * 512 / parts) for k in range(parts + 1)]
        cuts = sharding.balanced_row_cuts(w, parts)
        assert sharding.parallel_efficiency(w, equal) < eq_max
        assert sharding.parallel_efficiency(w, cuts) > 0.95
    # a coarse profile (every 8th row) interpolated to all rows still balances well
    prof = sharding.interpolate_row_profile(np.arange(0, 512, 8), w[::8], 512)
    assert sharding.parallel_efficiency(w, sharding.balanced_row_cuts(prof, 8)) > 0.9


def test_refine_cuts_from_measurements():
    """Measured rebalancing: a hidden per-row overhead the estimate does not see is recovered from block timings."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    rng = np.random.default_rng(1)
    est = rng.random(4000) ** 6 * 1000 + 1.0
    true = est + 40.0                                   # constant per-row cost missing from the estimate
    cuts = sharding.balanced_row_cuts(est, 8)
    before = sharding.parallel_efficiency(true, cuts)
    for _ in range(2):
        measured = [true[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])]
        cuts = sharding.refine_cuts(est, cuts, measured)
    after = sharding.parallel_efficiency(true, cuts)
    assert after > before and after > 0.97
    assert cuts[0] == 0 and cuts[-1] == 4000 and all(b > a for a, b in zip(cuts[:-1], cuts[1:]))


def _free_port() -> int:
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank: int, world: int, port: int, q):
    try:
        sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        from helpers import lines_equal, link_records_numpy, records_from_dwell
        from oracle import oracle
        from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
        dist.init_process_group("gloo", rank=rank, world_size=world)
        xs = np.linspace(-2.1, 0.9, 140); ys = np.linspace(-1.5, 1.5, 120)
        mi, lvl = 200, 192.0
        # every rank derives the same cuts from the same (deterministic) coarse profile
        coarse_rows = np.arange(0, ys.size, 4)
        dc, _ = oracle.dwell_grid(xs[::2], ys[coarse_rows], mi)
        # the product's plan on the CPU: per-pixel cost model on the coarse dwell samples, the linking rank cut lighter
        prof = sharding.interpolate_row_profile(coarse_rows, sharding.pixel_cost(dc, mi).sum(axis=1) * 2.0, ys.size)
        extra = np.zeros(world); extra[0] = 0.02 * prof.sum() / world
        cuts = sharding.balanced_row_cuts(prof, world, extra)
        assert sharding.parallel_efficiency(prof, cuts, extra) > 0.9
        r0, r1 = cuts[rank], cuts[rank + 1]
        mine, _ = oracle.dwell_grid(xs, ys[r0:r1], mi)                       # this rank's K1 rows
        firsts = sharding.exchange_first_rows(torch.from_numpy(mine[0].copy()))
        assert firsts.shape == (world, xs.size)
        has_halo = rank < world - 1
        block = np.vstack([mine, firsts[rank + 1].numpy()[None, :]]) if has_halo else mine
        recs = records_from_dwell(block, xs, ys[r0:r1 + (1 if has_halo else 0)], lvl, row_offset=r0, nx_global=xs.size)
        rec_t = torch.zeros((len(recs) + 5, 8), dtype=torch.int64)          # capacity > count, like the device buffer
        rec_t[: len(recs)] = torch.from_numpy(recs)
        allrec, counts = sharding.gather_records(rec_t, len(recs), 0)
        ok = len(counts) == world and counts[rank] == len(recs)
        if rank == 0:
            full, _ = oracle.dwell_grid(xs, ys, mi)
            ok = ok and np.array_equal(allrec.numpy(), records_from_dwell(full, xs, ys, lvl))
            lines = link_records_numpy(allrec.numpy(), xs, ys, lvl)          # restates the device linker (no GPU here)
            ok = ok and lines_equal(lines, oracle.contour_lines(xs, ys, full.astype(float), lvl))
        else:
            ok = ok and allrec is None
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:          # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_boundary_gloo(shim, oracle, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in results:
        assert ok, f"rank {rank}: {err}"


def _cloud_worker(rank: int, world: int, port: int, q):
    try:
        sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        from oracle import oracle
        from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
        dist.init_process_group("gloo", rank=rank, world_size=world)
        # config-5 shaped batch; every rank owns a contiguous slice of the polynomials
        rng = np.random.default_rng(0)
        npoly = 41
        deg = rng.integers(2, 9, size=npoly)
        tops = [np.where(np.arange(d) == d - 1, 1.0, rng.integers(0, 3, size=d).astype(float)) for d in deg]
        cuts = sharding.item_slices(npoly, world)
        mine = [oracle.inverse_eigenvalues_toprow(tops[k], 1e-12) for k in range(cuts[rank], cuts[rank + 1])]
        cloud = np.concatenate(mine) if mine else np.zeros(0, complex)
        gx = np.linspace(-2, 2, 17); gy = np.linspace(-2, 2, 13)
        # raw per-cell sums of this rank's slice (what lm_log_potential_sums_dev produces on the GPU)
        pts = np.column_stack([cloud.real, cloud.imag])
        sums = oracle.log_potential(pts, gx, gy, 1e-12, 0) * max(len(cloud), 1) if len(cloud) else np.zeros((13, 17))
        t, n_total = sharding.allreduce_field_sums(torch.from_numpy(sums.copy()), len(cloud))
        U = t.numpy() / n_total
        full = np.concatenate([oracle.inverse_eigenvalues_toprow(tp, 1e-12) for tp in tops])
        want = oracle.log_potential(np.column_stack([full.real, full.imag]), gx, gy, 1e-12, 0)
        ok = n_total == full.size and np.allclose(U, want, rtol=1e-12, atol=1e-13)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception:          # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


def test_item_slices():
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    assert sharding.item_slices(10, 3) == [0, 3, 6, 10]
    assert sharding.item_slices(2, 4) == [0, 0, 1, 1, 2]
    c = sharding.item_slices(10 ** 7, 8)
    assert c[0] == 0 and c[-1] == 10 ** 7 and max(np.diff(c)) - min(np.diff(c)) <= 1


@pytest.mark.parametrize("world", [2])
def test_sharded_cloud_field_gloo(oracle, world):
    """K4a with the cloud sharded over the ranks: all-reduce(sum) of the raw per-cell sums and of the cloud
    sizes, then the 1/N normalisation -- equals the single-process field."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cloud_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in results:
        assert ok, f"rank {rank}: {err}"


def _field_worker(rank: int, world: int, port: int, q):
    try:
        sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        from oracle import oracle
        from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
        dist.init_process_group("gloo", rank=rank, world_size=world)
        xs = np.linspace(-2.1, 0.9, 90); ys = np.linspace(-1.5, 1.5, 77)
        cuts = sharding.balanced_row_cuts(np.ones(ys.size) + (np.arange(ys.size) % 7), world)     # ragged blocks
        r0, r1 = cuts[rank], cuts[rank + 1]
        _, mine = oracle.potential_grid(xs, ys[r0:r1], 120, 2.0, oracle.FIELD_GREEN)       # this rank's K1 field rows
        _, full = oracle.potential_grid(xs, ys, 120, 2.0, oracle.FIELD_GREEN)
        block = torch.from_numpy(mine.copy())
        got = sharding.allgather_rows(block, cuts).numpy()
        ok = np.array_equal(got, full)
        h = float(xs[1] - xs[0])
        # periodic Laplacian of the sharded field through halo rows == rows of the Laplacian of the full field
        above, below = sharding.exchange_halo_rows(block, periodic=True)
        ext = np.vstack([above.numpy()[None], mine, below.numpy()[None]])
        ok = ok and np.array_equal(oracle.laplacian(ext, h)[1:-1], oracle.laplacian(full, h)[r0:r1])
        # interior 5-point average: the outer ranks have no halo on the outside (copied border)
        above, below = sharding.exchange_halo_rows(block, periodic=False)
        parts = ([above.numpy()[None]] if above is not None else []) + [mine] + ([below.numpy()[None]] if below is not None else [])
        sm = oracle.smooth5(np.vstack(parts))
        sm = sm[(1 if above is not None else 0): sm.shape[0] - (1 if below is not None else 0)]
        ok = ok and np.array_equal(sm, oracle.smooth5(full)[r0:r1])
        ok = ok and (above is None) == (rank == 0) and (below is None) == (rank == world - 1)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception:          # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_potential_field_gloo(oracle, world):
    """The final potential field of a row-sharded grid: all-gather of ragged row blocks, and the halo rows that
    let each rank apply the 5-point stencils to its own block (periodic wrap across the first / last rank)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_field_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in results:
        assert ok, f"rank {rank}: {err}"


def test_balanced_row_cuts_with_a_serial_stage():
    """The rank that also links the records gets less row work: block cost + extra is what is balanced."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    rng = np.random.default_rng(3)
    w = rng.random(5000) + 0.05
    extra = [40.0, 0.0, 0.0, 0.0, 0.0]
    cuts = sharding.balanced_row_cuts(w, 5, extra)
    tot = np.array([w[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])]) + np.array(extra)
    assert cuts[0] == 0 and cuts[-1] == 5000 and all(b > a for a, b in zip(cuts[:-1], cuts[1:]))
    assert tot.max() / tot.mean() < 1.002
    assert sharding.parallel_efficiency(w, cuts, extra) > 0.998
    assert sharding.balanced_row_cuts(w, 5, None) == sharding.balanced_row_cuts(w, 5, [0.0] * 5)
    huge = sharding.balanced_row_cuts(w, 5, [1e12, 0, 0, 0, 0])            # capped: the block keeps half a share
    first = w[: huge[1]].sum()
    assert 0.35 * w.sum() / 5 < first < 0.65 * w.sum() / 5


def test_pixel_cost_model():
    """cost = it + b + d_late * [escaped and it > knee], it = min(dwell + 1, max_iter)."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding
    m = sharding.COST_MODEL_B200
    d = np.array([0, 3, 62, 63, 64, 500, 999, 1000])
    c = sharding.pixel_cost(d, 1000)
    it = np.minimum(d + 1, 1000)
    late = (d < 1000) & (it > m["knee"])
    assert np.array_equal(c, it + m["b_pixel"] + m["d_late"] * late)
    assert list(late) == [False, False, False, False, True, True, True, False]      # interior pixels are not "late"
    assert sharding.pixel_cost(np.array([1000]), 1000)[0] == 1000 + m["b_pixel"]
    custom = {"b_pixel": 1.0, "d_late": 0.0, "knee": 64}
    assert np.array_equal(sharding.pixel_cost(d, 1000, custom), it + 1.0)
    # the one-rank plan needs no device
    plan = sharding.plan_row_cuts(np.linspace(-2, 1, 10), np.linspace(-1, 1, 8), 50, 1)
    assert plan["cuts"] == [0, 8] and plan["balance_estimate"] == 1.0

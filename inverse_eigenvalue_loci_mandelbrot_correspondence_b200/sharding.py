"""Row sharding of the escape-time grid over the GPUs of one node (one process per GPU).

Rows of the grid are independent for K1; K2 needs one halo row from the next shard.  The
reference is single-process, so this layer has no reference counterpart:
it only decides which rows a rank computes and moves the two small things that cross ranks,
  * the first dwell row of every shard (all-gather, nx*4 bytes per rank) -> halo for K2,
  * the per-shard crossing records (gather to rank 0) -> one global linking pass.
Row work varies by three orders of magnitude across the window, so shards are contiguous row
blocks cut at equal ESTIMATED work (a coarse K1 pre-pass), not equal row counts.

The collective helpers take torch tensors and work with any torch.distributed backend (NCCL on
the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def balanced_row_cuts(row_work, nparts: int, extra=None) -> list[int]:
    """Cut rows [0, ny) into `nparts` contiguous blocks of (nearly) equal cumulative work.

    extra: optional fixed cost per block in the same units (e.g. the serial linking stage on the rank that also
    chains the records): block k then receives row work T - extra[k], with T the common total per block.
    Returns nparts+1 increasing cut positions starting at 0 and ending at ny; every block has at
    least one row when ny >= nparts.
    """
    w = np.asarray(row_work, dtype=np.float64).ravel()
    ny = w.size
    if nparts < 1:
        raise ValueError("nparts must be >= 1")
    if ny < nparts:
        raise ValueError(f"cannot cut {ny} rows into {nparts} non-empty blocks")
    w = np.maximum(w, 0.0) + 1e-12 * max(float(w.max(initial=0.0)), 1.0)     # keep the prefix strictly increasing
    cum = np.concatenate([[0.0], np.cumsum(w)])
    ex = np.zeros(nparts) if extra is None else np.asarray(extra, dtype=np.float64).ravel()
    if ex.size != nparts:
        raise ValueError("one extra cost per block expected")
    ex = np.clip(ex, 0.0, 0.5 * cum[-1] / nparts)        # a block keeps at least half of an even share of the row work
    per_block = (cum[-1] + ex.sum()) / nparts
    cuts = [0]
    target = 0.0
    for k in range(1, nparts):
        target += per_block - ex[k - 1]
        c = int(np.searchsorted(cum, target, side="left"))
        if c > 0 and abs(cum[c - 1] - target) <= abs(cum[min(c, ny)] - target):
            c -= 1
        c = max(c, cuts[-1] + 1)              # non-empty block
        c = min(c, ny - (nparts - k))         # leave rows for the remaining blocks
        cuts.append(c)
    cuts.append(ny)
    return cuts


def parallel_efficiency(row_work, cuts, extra=None) -> float:
    """mean block cost / max block cost for the given cuts (1.0 = perfectly balanced); extra as in balanced_row_cuts."""
    w = np.asarray(row_work, dtype=np.float64).ravel()
    blocks = np.array([w[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])])
    if extra is not None:
        blocks = blocks + np.asarray(extra, dtype=np.float64).ravel()
    return float(np.mean(blocks) / max(np.max(blocks), 1e-300))


def interpolate_row_profile(coarse_rows: np.ndarray, coarse_work: np.ndarray, ny: int) -> np.ndarray:
    """Per-row work estimate for all ny rows from the work of a subset of rows."""
    return np.interp(np.arange(ny, dtype=np.float64), np.asarray(coarse_rows, dtype=np.float64),
                     np.asarray(coarse_work, dtype=np.float64))


# K1 cost of one pixel on a B200, in units of one blind-path pixel-iteration, as a function of its iteration count
# `it` = min(dwell + 1, max_iter) and of whether it escaped:
#     cost = it + B_PIXEL + D_LATE * [escaped and it > KNEE]
# B_PIXEL: refill (ballot / popc rank, coordinate loads), the first careful iterations, staging and the store of the
# result.  D_LATE: a pixel that escapes inside a blind block sends its lane back over that block with the exact test
# while the other 31 lanes of the warp wait -- about 64 warp-iterations, i.e. ~2000 lane-iterations per event.
# The constants were fitted ONCE, offline, to CUDA-event timings of 2 x 48 row bands of BASELINE.json configs 3 and 4 on
# a B200 (joint least squares, rms error 0.9 %, max 3.4 %; scripts/balance_probe.py, profiles/r02_balance_*.json):
# they are properties of the kernel and the device, not of the workload, so one-shot cuts need no calibration pass
# over the answer.
COST_MODEL_B200 = {"b_pixel": 15.0, "d_late": 2000.0, "knee": 64,
                   # absolute scale of one cost unit (2.96e12 blind pixel-iterations per second on a B200) and the
                   # serial stage the linking rank runs after its K1 while the others wait for it in the next step:
                   # record gather (NCCL recv from every rank) + device linker + its two host syncs, ~1.3 ms for the
                   # 5.5e5 records of config 3 at N = 8 (`sub_rooflines.k2_records` of the bench line)
                   "seconds_per_unit": 3.4e-13, "link_seconds": 1.3e-3}


def pixel_cost(dwell, max_iter: int, model: dict | None = None) -> np.ndarray:
    """Estimated K1 cost (blind pixel-iteration units) of pixels with the given dwell values."""
    m = COST_MODEL_B200 if model is None else model
    d = np.asarray(dwell, dtype=np.int64)
    it = np.minimum(d + 1, max_iter).astype(np.float64)
    late = (d < max_iter) & (it > m["knee"])
    return it + m["b_pixel"] + m["d_late"] * late


def _sample_indices(n: int, k: int) -> np.ndarray:
    return np.unique(np.linspace(0, n - 1, min(k, n)).round().astype(np.int64))


def coarse_row_profile(xs, ys, max_iter: int, rows: int = 2048, cols: int = 2048, model: dict | None = None,
                       iterations_only: bool = False, device=None) -> np.ndarray:
    """Estimated K1 cost of every grid row from a coarse K1 pass on the GPU (subsampled rows / columns), weighted by
    the per-pixel cost model (iterations_only=True: round 1's profile, the bare iteration counts).
    With `device` (a torch CUDA device) the sampled dwell grid never leaves the GPU: K1 through the device-resident
    entry point, the per-row cost sums in exact int64 arithmetic (identical on every rank), 16 KB back to the host;
    without it the host-buffer API and numpy are used."""
    xs = np.asarray(xs, dtype=np.float64); ys = np.asarray(ys, dtype=np.float64)
    ri = _sample_indices(ys.size, rows)
    ci = _sample_indices(xs.size, cols)
    m = COST_MODEL_B200 if model is None else model
    if device is not None:
        import ctypes as C
        import torch
        from . import _shim
        xd = torch.from_numpy(np.ascontiguousarray(xs[ci])).to(device)
        yd = torch.from_numpy(np.ascontiguousarray(ys[ri])).to(device)
        d = torch.empty((ri.size, ci.size), dtype=torch.int32, device=device)
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _shim.call("lm_escape_grid_f64_dev", C.c_void_p(xd.data_ptr()), ci.size, C.c_void_p(yd.data_ptr()), ri.size,
                   int(max_iter), 2.0, 0, C.c_void_p(d.data_ptr()), None, None, None, stream)
        it = torch.clamp(d.to(torch.int64) + 1, max=int(max_iter))
        if iterations_only:
            work = it.sum(dim=1)
        else:
            late = ((d < int(max_iter)) & (it > int(m["knee"]))).to(torch.int64)
            work = (it + int(round(m["b_pixel"])) + int(round(m["d_late"])) * late).sum(dim=1)
        work = work.cpu().numpy().astype(np.float64)
    else:
        from . import escape
        d, _, _ = escape.escape_grid(xs[ci], ys[ri], max_iter)
        if iterations_only:
            work = np.minimum(d.astype(np.int64) + 1, max_iter).sum(axis=1).astype(np.float64)
        else:
            work = pixel_cost(d, max_iter, m).sum(axis=1)
    return interpolate_row_profile(ri, work * (xs.size / ci.size), ys.size)


def plan_row_cuts(xs, ys, max_iter: int, nparts: int, device=None, model: dict | None = None, linker_rank: int | None = 0) -> dict:
    """One-shot row cuts for `nparts` ranks: coarse K1 pre-pass (<= 2048 x 2048 samples, every rank runs the same
    deterministic pass and gets the same cuts) -> per-row cost estimate -> contiguous blocks of equal estimated cost.
    linker_rank: the rank that also gathers and chains the crossing records (rank 0 in ShardedBoundary); its block is
    cut lighter by the model's estimate of that serial stage, so that the other ranks do not wait for it in the next
    step's halo exchange (None: pure K1 balance).
    -> {"cuts", "profile", "balance_estimate", "model", "setup", "extra"}"""
    import time
    ys = np.asarray(ys, dtype=np.float64)
    m = COST_MODEL_B200 if model is None else model
    desc = (f"cost = it + {m['b_pixel']:g} + {m['d_late']:g}*[escaped, it>{m['knee']}] per pixel, constants fitted offline on a B200")
    if nparts == 1:
        return {"cuts": [0, int(ys.size)], "profile": np.ones(ys.size), "balance_estimate": 1.0, "model": desc,
                "setup": {"coarse_pass_ms": 0.0}, "extra": None}
    t0 = time.perf_counter()
    profile = coarse_row_profile(xs, ys, max_iter, model=m, device=device)
    extra = None
    if linker_rank is not None and m.get("link_seconds") and m.get("seconds_per_unit"):
        # the estimate is a constant measured on a 5.5e5-record boundary; small workloads have small boundaries, so it
        # is capped at 2 % of an even share (config 3 at N = 8: 1.3 ms of a 77 ms share is inside the cap)
        units = min(float(m["link_seconds"]) / float(m["seconds_per_unit"]), 0.02 * float(profile.sum()) / nparts)
        extra = np.zeros(nparts)
        extra[int(linker_rank)] = units
        desc += (f"; the linking rank {int(linker_rank)} is cut lighter by {1e3 * units * m['seconds_per_unit']:.2f} ms of K1")
    cuts = balanced_row_cuts(profile, nparts, extra)
    return {"cuts": cuts, "profile": profile, "balance_estimate": parallel_efficiency(profile, cuts, extra), "model": desc,
            "setup": {"coarse_pass_ms": 1e3 * (time.perf_counter() - t0), "coarse_samples": "<= 2048 x 2048",
                      "where": "device resident" if device is not None else "host-buffer API + numpy"},
            "extra": None if extra is None else extra.tolist()}


def refine_cuts(row_work, cuts, measured) -> list[int]:
    """One step of measured rebalancing: rescale the estimated work of every block so that its total equals the
    time (or any cost) MEASURED for that block, then cut again.  The iteration-count profile misses per-pixel
    overheads (exterior rows are cheap in iterations but not free), the measurement does not."""
    w = np.asarray(row_work, dtype=np.float64).ravel().copy()
    measured = np.asarray(measured, dtype=np.float64).ravel()
    if measured.size != len(cuts) - 1:
        raise ValueError("one measurement per block expected")
    for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        tot = w[a:b].sum()
        if tot > 0 and measured[k] > 0:
            w[a:b] *= measured[k] / tot
    return balanced_row_cuts(w, len(cuts) - 1)


# ---- collectives (torch.distributed; backend-agnostic) -------------------------------------
def exchange_first_rows(first_row, group=None):
    """All-gather every rank's first dwell row; returns the [world, nx] tensor.
    Rank r's K2 halo is row r+1 of the result (the last rank has none)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world, first_row.numel()), dtype=first_row.dtype, device=first_row.device)
    dist.all_gather_into_tensor(out, first_row.contiguous().view(1, -1), group=group)
    return out


def gather_records(records, n_records: int, dst: int = 0, group=None):
    """Gather the per-rank record arrays to `dst`, concatenated in rank order (= global raster order for
    contiguous row blocks), WITHOUT leaving the device: `records` is a [cap, 8] int64 tensor (on the GPU for
    NCCL, on the CPU for gloo) of which the first n_records rows count.  The counts are all-gathered, then every
    rank sends exactly its rows and `dst` receives them at their offsets of one tensor (point-to-point over
    NVLink; no padding, no host bounce).  Returns (tensor [total, 8], counts) on dst, (None, counts) elsewhere."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = torch.zeros(world, dtype=torch.int64, device=records.device)
    mine = torch.tensor([int(n_records)], dtype=torch.int64, device=records.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_h = [int(c) for c in counts.cpu().tolist()]
    if rank != dst:
        if n_records:
            dist.send(records[:n_records].contiguous(), dst=dst, group=group)
        return None, counts_h
    offs = np.concatenate([[0], np.cumsum(counts_h)]).astype(np.int64)
    out = torch.empty((int(offs[-1]), 8), dtype=torch.int64, device=records.device)
    if n_records:
        out[int(offs[rank]):int(offs[rank + 1])] = records[:n_records]
    for r in range(world):
        if r != dst and counts_h[r]:
            dist.recv(out[int(offs[r]):int(offs[r + 1])], src=r, group=group)
    return out, counts_h


# ---- Lucas-Loci cloud (config 5): polynomials are independent -> contiguous slices, one all-reduce ----
def item_slices(n_items: int, nparts: int) -> list[int]:
    """nparts+1 cut positions of [0, n_items) into contiguous, nearly equal slices (K3 is embarrassingly
    parallel over polynomials; random degrees balance themselves)."""
    if nparts < 1:
        raise ValueError("nparts must be >= 1")
    return [(n_items * k) // nparts for k in range(nparts + 1)]


def allreduce_field_sums(sums, n_points_local: int, group=None):
    """Sum the per-cell raw log sums (a torch tensor, on the GPU for NCCL) and the cloud sizes over the
    ranks: the one data-path collective of the cloud stage (K4a with the points sharded).
    Returns (sums, n_points_total); `sums` is reduced in place."""
    import torch
    import torch.distributed as dist
    cnt = torch.tensor([int(n_points_local)], dtype=torch.int64, device=sums.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return sums, int(cnt.item())


# ---- potential field of a row-sharded grid: all-gather of the row blocks, halo rows for the stencils ----
def allgather_rows(block, cuts, group=None):
    """All-gather the row blocks of a row-sharded field ([rows_r, nx] on rank r, rows cut at `cuts`) into the
    full [ny, nx] field on every rank (NCCL all-gather of equal-size padded blocks).  This is how the final
    potential field reaches every GPU for K4 / K4a."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return block
    rows = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
    nx = block.shape[1]
    pad = max(rows)
    mine = torch.zeros((pad, nx), dtype=block.dtype, device=block.device)
    mine[: block.shape[0]] = block
    gathered = torch.empty((world, pad, nx), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(gathered, mine.view(1, pad, nx), group=group)
    return torch.cat([gathered[r, : rows[r]] for r in range(world)], dim=0)


def exchange_halo_rows(block, periodic: bool, group=None):
    """Halo rows of a row-sharded field for the 5-point stencils: returns (row_above, row_below) for this rank,
    taken from the neighbours' last / first rows (one all-gather of 2 rows per rank).  periodic=True wraps
    first <-> last rank (np.roll semantics of laplacian, Laplacian_C-M.py:49-59); otherwise the outer ranks get
    None (interior stencil with copied border, variograms_construct_mandelbrot.py:169-173)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nx = block.shape[1]
    edges = torch.stack([block[0], block[-1]]).contiguous()               # [2, nx]: my first and last row
    allv = torch.empty((world, 2, nx), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(allv, edges.view(1, 2, nx), group=group)
    above = allv[(rank - 1) % world, 1] if (periodic or rank > 0) else None
    below = allv[(rank + 1) % world, 0] if (periodic or rank < world - 1) else None
    return above, below


# ---- the boundary stage of the script on N GPUs (one process per GPU, torch.distributed initialised) ----
class ShardedBoundary:
    """compute_grid + extract_contour (mandelbrot_boundary_sample.py:66-67) with the rows sharded over the ranks.

    Every rank computes K1 on its row block through the host-buffer shard call (the block comes back to a pinned
    host array AND stays in HBM with a halo slot), the first dwell row of every block is all-gathered over NCCL
    straight from / into those blocks, K2 classifies each block on the device, the records (device tensors) are
    sent to rank 0's GPU over NCCL and chained there by the device linker; only the finished polylines go to the
    host.  Cuts come from a coarse K1 pre-pass and a per-device cost model; `refine()` re-cuts from measured
    block times.  All device work is enqueued on torch's CURRENT stream (the collectives order against it).
    Works with world size 1 too (then it is just the fused single-GPU call)."""

    def __init__(self, xs, ys, max_iter: int, level: float, device=None, with_potential: bool = False, cuts=None,
                 profile=None):
        import torch
        import torch.distributed as dist
        from . import _shim
        self.xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
        self.ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
        self.max_iter, self.level, self.with_potential = int(max_iter), float(level), bool(with_potential)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        if cuts is None:
            plan = plan_row_cuts(self.xs, self.ys, self.max_iter, self.world, device=self.device)
            self.profile, cuts = plan["profile"], plan["cuts"]
        else:
            self.profile = np.ones(self.ys.size) if profile is None else np.asarray(profile, dtype=np.float64)
        self._set_cuts(cuts)
        self._shim = _shim
        self.last_work_units = 0

    def _set_cuts(self, cuts):
        from . import _shim
        import torch
        self.cuts = [int(c) for c in cuts]
        self.r0, self.r1 = self.cuts[self.rank], self.cuts[self.rank + 1]
        self.rows = self.r1 - self.r0
        self.has_halo = self.rank < self.world - 1
        nx = self.xs.size
        self.dwell = _shim.pinned_empty((self.rows, nx), np.int32)             # this rank's rows, page-locked
        self.potential = _shim.pinned_empty((self.rows, nx), np.float64) if self.with_potential else None
        self.ys_rows = np.ascontiguousarray(self.ys[self.r0:self.r1])
        self.ys_block = np.ascontiguousarray(self.ys[self.r0:self.r1 + (1 if self.has_halo else 0)])
        self._edge = torch.empty(nx, dtype=torch.int32, device=self.device)
        self._field = (torch.empty((self.rows, nx), dtype=torch.float64, device=self.device)
                       if (self.with_potential and self.world > 1) else None)
        self._records = torch.empty((max(int(0.002 * self.rows * nx) + 4096, 1 << 16), 8), dtype=torch.int64, device=self.device)
        self.full_potential = None
        # the row-cost profile of this rank's rows: lm_shard_escape computes its chunks cheapest first
        self._row_cost = (np.ascontiguousarray(self.profile[self.r0:self.r1], dtype=np.float64)
                          if np.asarray(self.profile).size == self.ys.size and np.ptp(self.profile) > 0 else None)
        self._lines_buf = None           # rank 0: page-locked (verts, offsets) the linked lines are copied into
        self.last_phases_ms = {}

    def refine(self, rounds: int = 2):
        """Measured rebalancing: time K1 on the current blocks on every rank, rescale the profile, cut again."""
        import time
        import ctypes as C
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.cuts
        for _ in range(rounds):
            st = self._shim.Stats()
            blk = C.c_void_p(); pblk = C.c_void_p()
            dist.barrier()
            self._shim.call("lm_shard_escape", self._shim.ptr(self.xs), self.xs.size, self._shim.ptr(self.ys_rows), self.rows,
                            self.max_iter, None, None, 1, None, C.byref(blk), C.byref(pblk), C.byref(st))
            t_all = torch.zeros(self.world, dtype=torch.float64, device=self.device)
            t_all[self.rank] = float(st.kernel_ms)
            dist.all_reduce(t_all)
            self._set_cuts(refine_cuts(self.profile, self.cuts, t_all.cpu().numpy()))
        return self.cuts

    def run(self):
        """One pass.  Returns the polylines on rank 0 (None elsewhere); self.dwell / self.potential hold this rank's
        rows, self.full_potential the all-gathered field (device tensor) when asked for."""
        import ctypes as C
        import torch
        from . import contour
        shim = self._shim
        nx = self.xs.size
        if self.world == 1:
            lines, st = contour.boundary_sample(self.xs, self.ys, self.max_iter, self.level, dwell_out=self.dwell,
                                                potential_out=self.potential)
            self.last_work_units = int(st["work_units"])
            return lines
        import time
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        st = shim.Stats()
        blk = C.c_void_p(); pblk = C.c_void_p()
        t0 = time.perf_counter()
        shim.call("lm_shard_escape", shim.ptr(self.xs), nx, shim.ptr(self.ys_rows), self.rows, self.max_iter,
                  shim.ptr(self.dwell), shim.ptr(self.potential), 1, shim.ptr(self._row_cost), C.byref(blk), C.byref(pblk), C.byref(st))
        t1 = time.perf_counter()
        self.last_work_units = int(st.work_units)
        if self.with_potential:     # final potential field on every GPU: all-gather straight from the resident block
            shim.call("lm_memcpy_d2d", C.c_void_p(self._field.data_ptr()), pblk, self.rows * nx * 8, stream)
            self.full_potential = allgather_rows(self._field, self.cuts)
        shim.call("lm_memcpy_d2d", C.c_void_p(self._edge.data_ptr()), blk, nx * 4, stream)
        firsts = exchange_first_rows(self._edge)
        if self.has_halo:
            shim.call("lm_memcpy_d2d", C.c_void_p(blk.value + self.rows * nx * 4), C.c_void_p(firsts[self.rank + 1].data_ptr()),
                      nx * 4, stream)
        n_rec = C.c_int64(0)
        lib = shim.load()
        while True:
            rc = lib.lm_contour_records_dev(blk, shim.ptr(self.xs), nx, shim.ptr(self.ys_block), self.ys_block.size, self.r0,
                                            self.level, C.c_void_p(self._records.data_ptr()), self._records.shape[0],
                                            C.byref(n_rec), stream)
            if rc == shim.LM_E_CAP:
                self._records = torch.empty((n_rec.value + 1024, 8), dtype=torch.int64, device=self.device)
                continue
            shim.check(rc)
            break
        t2 = time.perf_counter()                 # includes waiting for the slowest rank's K1 in the all-gather
        self.n_records = int(n_rec.value)
        allrec, counts = gather_records(self._records, self.n_records, 0)
        self.n_records_total = int(sum(counts))
        t3 = time.perf_counter()
        lines = None
        if self.rank == 0:
            lines = self._link(allrec, stream)
        t4 = time.perf_counter()
        self.last_phases_ms = {"k1_host_call": 1e3 * (t1 - t0), "k1_kernels": float(st.kernel_ms),
                               "halo_exchange_and_records": 1e3 * (t2 - t1), "record_gather": 1e3 * (t3 - t2),
                               "link_and_fetch": 1e3 * (t4 - t3)}
        return lines

    def _link(self, allrec, stream):
        """rank 0: chain all records on the device, copy the lines into a cached page-locked buffer (grown when
        too small) and hand out exact-size copies."""
        import ctypes as C
        from . import contour
        shim = self._shim
        lib = shim.load()
        n = int(allrec.shape[0])
        if self._lines_buf is None or self._lines_buf[0].shape[0] < 2 * n + 16:
            self._lines_buf = (shim.pinned_empty((4 * n + 64, 2), np.float64), shim.pinned_empty(2 * n + 65, np.int64))
        verts, offs = self._lines_buf
        nv = C.c_int64(0); nl = C.c_int64(0)
        rc = lib.lm_contour_link_dev(C.c_void_p(allrec.data_ptr()), n, shim.ptr(self.xs), self.xs.size, shim.ptr(self.ys), self.ys.size,
                                     float(self.level), shim.ptr(verts), verts.shape[0], C.byref(nv), shim.ptr(offs), offs.size - 1,
                                     C.byref(nl), stream)
        shim.check(rc)
        return contour.Polylines(np.array(verts[: nv.value]), np.array(offs[: nl.value + 1]))


def sharded_cloud_fields(toprows_local, deg_local, grid_x, grid_y, tol: float = 1e-12, eps: float = 1e-12, variant: int = 0,
                         h: float | None = None, potential: tuple | None = None, device=None) -> dict:
    """lucas.cloud_fields with the polynomials sharded over the ranks (one process per GPU): every rank passes ITS
    slice of the batch (see item_slices), runs K3 -> cloud -> K1d -> K4a partial sums on it, the per-cell sums and
    the cloud sizes are all-reduced (NCCL), and every rank ends with the same U and Laplacian.
    Returns {"cloud" (this rank's points, complex), "g", "it" (this rank's), "U", "lapU", "n_points_total"}."""
    import ctypes as C
    import torch
    from . import _shim
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    # every kernel goes to torch's CURRENT stream: the NCCL collectives below order against that stream only
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())
    top_h = np.ascontiguousarray(toprows_local, dtype=np.float64)
    deg_h = np.ascontiguousarray(deg_local, dtype=np.int32).reshape(-1)
    npoly, maxdeg = top_h.shape
    nroots = int(np.clip(deg_h, 0, None).sum())
    top = torch.from_numpy(top_h).to(dev); deg = torch.from_numpy(deg_h).to(dev)
    re = torch.empty((max(npoly, 1), maxdeg), dtype=torch.float64, device=dev); im = torch.empty_like(re)
    kept = torch.empty(max(npoly, 1), dtype=torch.int32, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    px = torch.empty(max(nroots, 1), dtype=torch.float64, device=dev); py = torch.empty_like(px)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    _shim.call("lm_roots_batched_dev", P(top), P(deg), npoly, maxdeg, 1, float(tol), P(re), P(im), P(kept), None, P(status), stream)
    _shim.call("lm_cloud_compact_dev", P(re), P(im), P(kept), npoly, maxdeg, P(px), P(py), nroots, P(cnt), stream)
    n_local = int(cnt.item())
    # the status flags are all-reduced (MAX) BEFORE anybody raises: a rank-local exception ahead of the collectives
    # below would leave the other ranks waiting in them forever, and every rank must report the same outcome
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(status, op=dist.ReduceOp.MAX)
    flags = status.cpu().numpy()
    if flags[1]:
        raise ValueError("a degree outside [1, maxdeg] (on some rank)")
    out = {"cloud": px[:n_local].cpu().numpy() + 1j * py[:n_local].cpu().numpy(), "g": None, "it": None}
    if potential is not None:
        g = torch.empty(max(n_local, 1), dtype=torch.float64, device=dev)
        it = torch.empty(max(n_local, 1), dtype=torch.int64, device=dev)
        _shim.call("lm_escape_points_f64_dev", P(px), P(py), n_local, int(potential[0]), float(potential[1]), P(g), P(it),
                   None, None, None, stream)
        out["g"], out["it"] = g[:n_local].cpu().numpy(), it[:n_local].cpu().numpy()
    gx = torch.from_numpy(np.ascontiguousarray(grid_x, dtype=np.float64).ravel()).to(dev)
    gy = torch.from_numpy(np.ascontiguousarray(grid_y, dtype=np.float64).ravel()).to(dev)
    nx, ny = gx.numel(), gy.numel()
    sums = torch.empty(nx * ny, dtype=torch.float64, device=dev)
    _shim.call("lm_log_potential_sums_dev", P(px), P(py), n_local, P(gx), nx, P(gy), ny, float(eps), int(variant), P(sums), stream)
    _, n_total = allreduce_field_sums(sums, n_local)
    U = torch.empty_like(sums); lap = torch.empty_like(sums)
    _shim.call("lm_log_potential_finish_dev", P(sums), nx * ny, n_total, int(variant), P(U), stream)
    if h is None:
        h = float(gx[1] - gx[0]) if nx > 1 else 1.0
    _shim.call("lm_laplacian5_periodic_dev", P(U), ny, nx, float(h), P(lap), stream)
    out["U"] = U.cpu().numpy().reshape(ny, nx); out["lapU"] = lap.cpu().numpy().reshape(ny, nx)
    out["n_points_total"] = n_total
    if flags[0]:
        raise RuntimeError("the Aberth iteration did not converge for some polynomial (on some rank)")
    return out

#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 hot path (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1|cfg4|cfg5]

Metric (BASELINE.json): G pixel-iterations/s in fp64, work unit = sum over pixels of
min(dwell+1, max_iter) counted exactly by the kernel.  A step is one pass of the boundary stage
over the whole window: K1 (escape-time dwell grid, fp64, bit-exact) followed by K2 (level set:
crossing records -> ORDERED boundary polylines, all on the device); with N > 1 the rows are
sharded over the ranks at equal estimated cost, the shard-edge dwell rows are all-gathered over
NCCL for K2's halo and the shards' records are sent GPU-to-GPU to rank 0, which links them.

  value     whole-job throughput with inputs/outputs resident in HBM (device-timed, max over ranks)
  e2e       the same through the host-buffer C ABI (pinned numpy buffers in, dwell grid and ordered
            polylines out): H2D and D2H inside the region
  roofline  K1 against the FP64 peak measured live by lm_probe_fp64_peak (MEASURED_PEAKS.json has
            no FP64 entry): achieved = pixel_iters x 8 flops / K1 device time
  sub_rooflines   K2 (records) and K4 (5-point stencil) against the measured HBM copy bandwidth,
            K3 against the FP64 probe
  cpu_baseline / --impl reference
            the CPU port of the reference (oracle/lm_oracle.c, pthreads over all host cores) on a
            bounded row sample of the same workload, and beside it the reference's own pure-Python
            loop (oracle/pyref.py; the reference's `def` itself where /root/reference exists)
  legs      one-step runs of the other grid configs of BASELINE.json (cfg1, cfg2, cfg4)
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # BASELINE.json configs[2] (the grid the target is quoted on; fits one GPU), [1], [0] and [3]
    "cfg3": dict(xlim=(-2.1, 0.9), ylim=(-1.5, 1.5), res=32768, max_iter=10000, level=0.96),
    "cfg2": dict(xlim=(-2.1, 0.9), ylim=(-1.5, 1.5), res=8192, max_iter=2000, level=0.96),
    "cfg1": dict(xlim=(-2.1, 0.9), ylim=(-1.5, 1.5), res=2000, max_iter=500, level=0.96),
    "cfg4": dict(xlim=(-0.755, -0.735), ylim=(0.10, 0.12), res=16384, max_iter=100000, level=0.96),
}
FLOPS_PER_PIXEL_ITER = 8          # 3 mul + 5 add under the reference's unfused semantics

# BASELINE.json configs[4]: 10^7 generalized-Lucas polynomials of degree <= 25, generated in
# 16 independently seeded chunks so that every rank can build just its slice
CFG5 = dict(npoly=10_000_000, maxdeg=25, chunks=16, grid=400, pot_max_iter=20000, pot_radius=2.0, eps=1e-12)


def cfg5_chunk(k: int, npoly_total: int = CFG5["npoly"], chunks: int = CFG5["chunks"], maxdeg: int = CFG5["maxdeg"]):
    """Chunk k of the config-5 batch: degrees uniform in [2, maxdeg], first-row entries in {0,1,2}, a_d >= 1."""
    lo, hi = (npoly_total * k) // chunks, (npoly_total * (k + 1)) // chunks
    n = hi - lo
    rng = np.random.default_rng([0, k])
    deg = rng.integers(2, maxdeg + 1, size=n).astype(np.int32)
    top = rng.integers(0, 3, size=(n, maxdeg)).astype(np.float64)
    top[np.arange(maxdeg)[None, :] >= deg[:, None]] = 0.0
    last = top[np.arange(n), deg - 1]
    top[np.arange(n), deg - 1] = np.where(last == 0, 1.0, last)
    return top, deg


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["cfg5"], default="cfg3")
    ap.add_argument("--no-roots", action="store_true", help="skip the Lucas-roots leg of the default run")
    ap.add_argument("--no-legs", action="store_true", help="skip the one-step cfg1 / cfg2 / cfg4 legs of the default run")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--refine", type=int, default=0,
                    help="N > 1: rounds of measured re-cutting (untimed full K1 passes; charged to setup_ms). Default 0: one-shot cuts")
    ap.add_argument("--roots-f64-rows", action="store_true",
                    help="hand the Lucas first rows to the e2e call as float64 (8 B per coefficient) instead of int8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(w: dict) -> str:
    return (f"mandelbrot_boundary_sample --xlim {w['xlim'][0]} {w['xlim'][1]} --ylim {w['ylim'][0]} {w['ylim'][1]} "
            f"--res {w['res']} --max_iter {w['max_iter']} --level {w['level']}")


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of rows; the reference-speed Python loop
# ---------------------------------------------------------------------------------------------
def cpu_sample_rows(ny: int, every: int) -> np.ndarray:
    return np.arange(every // 2, ny, every, dtype=np.int64)


def cpu_pass(w: dict, every: int):
    """One pass of the CPU port over every `every`-th row of the workload. -> (pixel_iters, seconds, threads)"""
    from oracle import oracle
    xs = np.linspace(w["xlim"][0], w["xlim"][1], w["res"])
    ys = np.linspace(w["ylim"][0], w["ylim"][1], w["res"])
    rows = cpu_sample_rows(ys.size, every)
    t0 = time.perf_counter()
    _, work = oracle.dwell_grid(xs, ys[rows], w["max_iter"])
    dt = time.perf_counter() - t0
    return work, dt, oracle.num_threads(), rows.size


def pick_cpu_stride(w: dict, target_s: float) -> int:
    """Row stride so that one CPU pass takes roughly target_s (calibrated with a tiny pass)."""
    probe_every = max(w["res"] // 16, 1)
    work, dt, _, _ = cpu_pass(w, probe_every)
    rate = work / max(dt, 1e-6)
    total_est = work * probe_every
    every = int(max(1, round(total_est / (rate * target_s))))
    return min(every, probe_every)


def python_loop_baseline(w: dict, budget_s: float) -> dict:
    """The reference's own double loop over CPython complex numbers (mandelbrot_boundary_sample.py:22-39) on random
    rows of the workload, every 64th column: kind "reference" where the reference checkout exists (its `def` run as
    is), else the line-for-line restatement oracle/pyref.py (kind "port")."""
    from oracle import pyref
    xs = np.linspace(w["xlim"][0], w["xlim"][1], w["res"])
    ys = np.linspace(w["ylim"][0], w["ylim"][1], w["res"])
    stride = max(w["res"] // 512, 1)
    r = pyref.timed_sample(xs, ys, w["max_iter"], budget_s=budget_s, cols_stride=stride)
    return {"value": r["value"], "unit": "Gpixel-iter/s", "cores": 1, "kind": r["kind"], "language": "python",
            "sample": f"{r['rows']} random rows x every {stride}th column ({r['cols']} pixels per row) of the workload, "
                      f"{r['pixel_iters']} pixel-iterations in {r['seconds']:.1f} s (mandelbrot_dwell in CPython, one core)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    every = pick_cpu_stride(w, target_s=8.0)
    for _ in range(args.warmup):
        cpu_pass(w, every)
    t_total, work_total, threads, nrows = 0.0, 0, 1, 0
    for _ in range(args.steps):
        work, dt, threads, nrows = cpu_pass(w, every)
        t_total += dt; work_total += work
    value = work_total / t_total / 1e9
    sample = f"every {every}th row ({nrows} of {w['res']} rows, full width) of the workload per step"
    line = {
        "impl": "reference", "metric": "gpixel_iter_per_s_fp64", "value": value, "unit": "Gpixel-iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gpixel-iter/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gpixel-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "python_loop": python_loop_baseline(w, 10.0),
        "note": "CPU port of the reference's scalar loop (oracle/lm_oracle.c), all host threads; python_loop is the "
                "reference's own CPython loop on one core, timed beside it",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        return {}


def profiled_traffic(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the K1 launch from the committed ncu capture of THIS round
    (profiles/r02_k1_traffic.json, written by scripts/ncu_traffic.py from the --set full CSV); None when there is none."""
    try:
        t = json.loads((ROOT / "profiles" / "r02_k1_traffic.json").read_text())
        e = t.get(workload)
        return (float(e["dram_bytes"]), e["source"]) if e else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


# ---------------------------------------------------------------------------------------------
# GPU arm: one grid workload (the headline, or a one-step leg)
# ---------------------------------------------------------------------------------------------
class Ctx:
    pass


def dwell_checksum(dwell_rows, r0: int, nx: int) -> int:
    """Position-weighted sum of the dwell block mod 2^64 (int64 wrap-around): additive over row blocks, so the
    all-reduced value is the same for every sharding of the same grid."""
    import torch
    acc = torch.zeros((), dtype=torch.int64, device=dwell_rows.device)
    cols = torch.arange(nx, dtype=torch.int64, device=dwell_rows.device)[None, :]
    for a in range(0, dwell_rows.shape[0], 1024):
        blk = dwell_rows[a:a + 1024].to(torch.int64)
        rows = torch.arange(r0 + a, r0 + a + blk.shape[0], dtype=torch.int64, device=dwell_rows.device)[:, None]
        wgt = (rows * nx + cols) * 2654435761 + 40503
        acc += (blk * wgt).sum()
    return int(acc.item())


def lines_digest(verts: np.ndarray, offsets: np.ndarray) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(offsets, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(verts, dtype=np.float64).tobytes())
    return h.hexdigest()[:16]


def grid_leg(cx: Ctx, name: str, steps: int, warmup: int, want_e2e: bool, refine: int = 0) -> dict:
    """K1 + K2 on workload `name` over the ranks of cx.  -> fields of the JSON line (rank 0 builds the line)."""
    import torch
    import torch.distributed as dist
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, contour, sharding
    rank, world, dev, lib, stream = cx.rank, cx.world, cx.dev, cx.lib, cx.stream
    w = WORKLOADS[name]
    res, max_iter = w["res"], w["max_iter"]
    level = w["level"] * max_iter
    xs = np.linspace(w["xlim"][0], w["xlim"][1], res)
    ys = np.linspace(w["ylim"][0], w["ylim"][1], res)
    nx, ny = xs.size, ys.size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- row shard of this rank: contiguous blocks cut ONCE at equal estimated cost (coarse K1 pre-pass weighted by
    # the per-device cost model of sharding.plan_row_cuts); nothing is calibrated on the workload itself unless --refine
    barrier()
    t_setup = time.perf_counter()
    plan = sharding.plan_row_cuts(xs, ys, max_iter, world, device=dev)
    cuts = plan["cuts"]
    balance_estimate_only = plan["balance_estimate"]
    refined = 0
    for _ in range(refine if world > 1 else 0):
        ra, rb = cuts[rank], cuts[rank + 1]
        cal_ys = torch.from_numpy(np.ascontiguousarray(ys[ra:rb])).to(dev)
        cal_xs = torch.from_numpy(xs).to(dev)
        cal_d = torch.empty((rb - ra, nx), dtype=torch.int32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        _shim.call("lm_escape_grid_f64_dev", C.c_void_p(cal_xs.data_ptr()), nx, C.c_void_p(cal_ys.data_ptr()), rb - ra,
                   max_iter, 2.0, 0, C.c_void_p(cal_d.data_ptr()), None, None, None, stream)
        e1.record(); torch.cuda.synchronize()
        t_all = torch.zeros(world, dtype=torch.float64, device=dev)
        t_all[rank] = e0.elapsed_time(e1)
        dist.all_reduce(t_all)
        cuts = sharding.refine_cuts(plan["profile"], cuts, t_all.cpu().numpy())
        refined += 1
        del cal_d, cal_ys, cal_xs
    barrier()
    setup_ms = 1e3 * (time.perf_counter() - t_setup)
    r0, r1 = cuts[rank], cuts[rank + 1]
    rows = r1 - r0
    has_halo = rank < world - 1
    ys_block = np.ascontiguousarray(ys[r0:r1 + (1 if has_halo else 0)])

    xs_d = torch.from_numpy(xs).to(dev)
    ys_d = torch.from_numpy(ys_block).to(dev)
    dwell_d = torch.empty((rows + 1, nx), dtype=torch.int32, device=dev)      # +1: halo row slot
    # BASELINE.json configs[1] asks for the smooth potential with the dwell grid; at N > 1 the final field is
    # all-gathered so that every GPU holds it (north_star)
    with_pot = name == "cfg2"
    field_d = torch.empty((rows, nx), dtype=torch.float64, device=dev) if with_pot else None
    work_d = torch.zeros(1, dtype=torch.int64, device=dev)
    rec_d = torch.empty((max(int(0.002 * rows * nx) + 4096, 1 << 16), 8), dtype=torch.int64, device=dev)
    n_rec = C.c_int64(0)
    nv = C.c_int64(0); nl = C.c_int64(0)
    first = np.zeros(1, dtype=np.int64)
    launches = {"n": 0}
    state = {"rec_d": rec_d, "n_rec_total": 0}

    def k1():
        _shim.call("lm_escape_grid_f64_dev", C.c_void_p(xs_d.data_ptr()), nx, C.c_void_p(ys_d.data_ptr()), rows,
                   max_iter, 2.0, 1 if with_pot else 0, C.c_void_p(dwell_d.data_ptr()), None,
                   C.c_void_p(field_d.data_ptr()) if with_pot else None, C.c_void_p(work_d.data_ptr()), stream)
        launches["n"] += 1

    def k2_records():
        """halo exchange + mark / scan / emit: this rank's crossing records, resident"""
        if world > 1:
            firsts = sharding.exchange_first_rows(dwell_d[0])
            if has_halo:
                dwell_d[rows].copy_(firsts[rank + 1])
            if with_pot:
                state["full_field"] = sharding.allgather_rows(field_d, cuts)
        while True:
            rc = lib.lm_contour_records_dev(C.c_void_p(dwell_d.data_ptr()), _shim.ptr(xs), nx, _shim.ptr(ys_block),
                                            rows + (1 if has_halo else 0), r0, float(level),
                                            C.c_void_p(state["rec_d"].data_ptr()), state["rec_d"].shape[0], C.byref(n_rec), stream)
            if rc == _shim.LM_E_CAP:
                state["rec_d"] = torch.empty((n_rec.value + 1024, 8), dtype=torch.int64, device=dev)
                continue
            _shim.check(rc)
            break
        launches["n"] += 3 if n_rec.value else 2

    def k2_link():
        """records of all shards -> rank 0 (GPU to GPU) -> ordered polylines, left on the device"""
        if world > 1:
            allrec, counts = sharding.gather_records(state["rec_d"], n_rec.value, 0)
            state["n_rec_total"] = int(sum(counts))
        else:
            allrec = state["rec_d"]
            state["n_rec_total"] = n_rec.value
        if rank == 0:
            rc = lib.lm_contour_link_dev(C.c_void_p(allrec.data_ptr()), state["n_rec_total"], _shim.ptr(xs), nx, _shim.ptr(ys), ny,
                                         float(level), None, 0, C.byref(nv), _shim.ptr(first), 0, C.byref(nl), stream)
            if rc != _shim.LM_E_CAP:          # LM_E_CAP with zero capacity = "linked, kept on the device"
                _shim.check(rc)
            launches["n"] += 2 if state["n_rec_total"] else 0

    for _ in range(warmup):
        k1(); k2_records(); k2_link()
    barrier()
    launches["n"] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 + 4 * steps)]
    work_steps_d = torch.zeros(steps, dtype=torch.int64, device=dev)     # per-step work counters, read after the timed region
    _ = int(work_d.item())
    barrier()
    ev[0].record()
    for k in range(steps):
        ev[2 + 4 * k].record()
        k1()
        ev[3 + 4 * k].record()
        k2_records()
        ev[4 + 4 * k].record()
        k2_link()
        ev[5 + 4 * k].record()
        work_steps_d[k:k + 1].copy_(work_d)
    ev[1].record()
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    k1_ms = sum(ev[2 + 4 * k].elapsed_time(ev[3 + 4 * k]) for k in range(steps))
    k2a_ms = sum(ev[3 + 4 * k].elapsed_time(ev[4 + 4 * k]) for k in range(steps))
    k2b_ms = sum(ev[4 + 4 * k].elapsed_time(ev[5 + 4 * k]) for k in range(steps))
    step_ms = [ev[2 + 4 * k].elapsed_time(ev[5 + 4 * k]) for k in range(steps)]
    gap_ms = [ev[0].elapsed_time(ev[2])] + [ev[5 + 4 * k].elapsed_time(ev[6 + 4 * k]) for k in range(steps - 1)] + [ev[1 + 4 * steps].elapsed_time(ev[1])]
    my_work = int(work_steps_d.sum().item())
    t = torch.tensor([elapsed_ms, k1_ms, k2a_ms, k2b_ms], dtype=torch.float64, device=dev)
    tmin = torch.tensor([k1_ms], dtype=torch.float64, device=dev)
    wk = torch.tensor([my_work], dtype=torch.int64, device=dev)
    crc = torch.tensor([dwell_checksum(dwell_d[:rows], r0, nx)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.SUM)
        dist.all_reduce(wk, op=dist.ReduceOp.SUM)
        dist.all_reduce(crc, op=dist.ReduceOp.SUM)
    elapsed_ms, k1_ms_max, k2a_ms_max, k2b_ms_max = (float(v) for v in t)
    total_work = int(wk[0])
    value = total_work / (elapsed_ms * 1e-3) / 1e9
    out = {"name": name, "w": w, "value": value, "elapsed_ms": elapsed_ms, "ms_per_step": elapsed_ms / steps,
           "total_work": total_work, "cuts": cuts, "plan": plan, "setup_ms": setup_ms, "refined": refined,
           "balance_estimate_only": balance_estimate_only,
           "balance_measured": float(tmin[0]) / world / max(k1_ms_max, 1e-9),      # mean / max of the ranks' K1 time
           "k1_ms_per_step_max_rank": k1_ms_max / steps, "gpu_launches": launches["n"], "with_pot": with_pot,
           "dwell_crc": f"{int(crc[0]) & 0xFFFFFFFFFFFFFFFF:016x}", "n_records": state["n_rec_total"],
           "step_ms_rank0": step_ms, "gap_ms_rank0": gap_ms}

    # the device-resident lines of the last step (rank 0) -> digest
    if rank == 0:
        verts = np.empty((nv.value, 2)); offs = np.empty(nl.value + 1, dtype=np.int64)
        _shim.check(lib.lm_contour_fetch_last(_shim.ptr(verts), nv.value, C.byref(nv), _shim.ptr(offs), nl.value, C.byref(nl)))
        out["boundary_sha256"] = lines_digest(verts, offs)
        out["boundary_lines"] = int(nl.value)
        out["boundary_vertices_total"] = int(nv.value)
        out["boundary_vertices"] = int(np.diff(offs).max()) if nl.value else 0

    # ---- roofline of the dominant kernel (K1) on this rank, K2 against HBM
    k1_gpi = my_work / (k1_ms * 1e-3) / 1e9
    achieved_tflops = k1_gpi * 1e9 * FLOPS_PER_PIXEL_ITER / 1e12
    peak = cx.peak_tflops
    traffic, traffic_src = profiled_traffic(name) if world == 1 else (None, None)
    out["roofline"] = {
        "bound": "fp64", "kernel": "lm_escape_kernel<grid, dwell + potential>" if with_pot else "lm_escape_kernel<grid, dwell>",
        "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tflops / peak,
        "peak_source": "lm_probe_fp64_peak (dependent-free DFMA loop, 2 flops/DFMA), measured live; "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "algorithmic_flops_per_unit": FLOPS_PER_PIXEL_ITER,
        "fp64_pipe_instr_util": k1_gpi * 1e9 * 6 / (peak * 1e12 / 2),
        "k1_gpixel_iter_per_s": k1_gpi, "traffic": traffic, "traffic_source": traffic_src,
    }
    hbm = cx.hbm_gbs
    k2_bytes = 4.0 * (rows + (1 if has_halo else 0)) * nx * steps
    out["k2_roofline"] = {
        "bound": "hbm", "kernel": "contour_mark_bulk_kernel + contour_scan_kernel + contour_emit_kernel (+ NCCL halo row at N > 1)",
        "achieved": k2_bytes / (k2a_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": k2_bytes / (k2a_ms * 1e-3) / 1e9 / hbm,
        "algorithmic_bytes_per_unit": 4, "ms_per_step": k2a_ms / steps,
        "link_ms_per_step_max_rank": k2b_ms_max / steps, "records_ms_per_step_max_rank": k2a_ms_max / steps,
        "note": "timed with CUDA events around lm_contour_records_dev on this rank; it includes the host sync that reads the "
                "record count; link = record gather to rank 0 + link_build_kernel + link_rank_kernel"}

    # ---- e2e through the host-buffer C ABI
    if want_e2e:
        out_h = _shim.pinned_empty((rows, nx), np.int32)
        xs_p = _shim.pinned_empty(nx, np.float64); xs_p[:] = xs
        ys_p = _shim.pinned_empty(rows, np.float64); ys_p[:] = ys[r0:r1]
        pot_p = _shim.pinned_empty((rows, nx), np.float64) if with_pot else None
        job = (sharding.ShardedBoundary(xs, ys, max_iter, level, device=dev, with_potential=with_pot, cuts=cuts, profile=plan["profile"])
               if world > 1 else None)

        def e2e_step():
            if world == 1:
                # the fused host-buffer call (compute_grid + extract_contour of the script's main()):
                # H2D of xs/ys, chunked K1, dwell grid copied back to the pinned host buffer while
                # K1/K2 still run, K2 records -> ordered polylines on the device -> host
                lines, stx = contour.boundary_sample(xs_p, ys_p, max_iter, level, dwell_out=out_h, potential_out=pot_p)
                return stx["work_units"], lines
            # N > 1: sharding.ShardedBoundary.run: K1 on this rank's rows through the host-buffer shard call (block
            # returned to pinned host memory AND kept in HBM), shard-edge rows all-gathered over NCCL straight from /
            # into those blocks, K2 records on them, records sent GPU-to-GPU to rank 0 and linked there on the device
            lines = job.run()
            return job.last_work_units, lines

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_work = 0
        lines = None
        for _ in range(steps):
            wu, lines = e2e_step()
            e2e_work += int(wu)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        ww = torch.tensor([e2e_work], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        host_dwell = job.dwell if job else out_h
        host_crc = torch.tensor([dwell_checksum(torch.from_numpy(host_dwell[: min(rows, 4096)]).to(dev), r0, nx)], dtype=torch.int64, device=dev)
        dev_crc = torch.tensor([dwell_checksum(dwell_d[: min(rows, 4096)], r0, nx)], dtype=torch.int64, device=dev)
        lines_bytes = int(lines.verts.nbytes + lines.offsets.nbytes) if lines is not None else 0
        out["e2e"] = {"value": int(ww[0]) / float(tt[0]) / 1e9, "unit": "Gpixel-iter/s",
                      "h2d_bytes_per_step": int((nx + rows) * 8),
                      "d2h_bytes_per_step": int(rows * nx * (12 if with_pot else 4)) + lines_bytes,
                      "ms_per_step": 1e3 * float(tt[0]) / steps,
                      "boundary_vertices": int(lines.lengths().max()) if lines is not None and len(lines) else 0,
                      "boundary_sha256": lines_digest(lines.verts, lines.offsets) if lines is not None else None,
                      "host_dwell_matches_device": bool(int(host_crc[0]) == int(dev_crc[0])),
                      "phases_ms_rank0_last_step": (job.last_phases_ms if job else None),
                      "api": ("lm_boundary_sample (pinned numpy buffers in, dwell grid + ordered boundary polylines out)" if world == 1 else
                              "sharding.ShardedBoundary.run: lm_shard_escape (pinned numpy buffers; block kept in HBM) + NCCL edge rows + "
                              "lm_contour_records_dev + NCCL send of the records to rank 0 + lm_contour_link_dev")}
    del dwell_d, field_d, rec_d
    state.clear()
    torch.cuda.empty_cache()
    return out


def f32_leg(cx: Ctx, name: str = "cfg3") -> dict:
    """The optional single-precision variant of K1 (same persistent kernel in binary32) on the headline grid, device
    resident: throughput, share of the FP32 issue peak, and its stated-tolerance comparison with the fp64 dwell grid."""
    import torch
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim
    w = WORKLOADS[name]
    res, mi = w["res"], w["max_iter"]
    P = lambda t: C.c_void_p(t.data_ptr())
    xs = torch.from_numpy(np.linspace(w["xlim"][0], w["xlim"][1], res)).to(cx.dev)
    ys = torch.from_numpy(np.linspace(w["ylim"][0], w["ylim"][1], res)).to(cx.dev)
    d64 = torch.empty((res, res), dtype=torch.int32, device=cx.dev); d32 = torch.empty_like(d64)
    wk = torch.zeros(1, dtype=torch.int64, device=cx.dev)
    _shim.call("lm_escape_grid_f64_dev", P(xs), res, P(ys), res, mi, 2.0, 0, P(d64), None, None, None, cx.stream)
    ms = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _shim.call("lm_escape_grid_f32_dev", P(xs), res, P(ys), res, mi, 2.0, P(d32), P(wk), cx.stream)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms[1:]))
    work = int(wk.item())
    info = _shim.DeviceInfo()
    _shim.call("lm_get_device_info", C.byref(info))
    clock_hz = float(measured_peaks().get("sm_max_mhz", info.clock_khz / 1e3)) * 1e6
    peak_instr = info.sm_count * 128 * clock_hz                     # FP32 lane-instructions / s (FMUL / FADD / FFMA issue)
    gpi = work / (t * 1e-3) / 1e9
    mism = float((d64 != d32).double().mean().item())
    mask = float(((d64 == mi) == (d32 == mi)).double().mean().item())
    return {"workload": workload_name(w) + " (dwell only, device resident, 3 timed launches after 1 warm-up)",
            "kernel": "lm_escape_kernel<grid, dwell, float>", "value": gpi, "unit": "Gpixel-iter/s", "ms": t, "dtype": "f32",
            "fp32_pipe_instr_util": gpi * 1e9 * 6 / peak_instr,
            "fp32_peak": f"{info.sm_count} SMs x 128 FP32 lanes x {clock_hz / 1e6:.0f} MHz, 6 FP32 instructions per pixel-iteration",
            "dwell_mismatch_frac_vs_f64": mism, "interior_mask_agreement_vs_f64": mask,
            "stated_tolerance": "mismatch < 0.5 %, interior mask >= 99.95 % on configs 1-3 (12 % / 99.95 % in config 4's zoom)"}


def k4_roofline(cx: Ctx, n: int = 8192, reps: int = 20) -> dict:
    """5-point periodic Laplacian (Laplacian_C-M.py:49-59) on an n x n float64 field resident in HBM: 16 B / pixel."""
    import torch
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim
    U = torch.rand((n, n), dtype=torch.float64, device=cx.dev)
    out = torch.empty_like(U)
    P = lambda t: C.c_void_p(t.data_ptr())
    for _ in range(3):
        _shim.call("lm_laplacian5_periodic_dev", P(U), n, n, 1e-3, P(out), cx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        _shim.call("lm_laplacian5_periodic_dev", P(U), n, n, 1e-3, P(out), cx.stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = 16.0 * n * n / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "laplacian5_kernel (periodic 5-point stencil, float64)", "achieved": gbs, "peak": cx.hbm_gbs,
            "unit": "GB/s", "frac": gbs / cx.hbm_gbs, "algorithmic_bytes_per_unit": 16, "ms": ms,
            "workload": f"{n} x {n} float64 field ({8 * n * n / 1e6:.0f} MB in, same out: larger than L2), {reps} launches back to back"}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    import torch
    import torch.distributed as dist
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build

    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    _shim.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cx = Ctx()
    cx.rank, cx.world, cx.dev, cx.lib = rank, world, dev, _shim.load()
    cx.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # ---- FP64 peak for the roofline (not in MEASURED_PEAKS.json); HBM peak from the driver's file
    peak_tflops = C.c_double(0.0); mix = C.c_double(0.0)
    _shim.call("lm_probe_fp64_peak", 2000, C.byref(peak_tflops), C.byref(mix))
    cx.peak_tflops = peak_tflops.value
    mp = measured_peaks()
    cx.hbm_gbs = float(mp.get("hbm_gbs", 6500.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in mp else "fallback 6500 GB/s (B200_PROFILING.md)"

    if world > 1:
        # process warm-up: the first torch / NCCL kernels of a process load lazily (~150 ms); the plan below is timed
        from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import sharding as _sh
        _sh.plan_row_cuts(np.linspace(-2.0, 1.0, 64), np.linspace(-1.5, 1.5, 64), 50, world, device=dev)
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("LM_BENCH_NO_SAMPLER"):
        sampler.start()             # before the warm-up (same load): nvidia-smi needs ~0.2 s for its first sample
    main = grid_leg(cx, args.workload, args.steps, args.warmup, want_e2e=not args.no_e2e, refine=args.refine)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the other grid configs of BASELINE.json, one timed step each (builder-visible in round 1 only)
    legs = {}
    if args.workload == "cfg3" and not args.no_legs:
        for nm in ("cfg1", "cfg2", "cfg4"):
            try:
                g = grid_leg(cx, nm, steps=1 if nm == "cfg4" else 3, warmup=1 if nm == "cfg4" else 3, want_e2e=(nm != "cfg4") and not args.no_e2e)
                legs[nm] = {"workload": workload_name(g["w"]), "value": g["value"], "unit": "Gpixel-iter/s", "ms_per_step": g["ms_per_step"],
                            "steps": 1 if nm == "cfg4" else 3, "pixel_iters_per_step": g["total_work"] // (1 if nm == "cfg4" else 3),
                            "roofline_frac": g["roofline"]["frac"], "fp64_pipe_instr_util": g["roofline"]["fp64_pipe_instr_util"],
                            "k1_gpixel_iter_per_s": g["roofline"]["k1_gpixel_iter_per_s"],
                            "balance_estimate_only": g["balance_estimate_only"], "balance_measured": g["balance_measured"],
                            "setup_ms": g["setup_ms"], "dwell_crc": g["dwell_crc"], "boundary_sha256": g.get("boundary_sha256"),
                            "with_potential": g["with_pot"], "e2e": g.get("e2e")}
            except Exception as e:      # the headline line must survive a failure of a secondary leg
                legs[nm] = {"error": f"{type(e).__name__}: {e}"}

    if args.workload == "cfg3" and not args.no_legs and world == 1:
        try:
            legs["cfg3_f32_variant"] = f32_leg(cx)
        except Exception as e:
            legs["cfg3_f32_variant"] = {"error": f"{type(e).__name__}: {e}"}

    sub = None
    try:
        sub = {"k2_records": main["k2_roofline"], "k4_stencil": k4_roofline(cx) if rank == 0 or world == 1 else None,
               "hbm_peak_source": hbm_src}
    except Exception as e:
        sub = {"error": f"{type(e).__name__}: {e}"}

    # ---- the other half of BASELINE.json's metric: Lucas roots/s (K3) on the config-5 batch, sharded by polynomial
    lucas_roots = None
    if not args.no_roots:
        try:
            r = measure_roots(args, rank, world, dev, cx.stream, full_fields=False, peak_tflops=cx.peak_tflops)
            lucas_roots = {k: r[k] for k in ("value", "unit", "ms_per_step", "roots", "mean_sweeps", "not_converged", "roofline") if k in r}
            lucas_roots["workload"] = r["config"]["workload"]
            if "e2e" in r:
                lucas_roots["e2e"] = r["e2e"]
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                lucas_roots["cpu_numpy"] = cpu_roots_baseline(8.0)
        except Exception as e:          # the headline line must survive a failure of the secondary leg
            lucas_roots = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baselines beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        w = main["w"]
        every = pick_cpu_stride(w, target_s=15.0)
        work, dt, threads, nrows = cpu_pass(w, every)
        cpu = {"value": work / dt / 1e9, "unit": "Gpixel-iter/s", "cores": threads, "kind": "port",
               "sample": f"every {every}th row ({nrows} of {w['res']} rows, full width), one pass, {dt:.1f} s",
               "python_loop": python_loop_baseline(w, 8.0)}

    if rank == 0:
        w = main["w"]
        plan = main["plan"]
        line = {
            "metric": "gpixel_iter_per_s_fp64", "value": main["value"], "unit": "Gpixel-iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(w), "pixel_iters_per_step": main["total_work"] // args.steps,
                       "step": "K1 dwell grid" + (" + smooth potential" if main["with_pot"] else "") +
                               " + K2 crossing records + ordered boundary polylines (device linker)" +
                               ((" + NCCL all-gather of shard-edge rows" + (" and of the potential field" if main["with_pot"] else "") +
                                 " + NCCL send of the records to rank 0") if world > 1 else ""),
                       "sharding": (f"contiguous row blocks cut once at equal estimated cost ({plan['model']}), cuts={main['cuts']}" +
                                    (f", then refined {main['refined']}x from measured per-rank K1 times" if main["refined"] else "")),
                       "l2": "FP64-bound; per step every rank writes its dwell block (>= L2 for cfg2/cfg3) and reads 2*res coordinates"},
            "roofline": main["roofline"], "sub_rooflines": sub, "cpu_baseline": cpu, "e2e": main.get("e2e"),
            "gpu_launches": main["gpu_launches"], "clocks": clocks,
            "k1_ms_per_step_max_rank": main["k1_ms_per_step_max_rank"],
            "balance_estimate_only": main["balance_estimate_only"], "balance_measured": main["balance_measured"],
            "setup_ms": main["setup_ms"], "setup": plan.get("setup"),
            "dwell_crc": main["dwell_crc"], "boundary_sha256": main.get("boundary_sha256"),
            "boundary_lines": main.get("boundary_lines"), "boundary_vertices": main.get("boundary_vertices"),
            "boundary_vertices_total": main.get("boundary_vertices_total"), "n_records": main["n_records"],
            "step_ms_rank0": main["step_ms_rank0"], "gap_ms_rank0": main["gap_ms_rank0"],
            "lucas_roots": lucas_roots, "legs": legs or None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# Lucas-Loci cloud (BASELINE.json configs[4]; the "Lucas roots/s vs CPU numpy" half of the metric)
# ---------------------------------------------------------------------------------------------
def _eigvals_stack(args):
    """np.linalg.eigvals on the stacked companions of one degree (the reference's own call, batched)."""
    top, d = args
    M = np.zeros((top.shape[0], d, d))
    M[:, 0, :] = top[:, :d]
    idx = np.arange(1, d)
    M[:, idx, idx - 1] = 1.0
    return int(np.linalg.eigvals(M).size)


def cpu_roots_pass(npoly_sample: int, procs: int):
    """numpy (LAPACK dgeev) on the first npoly_sample polynomials of chunk 0, `procs` worker processes.
    -> (roots, seconds)"""
    import multiprocessing as mp
    per_chunk = CFG5["npoly"] // CFG5["chunks"]
    parts = [cfg5_chunk(k) for k in range(min(CFG5["chunks"], (npoly_sample + per_chunk - 1) // per_chunk))]
    top = np.concatenate([p[0] for p in parts])[:npoly_sample]
    deg = np.concatenate([p[1] for p in parts])[:npoly_sample]
    jobs = []
    for d in range(2, CFG5["maxdeg"] + 1):
        sel = np.where(deg == d)[0]
        for part in np.array_split(sel, max(1, procs // 2)):
            if part.size:
                jobs.append((top[part], d))
    jobs.sort(key=lambda j: -j[0].shape[0] * j[1] ** 3)
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_eigvals_stack, jobs[:procs])                       # spin the workers up
        t0 = time.perf_counter()
        roots = sum(pool.map(_eigvals_stack, jobs, chunksize=1))
        dt = time.perf_counter() - t0
    return roots, dt


def cpu_roots_baseline(target_s: float = 10.0) -> dict:
    procs = os.cpu_count() or 1
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    roots, dt = cpu_roots_pass(4000 * procs, procs)
    n = int(min(max(4000 * procs * target_s / max(dt, 1e-3), 4000 * procs), CFG5["npoly"] // 2))
    roots, dt = cpu_roots_pass(n, procs)
    return {"value": roots / dt / 1e6, "unit": "Mroots/s", "cores": procs, "kind": "reference",
            "sample": f"np.linalg.eigvals on the stacked companion matrices of the first {n} polynomials of the "
                      f"config-5 batch ({roots} roots), {procs} worker processes, {dt:.1f} s"}


def measure_roots(args, rank, world, dev, stream, full_fields: bool, peak_tflops: float | None = None):
    """K3 (+ cloud compaction) over this rank's slice of the 10^7-polynomial batch, device resident; optionally the
    field stage (K1d, K4a with an NCCL all-reduce of the per-cell sums, K4).  -> dict for the JSON line."""
    import torch
    import torch.distributed as dist
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, lucas, sharding
    P = lambda t: C.c_void_p(t.data_ptr())
    cuts = sharding.item_slices(CFG5["chunks"], world)
    mine = range(cuts[rank], cuts[rank + 1])
    parts = [cfg5_chunk(k) for k in mine]
    maxdeg = CFG5["maxdeg"]
    top_h = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, maxdeg))
    deg_h = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int32)
    npoly = int(deg_h.size)
    nroots = int(deg_h.sum())
    top = torch.from_numpy(top_h).to(dev); deg = torch.from_numpy(deg_h).to(dev)
    re = torch.empty((max(npoly, 1), maxdeg), dtype=torch.float64, device=dev); im = torch.empty_like(re)
    kept = torch.empty(max(npoly, 1), dtype=torch.int32, device=dev)
    iters = torch.empty(max(npoly, 1), dtype=torch.int32, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    px = torch.empty(max(nroots, 1), dtype=torch.float64, device=dev); py = torch.empty_like(px)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)

    def step():
        _shim.call("lm_roots_batched_dev", P(top), P(deg), npoly, maxdeg, 1, 1e-12, P(re), P(im), P(kept), P(iters), P(status), stream)
        _shim.call("lm_cloud_compact_dev", P(re), P(im), P(kept), npoly, maxdeg, P(px), P(py), nroots, P(cnt), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    tot = torch.tensor([nroots, npoly, int(cnt.item()), int(status[0].item()), int(iters[:npoly].sum().item())], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t[0]) / args.steps
    roots_total, npoly_total, pts_total, failed, sweeps = (int(v) for v in tot)
    out = {"metric": "lucas_roots_per_s", "value": roots_total / ms / 1e3, "unit": "Mroots/s", "ms_per_step": ms,
           "config": {"workload": f"{npoly_total} generalized-Lucas characteristic polynomials, degree 2..{maxdeg}, first-row entries in "
                                  "{0,1,2} (BASELINE.json configs[4]), K3 roots + cloud compaction, device resident",
                      "sharding": f"{CFG5['chunks']} independently seeded chunks dealt contiguously to the ranks, no collective"},
           "roots": roots_total, "cloud_points": pts_total, "mean_sweeps": sweeps / max(npoly_total, 1), "not_converged": failed,
           "gpu_launches": 9 * args.steps}
    # K3 against the FP64 peak: SURVEY 8d's flop model (28 flops per inner step, 2 d inner steps per root update:
    # d Horner steps for p, p' and the error bound + d Aberth terms; root updates ~ roots x sweeps, an UPPER bound
    # because converged roots drop out of the later sweeps) and the FP64 instructions the kernel issues per inner
    # step (SASS of the two unrolled loops: 9 per Horner step, 6 per Aberth term)
    if peak_tflops is None:
        pk = C.c_double(0.0); mx = C.c_double(0.0)
        _shim.call("lm_probe_fp64_peak", 2000, C.byref(pk), C.byref(mx))
        peak_tflops = pk.value
    sw_d2 = (iters[:npoly].abs().double() * deg[:npoly].double() ** 2).sum().reshape(1) if npoly else torch.zeros(1, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sw_d2, op=dist.ReduceOp.SUM)
    inner_steps = 2.0 * float(sw_d2[0])                      # sum over polynomials of sweeps * 2 d^2
    flops = 14.0 * inner_steps                               # 28 d^2 per sweep = 14 per inner step
    instr = 7.5 * inner_steps                                # (9 + 6) / 2 FP64 instructions per inner step
    out["roofline"] = {"bound": "fp64", "kernel": "roots_pool_kernel (Aberth-Ehrlich, 4 lanes per polynomial) + sort / compaction kernels",
                       "unit": "TFLOP/s", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak_tflops * world,
                       "frac": flops / (ms * 1e-3) / 1e12 / (peak_tflops * world),
                       "fp64_pipe_instr_util": instr / (ms * 1e-3) / (peak_tflops * world * 1e12 / 2),
                       "flop_model": "28 d^2 flops per polynomial and sweep (SURVEY 8d) x the sweeps each polynomial ran: an upper bound, "
                                     "converged roots leave the later sweeps",
                       "peak_source": "lm_probe_fp64_peak, measured live"}

    # e2e: host numpy arrays in, cloud on the host out, through the fused host-buffer call
    if not args.no_e2e:
        # the workload's first rows are integers in {0, 1, 2}: hand them over as int8 (exact; widened on the device)
        compact = not args.roots_f64_rows and bool(np.array_equal(top_h, top_h.astype(np.int8)))
        top_p = _shim.pinned_empty(top_h.shape, np.int8 if compact else np.float64); top_p[...] = top_h
        deg_p = _shim.pinned_empty(deg_h.shape, np.int32); deg_p[...] = deg_h
        cre_p = _shim.pinned_empty(max(nroots, 1), np.float64); cim_p = _shim.pinned_empty(max(nroots, 1), np.float64)
        lucas.cloud_fields(top_p, deg_p, cloud_out=(cre_p, cim_p))              # warm-up (workspace allocation)
        barrier()
        t0 = time.perf_counter()
        res = lucas.cloud_fields(top_p, deg_p, cloud_out=(cre_p, cim_p))
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": roots_total / float(dt[0]) / 1e6, "unit": "Mroots/s", "h2d_bytes_per_step": int(top_p.nbytes + deg_h.nbytes),
                      "d2h_bytes_per_step": int(res["n_points"] * 16), "ms_per_step": 1e3 * float(dt[0]),
                      "api": ("lm_lucas_cloud_fields_i8 (pinned int8 first rows in" if compact else "lm_lucas_cloud_fields (pinned float64 first rows in")
                             + ", cloud of 1/lambda out to pinned numpy buffers)"}

    if full_fields:
        g = torch.linspace(-2.0, 2.0, CFG5["grid"], dtype=torch.float64, device=dev)
        ncell = CFG5["grid"] ** 2
        sums = torch.empty(ncell, dtype=torch.float64, device=dev); U = torch.empty_like(sums); lap = torch.empty_like(sums)
        gpot = torch.empty_like(px); it = torch.empty(px.numel(), dtype=torch.int64, device=dev)
        work = torch.zeros(1, dtype=torch.int64, device=dev)
        n_local = int(cnt.item())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier()
        ev[0].record()
        _shim.call("lm_escape_points_f64_dev", P(px), P(py), n_local, CFG5["pot_max_iter"], CFG5["pot_radius"], P(gpot), P(it),
                   None, None, P(work), stream)
        ev[1].record()
        _shim.call("lm_log_potential_sums_dev", P(px), P(py), n_local, P(g), CFG5["grid"], P(g), CFG5["grid"], CFG5["eps"], 0,
                   P(sums), stream)
        ev[2].record()
        _, n_total = sharding.allreduce_field_sums(sums, n_local)
        _shim.call("lm_log_potential_finish_dev", P(sums), ncell, n_total, 0, P(U), stream)
        _shim.call("lm_laplacian5_periodic_dev", P(U), CFG5["grid"], CFG5["grid"], 4.0 / (CFG5["grid"] - 1), P(lap), stream)
        ev[3].record()
        barrier()
        tt = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], dtype=torch.float64, device=dev)
        ww = torch.tensor([int(work.item()), n_local * ncell], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        peak_tflops = C.c_double(0.0); mix = C.c_double(0.0)
        _shim.call("lm_probe_fp64_peak", 2000, C.byref(peak_tflops), C.byref(mix))
        pairs_per_s = n_local * ncell / (ev[1].elapsed_time(ev[2]) * 1e-3)
        out["field_stage"] = {
            "k1d_batch_potential": {"ms": float(tt[0]), "gpixel_iter_per_s": int(ww[0]) / float(tt[0]) / 1e6, "max_iter": CFG5["pot_max_iter"]},
            "k4a_log_potential": {"ms": float(tt[1]), "tpairs_per_s": int(ww[1]) / float(tt[1]) / 1e9, "grid": f"{CFG5['grid']}^2 over [-2,2]^2",
                                  "roofline": {"bound": "fp64", "kernel": "logpot_partial_kernel<SUM_SQRT, fast>", "unit": "T FP64 instr/s",
                                               "achieved": pairs_per_s * 4.5 / 1e12, "peak": peak_tflops.value / 2,
                                               "frac": pairs_per_s * 4.5 / 1e12 / (peak_tflops.value / 2),
                                               "algorithmic_fp64_instr_per_pair": 4.5}},
            "allreduce_finish_laplacian_ms": float(tt[2]), "field_checksum": float(U.sum().item()), "laplacian_abs_max": float(lap.abs().max().item()),
            "collective": "NCCL all-reduce(sum) of the 400^2 per-cell sums + cloud sizes" if world > 1 else "none (N=1)"}
    return out


def run_cfg5(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            cpu = cpu_roots_baseline(10.0 * max(args.steps, 1) / 2)
            line = {"impl": "reference", "metric": "lucas_roots_per_s", "value": cpu["value"], "unit": "Mroots/s", "n_gpus": args.gpus,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": "generalized-Lucas characteristic polynomials, degree 2..25 (BASELINE.json configs[4])",
                               "sample": cpu["sample"]},
                    "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "Mroots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            print(json.dumps(line), flush=True)
        return
    import torch
    import torch.distributed as dist
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank); _shim.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    out = measure_roots(args, rank, world, dev, stream, full_fields=True)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_roots_baseline()
        line = {"metric": out["metric"], "value": out["value"], "unit": out["unit"], "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": out["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": out["config"],
                "roofline": out["field_stage"]["k4a_log_potential"]["roofline"], "cpu_baseline": cpu, "e2e": out.get("e2e"),
                "gpu_launches": out["gpu_launches"], "clocks": clocks,
                "lucas": {k: out[k] for k in ("roots", "cloud_points", "mean_sweeps", "not_converged")},
                "field_stage": out["field_stage"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.workload == "cfg5":
        run_cfg5(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 hot path (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1]

Metric (BASELINE.json): G pixel-iterations/s in fp64, work unit = sum over pixels of
min(dwell+1, max_iter) counted exactly by the kernel.  A step is one pass of the boundary stage
over the whole window: K1 (escape-time dwell grid, fp64, bit-exact) followed by K2 (level-set
classification -> crossing records); with N > 1 the rows are sharded over the ranks at equal
estimated work and the shard-edge dwell rows are all-gathered over NCCL for K2's halo.

  value     whole-job throughput with inputs/outputs resident in HBM (device-timed, max over ranks)
  e2e       the same through the host-buffer C ABI (lm_escape_grid_f64 on pinned numpy buffers +
            K2 + gather of the records + ordered linking on rank 0): H2D and D2H inside the region
  roofline  K1 against the FP64 peak measured live by lm_probe_fp64_peak (MEASURED_PEAKS.json has
            no FP64 entry): achieved = pixel_iters x 8 flops / K1 device time
  cpu_baseline / --impl reference
            the CPU port of the reference (oracle/lm_oracle.c, pthreads over all host cores) on a
            bounded row sample of the same workload.  The reference itself is pure Python
            (0.84 M pixel-iter/s on one core) and cannot run the workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
    # BASELINE.json configs[2] (the grid the target is quoted on; fits one GPU), [1] and [0]
    "cfg3": dict(xlim=(-2.1, 0.9), ylim=(-1.5, 1.5), res=32768, max_iter=10000, level=0.96),
    "cfg2": dict(xlim=(-2.1, 0.9), ylim=(-1.5, 1.5), res=8192, max_iter=2000, level=0.96),
    "cfg1": dict(xlim=(-2.1, 0.9), ylim=(-1.5, 1.5), res=2000, max_iter=500, level=0.96),
    "cfg4": dict(xlim=(-0.755, -0.735), ylim=(0.10, 0.12), res=16384, max_iter=100000, level=0.96),
}
FLOPS_PER_PIXEL_ITER = 8          # 3 mul + 5 add under the reference's unfused semantics


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg3")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(w: dict) -> str:
    return (f"mandelbrot_boundary_sample --xlim {w['xlim'][0]} {w['xlim'][1]} --ylim {w['ylim'][0]} {w['ylim'][1]} "
            f"--res {w['res']} --max_iter {w['max_iter']} --level {w['level']}")


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of rows
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
        "note": "CPU port of the reference's scalar loop (oracle/lm_oracle.c), all host threads; "
                "the Python reference itself runs at ~0.84e-3 Gpixel-iter/s on one core",
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


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    import torch
    import torch.distributed as dist
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build, contour, sharding

    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    _shim.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _shim.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    w = WORKLOADS[args.workload]
    res, max_iter = w["res"], w["max_iter"]
    level = w["level"] * max_iter
    xs = np.linspace(w["xlim"][0], w["xlim"][1], res)
    ys = np.linspace(w["ylim"][0], w["ylim"][1], res)
    nx, ny = xs.size, ys.size

    # ---- row shard of this rank (contiguous block cut at equal estimated work)
    if world > 1:
        profile = sharding.coarse_row_profile(xs, ys, max_iter)
        cuts = sharding.balanced_row_cuts(profile, world)
        balance = sharding.parallel_efficiency(profile, cuts)
    else:
        cuts, balance = [0, ny], 1.0
    r0, r1 = cuts[rank], cuts[rank + 1]
    rows = r1 - r0
    has_halo = rank < world - 1
    ys_block = np.ascontiguousarray(ys[r0:r1 + (1 if has_halo else 0)])

    xs_d = torch.from_numpy(xs).to(dev)
    ys_d = torch.from_numpy(ys_block).to(dev)
    dwell_d = torch.empty((rows + 1, nx), dtype=torch.int32, device=dev)      # +1: halo row slot
    work_d = torch.zeros(1, dtype=torch.int64, device=dev)
    rec_cap = max(int(0.01 * rows * nx) + 4096, 1 << 16)
    records = np.empty((rec_cap, 8), dtype=np.int64)
    n_rec = C.c_int64(0)
    launches = {"n": 0}

    def k1():
        _shim.call("lm_escape_grid_f64_dev", C.c_void_p(xs_d.data_ptr()), nx, C.c_void_p(ys_d.data_ptr()), rows,
                   max_iter, 2.0, 0, C.c_void_p(dwell_d.data_ptr()), None, None, C.c_void_p(work_d.data_ptr()), stream)
        launches["n"] += 1

    def k2():
        nonlocal records
        if world > 1:
            firsts = sharding.exchange_first_rows(dwell_d[0])
            if has_halo:
                dwell_d[rows].copy_(firsts[rank + 1])
        nrows_k2 = rows + (1 if has_halo else 0)
        while True:
            rc = lib.lm_contour_classify_dev(C.c_void_p(dwell_d.data_ptr()), _shim.ptr(xs), nx, _shim.ptr(ys_block), nrows_k2,
                                             r0, float(level), _shim.ptr(records), records.shape[0], C.byref(n_rec), stream)
            if rc == _shim.LM_E_CAP:
                records = np.empty((n_rec.value + 1024, 8), dtype=np.int64)
                continue
            _shim.check(rc)
            break
        launches["n"] += 3 if n_rec.value else 2
        return records[: n_rec.value]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP64 peak for the roofline (not in MEASURED_PEAKS.json)
    peak_tflops = C.c_double(0.0); mix = C.c_double(0.0)
    _shim.call("lm_probe_fp64_peak", 2000, C.byref(peak_tflops), C.byref(mix))

    for _ in range(args.warmup):
        k1(); k2()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches["n"] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 + 2 * args.steps)]
    work_steps = []
    ev[0].record()
    for k in range(args.steps):
        ev[2 + 2 * k].record()
        k1()
        ev[3 + 2 * k].record()
        k2()
        work_steps.append(int(work_d.item()))
    ev[1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev[0].elapsed_time(ev[1])
    k1_ms = sum(ev[2 + 2 * k].elapsed_time(ev[3 + 2 * k]) for k in range(args.steps))
    my_work = sum(work_steps)
    t = torch.tensor([elapsed_ms, k1_ms], dtype=torch.float64, device=dev)
    wk = torch.tensor([my_work], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(wk, op=dist.ReduceOp.SUM)
    elapsed_ms, k1_ms_max = float(t[0]), float(t[1])
    total_work = int(wk[0])
    value = total_work / (elapsed_ms * 1e-3) / 1e9
    gpu_launches = launches["n"]

    # ---- roofline of the dominant kernel (K1) on this rank
    k1_gpi = my_work / (k1_ms * 1e-3) / 1e9
    achieved_tflops = k1_gpi * 1e9 * FLOPS_PER_PIXEL_ITER / 1e12
    roofline = {
        "bound": "fp64", "kernel": "lm_escape_kernel<grid, dwell>", "achieved": achieved_tflops, "peak": peak_tflops.value,
        "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops.value,
        "peak_source": "lm_probe_fp64_peak (dependent-free DFMA loop, 2 flops/DFMA), measured live; "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "algorithmic_flops_per_unit": FLOPS_PER_PIXEL_ITER,
        "fp64_pipe_instr_util": k1_gpi * 1e9 * 6 / (peak_tflops.value * 1e12 / 2),
        "k1_gpixel_iter_per_s": k1_gpi, "traffic": None,
    }

    # ---- e2e through the host-buffer C ABI
    e2e = None
    if not args.no_e2e:
        out = _shim.pinned_empty((rows, nx), np.int32)
        xs_p = _shim.pinned_empty(nx, np.float64); xs_p[:] = xs
        ys_p = _shim.pinned_empty(rows, np.float64); ys_p[:] = ys[r0:r1]
        st = _shim.Stats()

        def e2e_step():
            if world == 1:
                # the fused host-buffer call (compute_grid + extract_contour of the script's main()):
                # H2D of xs/ys, chunked K1, dwell grid copied back to the pinned host buffer while
                # K1/K2 still run, K2 records -> ordered polylines on the host
                lines, stx = contour.boundary_sample(xs_p, ys_p, max_iter, level, dwell_out=out)
                return stx["work_units"], lines
            # N > 1: K1 through the host-buffer entry point on this rank's rows, then the shard goes
            # back up for the halo exchange (NCCL) and K2, records are gathered and linked on rank 0
            _shim.call("lm_escape_grid_f64", _shim.ptr(xs_p), nx, _shim.ptr(ys_p), rows, max_iter, 2.0, 0,
                       _shim.ptr(out), None, None, C.byref(st))
            _shim.call("lm_memcpy_h2d", C.c_void_p(dwell_d.data_ptr()), _shim.ptr(out), out.nbytes, stream)
            recs_local = k2()
            allrec = sharding.gather_records(recs_local, dev, 0)
            lines = contour.link_records(allrec, xs, ys, level) if rank == 0 else None
            return st.work_units, lines

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_work = 0
        lines = None
        for _ in range(args.steps):
            wu, lines = e2e_step()
            e2e_work += int(wu)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        ww = torch.tensor([e2e_work], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        n_vertices = int(max((len(l) for l in lines), default=0)) if lines else 0
        e2e = {"value": int(ww[0]) / float(tt[0]) / 1e9, "unit": "Gpixel-iter/s",
               "h2d_bytes_per_step": int((nx + rows) * 8 + (rows * nx * 4 if world > 1 else 0)),
               "d2h_bytes_per_step": int(rows * nx * 4 + n_rec.value * 64), "ms_per_step": 1e3 * float(tt[0]) / args.steps,
               "boundary_vertices": n_vertices,
               "api": ("lm_boundary_sample (pinned numpy buffers in, dwell grid + ordered boundary polylines out)" if world == 1 else
                       "lm_escape_grid_f64 (pinned numpy buffers) + lm_contour_classify_dev + lm_contour_link")}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        every = pick_cpu_stride(w, target_s=15.0)
        work, dt, threads, nrows = cpu_pass(w, every)
        cpu = {"value": work / dt / 1e9, "unit": "Gpixel-iter/s", "cores": threads, "kind": "port",
               "sample": f"every {every}th row ({nrows} of {res} rows, full width), one pass, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": "gpixel_iter_per_s_fp64", "value": value, "unit": "Gpixel-iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(w), "pixel_iters_per_step": total_work // args.steps,
                       "step": "K1 dwell grid + K2 crossing records" + (" + NCCL all-gather of shard-edge rows" if world > 1 else ""),
                       "sharding": f"contiguous row blocks at equal estimated work, cuts={cuts}, balance={balance:.3f}",
                       "l2": "FP64-bound; per step every rank writes its dwell block (>= L2 for cfg2/cfg3) and reads 2*res coordinates"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
            "k1_ms_per_step_max_rank": k1_ms_max / args.steps,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Host-link study under torchrun (one rank per GPU): device-to-host bandwidth of page-locked copies, one rank at a time
and all ranks at once, with and without binding every rank to the CPUs NVML reports as local to its GPU
(--bind; the page-locked buffer is allocated after the binding, so it lands on that NUMA node)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, torch.distributed as dist
from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, device_affinity
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
bound = None
if "--bind" in sys.argv:
    bound = device_affinity.bind_to_gpu(local)
torch.cuda.set_device(local); _shim.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 512 << 20
host = _shim.pinned_empty(nbytes, np.uint8); host[:] = 1
devb = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
def copy_gbs(reps=5, h2d=False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d: _shim.call("lm_memcpy_h2d", devb.data_ptr(), _shim.ptr(host), nbytes, None)
        else: _shim.call("lm_memcpy_d2h", _shim.ptr(host), devb.data_ptr(), nbytes, None)
    torch.cuda.synchronize()
    return reps * nbytes / (time.perf_counter() - t0) / 1e9
copy_gbs(2)
solo = torch.zeros(world, dtype=torch.float64, device="cuda")
for r in range(world):
    dist.barrier()
    if r == rank: solo[rank] = copy_gbs()
dist.barrier(); dist.all_reduce(solo)
dist.barrier(); allv = torch.zeros(world, dtype=torch.float64, device="cuda"); allv[rank] = copy_gbs(); dist.all_reduce(allv)
dist.barrier(); allh = torch.zeros(world, dtype=torch.float64, device="cuda"); allh[rank] = copy_gbs(h2d=True); dist.all_reduce(allh)
if rank == 0:
    print(f"bind={bound is not None} cpus(rank0)={bound}", flush=True)
    print("D2H solo    GB/s per rank:", [round(float(v), 1) for v in solo])
    print("D2H all at once GB/s per rank:", [round(float(v), 1) for v in allv], "aggregate", round(float(allv.sum()), 1))
    print("H2D all at once GB/s per rank:", [round(float(v), 1) for v in allh], "aggregate", round(float(allh.sum()), 1))
dist.destroy_process_group()

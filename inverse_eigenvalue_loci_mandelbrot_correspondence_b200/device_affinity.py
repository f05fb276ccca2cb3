"""Bind a rank to the CPUs that are local to its GPU (one process per GPU, 8 GPUs on a two-socket host).

The host-buffer entry points move gigabytes between page-locked host memory and the GPU (the dwell block of a shard,
the Lucas-Loci cloud).  Page-locked memory is placed on the NUMA node of the thread that allocates it; with the launcher's
default placement several ranks end up allocating on the far socket and their copies cross the inter-socket link.  NVML
knows which CPUs sit next to each GPU (`nvmlDeviceGetCpuAffinity`); binding the process to them BEFORE any page-locked
allocation keeps buffers and copy threads on the GPU's own node.  Plumbing only (no reference counterpart)."""
from __future__ import annotations

import os


def local_cpus(gpu_index: int) -> list[int]:
    """CPUs NVML reports as local to the GPU (empty when NVML or the query is unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        pynvml.nvmlShutdown()
        return [c for c in cpus if c < ncpu]
    except Exception:
        return []


def bind_to_gpu(gpu_index: int) -> list[int] | None:
    """Restrict this process to the GPU's local CPUs (intersected with what the process is allowed to use).
    Returns the CPU list it bound to, or None when nothing was changed."""
    cpus = local_cpus(gpu_index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    allowed = os.sched_getaffinity(0)
    use = sorted(set(cpus) & set(allowed))
    if not use or len(use) == len(allowed):
        return None
    try:
        os.sched_setaffinity(0, use)
    except OSError:
        return None
    return use

"""torchrun worker of tests/test_gpu_sharded.py: sharding.ShardedBoundary on every rank (NCCL), results to an .npz."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    out_dir = Path(sys.argv[1])
    res, mi, with_pot = int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
    import torch
    import torch.distributed as dist
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, sharding
    local = int(os.environ["LOCAL_RANK"]); rank = int(os.environ["RANK"])
    torch.cuda.set_device(local); _shim.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res + 3)
    job = sharding.ShardedBoundary(xs, ys, mi, 0.96 * mi, with_potential=with_pot)
    lines = job.run()
    lines2 = job.run()                      # a second pass over the same buffers must give the same answer
    save = {"cuts": np.array(job.cuts), "dwell": np.array(job.dwell), "r0": job.r0}
    if with_pot:
        save["potential"] = np.array(job.potential)
        save["full_potential"] = job.full_potential.cpu().numpy()
    if rank == 0:
        assert lines is not None and lines2 is not None
        assert np.array_equal(lines.verts, lines2.verts) and np.array_equal(lines.offsets, lines2.offsets)
        save["verts"] = lines.verts; save["offsets"] = lines.offsets
    else:
        assert lines is None
    np.savez(out_dir / f"rank{rank}.npz", **save)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""GPU, >= 2 devices: the row-sharded boundary stage (sharding.ShardedBoundary under torchrun, NCCL: halo rows,
potential all-gather, GPU-to-GPU record gather, device linker on rank 0) against the single-GPU fused call.
Skipped on a box with one GPU (the driver's N=1 tier); the CPU / gloo coverage of the same logic is
tests/test_sharding_gloo.py."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("world,res,mi,with_pot", [(2, 1536, 400, False), (2, 1024, 300, True)])
def test_sharded_equals_unsharded(gpu, tmp_path, world, res, mi, with_pot):
    if gpu.shim.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", str(ROOT / "tests" / "_sharded_worker.py"), str(tmp_path), str(res), str(mi), str(int(with_pot))]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    parts = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    xs = np.linspace(-2.1, 0.9, res); ys = np.linspace(-1.5, 1.5, res + 3)
    d = np.empty((ys.size, xs.size), dtype=np.int32)
    pot = np.empty((ys.size, xs.size)) if with_pot else None
    lines, _ = gpu.contour.boundary_sample(xs, ys, mi, 0.96 * mi, dwell_out=d, potential_out=pot)
    cuts = parts[0]["cuts"]
    assert all(np.array_equal(p["cuts"], cuts) for p in parts) and cuts[0] == 0 and cuts[-1] == ys.size
    assert np.array_equal(np.concatenate([p["dwell"] for p in parts]), d)
    assert np.array_equal(parts[0]["verts"], lines.verts) and np.array_equal(parts[0]["offsets"], lines.offsets)
    if with_pot:
        assert np.array_equal(np.concatenate([p["potential"] for p in parts]), pot)
        for p in parts:
            assert np.array_equal(p["full_potential"], pot)

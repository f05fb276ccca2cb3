"""CPU: the evidence helpers of bench.py -- the dwell checksum must be additive over row blocks (so the all-reduced
value is the same for every sharding of a grid), the polyline digest must depend on values and on the line
structure, the workload table must be the one BASELINE.json names."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def test_dwell_checksum_is_additive_over_row_blocks():
    torch = pytest.importorskip("torch")
    import bench
    rng = np.random.default_rng(0)
    d = torch.from_numpy(rng.integers(0, 10001, size=(3000, 257)).astype(np.int32))
    whole = bench.dwell_checksum(d, 0, 257)
    mask = (1 << 64) - 1
    for cuts in ([0, 3000], [0, 1, 3000], [0, 1200, 1201, 2999, 3000], [0, 750, 1500, 2250, 3000]):
        parts = sum(bench.dwell_checksum(d[a:b], a, 257) for a, b in zip(cuts[:-1], cuts[1:]))
        assert parts & mask == whole & mask
    e = d.clone(); e[1234, 56] += 1
    assert bench.dwell_checksum(e, 0, 257) & mask != whole & mask            # one changed pixel changes it
    f = d.clone(); f[[10, 11]] = f[[11, 10]]
    if not torch.equal(f, d):
        assert bench.dwell_checksum(f, 0, 257) & mask != whole & mask        # position dependent: swapped rows change it


def test_lines_digest_depends_on_values_and_structure():
    import bench
    v = np.arange(20, dtype=np.float64).reshape(10, 2)
    a = bench.lines_digest(v, np.array([0, 4, 10]))
    assert a == bench.lines_digest(v.copy(), np.array([0, 4, 10], dtype=np.int64))
    assert a != bench.lines_digest(v, np.array([0, 5, 10]))
    w = v.copy(); w[3, 1] = np.nextafter(w[3, 1], 1e9)
    assert a != bench.lines_digest(w, np.array([0, 4, 10]))


def test_workloads_are_the_configs_of_baseline_json():
    import bench
    cfg = json.loads((ROOT / "BASELINE.json").read_text())["configs"]
    w1, w2, w3, w4 = bench.WORKLOADS["cfg1"], bench.WORKLOADS["cfg2"], bench.WORKLOADS["cfg3"], bench.WORKLOADS["cfg4"]
    assert "--res 2000 --max_iter 500 --level 0.96" in cfg[0] and (w1["res"], w1["max_iter"], w1["level"]) == (2000, 500, 0.96)
    assert "res 8192, max_iter 2000" in cfg[1] and (w2["res"], w2["max_iter"]) == (8192, 2000)
    assert "res 32768 max_iter 10000" in cfg[2] and (w3["res"], w3["max_iter"]) == (32768, 10000)
    assert "16384" in cfg[3] and "100000" in cfg[3] and (w4["res"], w4["max_iter"]) == (16384, 100000)
    assert "10^7" in cfg[4] and bench.CFG5["npoly"] == 10_000_000 and bench.CFG5["maxdeg"] == 25
    for w in (w1, w2, w3):
        assert w["xlim"] == (-2.1, 0.9) and w["ylim"] == (-1.5, 1.5)
    top, deg = bench.cfg5_chunk(3, npoly_total=1600, chunks=16)
    assert top.shape == (100, 25) and deg.min() >= 2 and deg.max() <= 25
    assert (top[np.arange(100), deg - 1] >= 1).all() and set(np.unique(top)) <= {0.0, 1.0, 2.0}
    assert (top[np.arange(25)[None, :] >= deg[:, None]] == 0).all()


def test_traffic_file_names_its_source():
    import bench
    t, src = bench.profiled_traffic("cfg3")
    assert t is not None and 4.2e9 < t < 4.8e9 and src and (ROOT / src.split(" ")[0]).exists()
    assert bench.profiled_traffic("cfg1") == (None, None)

"""The packaged CLI drop-ins for README steps 2 and 3 (construct_boundary_alpha.py, boundary_curvature_localpoly.py) against
the text outputs of the reference's own mains on the same inputs (tests/golden/cli/, made by oracle/gen_cli_golden.py).
Runs last on purpose (file name): the GPU cases exercise whole scripts."""
import shutil
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden" / "cli"


def _read_csv(path):
    lines = Path(path).read_text().splitlines()
    return lines[0], np.array([[float(v) for v in ln.split(",")] for ln in lines[1:]])


def _check_alpha_outputs(out: Path):
    for name in ("construct_boundary.csv", "construct_edges.csv", "construct_meta.txt"):
        assert (out / name).read_bytes() == (GOLD / name).read_bytes(), name
    assert (out / "construct_boundary.png").read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"


def _check_alpha_v2_outputs(out: Path):
    for name in ("construct_v2_boundary.csv", "construct_v2_edges.csv", "construct_v2_meta.txt"):
        assert (out / name).read_bytes() == (GOLD / name).read_bytes(), name


def _check_curvature_outputs(out: Path, rtol: float):
    head, got = _read_csv(out / "loop_curvature.csv")
    head_ref, want = _read_csv(GOLD / "loop_curvature.csv")
    assert head == head_ref and got.shape == want.shape
    assert np.array_equal(got[:, :3], want[:, :3])                        # idx, x, y: the same "%.10g" text
    scale = np.abs(want).max(axis=0)
    assert np.all(np.abs(got - want) <= rtol * scale[None, :])
    a = (out / "loop_summary.txt").read_text().splitlines(); b = (GOLD / "loop_summary.txt").read_text().splitlines()
    assert a[0] == b[0] and len(a) == len(b)
    for la, lb in zip(a[1:], b[1:]):
        ka, va = la.split(": "); kb, vb = lb.split(": ")
        assert ka == kb and abs(float(va) - float(vb)) <= rtol * max(abs(float(vb)), 1.0)
    for name in ("loop_curvature_hist.png", "loop_curvature_overlay.png"):
        assert (out / name).read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"


def test_cli_file_formats_with_oracle_compute_cpu(oracle, tmp_path, monkeypatch):
    """Argument handling and file writers of both scripts, the device calls replaced by the oracle: byte-identical alpha-shape
    files, curvature CSV / summary equal to the reference's to the text precision."""
    from scipy.spatial import Delaunay
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import boundary_curvature_localpoly as bc
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import construct_boundary_alpha as cb
    monkeypatch.setattr(cb, "alpha_shape_edges", lambda P, alpha: oracle.alpha_shape_edges(P, Delaunay(P).simplices, alpha)[2])

    def curv(P, neighbors=7, closed=True, stride=1):
        o = oracle.curvature_localpoly(P, neighbors, closed)
        return o[:, 0], o[:, 1], o[:, 2], dict(xprime=o[:, 3], yprime=o[:, 4], x2=o[:, 5], y2=o[:, 6])
    monkeypatch.setattr(bc, "compute_curvature_localpoly", curv)
    out = tmp_path / "outputs"
    cb.main(["--input_csv", str(GOLD / "construct_points.csv"), "--alpha", "6.0", "--output_prefix", str(out / "construct")])
    _check_alpha_outputs(out)
    # the v2 behaviour (README step 2): nine components at alpha 12, the outer loop kept and resampled to 1500 points
    cb.main(["--input_csv", str(GOLD / "construct_points.csv"), "--alpha", "12.0", "--output_prefix", str(out / "construct_v2"), "--main_loop"])
    _check_alpha_v2_outputs(out)
    bc.main(["--input_csv", str(GOLD / "loop_boundary.csv"), "--output_prefix", str(out / "loop"), "--neighbors", "7"])
    _check_curvature_outputs(out, rtol=1e-7)
    with pytest.raises(SystemExit):                                       # too few points for the window: exit code 2
        short = tmp_path / "short.csv"
        short.write_text("x,y\n0,0\n1,0\n1,1\n")
        bc.main(["--input_csv", str(short), "--output_prefix", str(out / "short")])


@pytest.mark.gpu
def test_cli_dropins_on_the_device(gpu, tmp_path):
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import boundary_curvature_localpoly as bc
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import construct_boundary_alpha as cb
    out = tmp_path / "outputs"
    cb.main(["--input_csv", str(GOLD / "construct_points.csv"), "--alpha", "6.0", "--output_prefix", str(out / "construct")])
    _check_alpha_outputs(out)
    cb.main(["--input_csv", str(GOLD / "construct_points.csv"), "--alpha", "12.0", "--output_prefix", str(out / "construct_v2"), "--main_loop"])
    _check_alpha_v2_outputs(out)
    bc.main(["--input_csv", str(GOLD / "loop_boundary.csv"), "--output_prefix", str(out / "loop"), "--neighbors", "7"])
    _check_curvature_outputs(out, rtol=1e-6)
    with pytest.raises(SystemExit):
        cb.main(["--input_csv", str(GOLD / "construct_points.csv"), "--alpha", "-1.0", "--output_prefix", str(out / "none")])


def _check_samplers(pt, golden):
    np.random.seed(11)
    got = pt.sample_mandelbrot_boundary(nx=120, ny=80, max_iter=200, nsamples=300)
    want = golden["stage1_boundary_sample_seed11"]
    assert got.shape == want.shape == (300, 2)
    # the seeded weighted draw: the same candidates in the same order, weights equal to ~1e-13 -> the same picks
    assert np.array_equal(got, want)
    got = pt.sample_mandelbrot_boundary(nx=60, ny=40, max_iter=150, nsamples=10 ** 6)       # no subsampling: every candidate
    assert np.array_equal(got, golden["stage1_boundary_sample_all"])
    got = pt.mandelbrot_boundary_points(N=160, dist_thresh=0.01, max_iter=200)
    want = golden["vario_boundary_points_N160"]
    # the reference iterates numpy complex ARRAYS (FMA multiply): a few pixels may escape one step apart
    common = np.intersect1d(got, want).size
    assert common >= 0.99 * max(got.size, want.size) and abs(got.size - want.size) <= 0.01 * want.size


def test_grid_samplers_with_oracle_compute_cpu(oracle, golden, monkeypatch):
    """sample_mandelbrot_boundary (construct_stage1_clean.py:60-80) and mandelbrot_boundary_points
    (variograms_construct_mandelbrot.py:90-104): host logic around the distance-estimator grid, the device call replaced
    by the oracle."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import potentials as pt

    def grid(xs, ys, max_iter, bailout, eps, variant):
        d, e = oracle.distance_grid(xs, ys, max_iter, bailout, eps, variant)
        return d, e.astype(bool)
    monkeypatch.setattr(pt, "distance_grid", grid)
    _check_samplers(pt, golden)


@pytest.mark.gpu
def test_grid_samplers_on_the_device(gpu, golden):
    _check_samplers(gpu.potentials, golden)

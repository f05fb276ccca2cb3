import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the reference's own Python functions (oracle/gen_golden.py)."""
    with np.load(ROOT / "tests" / "golden" / "reference_vectors.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def shim():
    """The product library, built in-tree (nvcc cross-compiles without a GPU)."""
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import _shim, build
    build.build()
    _shim.load()
    return _shim


@pytest.fixture(scope="session")
def gpu(shim):
    """Product modules, requiring a Blackwell device; fails loudly (no skip) when absent."""
    if shim.device_count() < 1:
        pytest.fail("gpu-marked test run without a CUDA device: " + shim.last_error())
    shim.set_device(0)
    import inverse_eigenvalue_loci_mandelbrot_correspondence_b200 as pkg
    from inverse_eigenvalue_loci_mandelbrot_correspondence_b200 import (contour, device, escape, lucas, potentials,
                                                                        stencils)

    class NS:
        pass

    ns = NS()
    ns.pkg, ns.contour, ns.device, ns.escape, ns.lucas, ns.potentials, ns.stencils, ns.shim = (
        pkg, contour, device, escape, lucas, potentials, stencils, shim)
    return ns


def match_sorted_complex(a: np.ndarray, b: np.ndarray) -> float:
    """Max relative distance between two complex multisets after greedy nearest matching."""
    a = np.asarray(a, dtype=np.complex128).ravel()
    b = np.asarray(b, dtype=np.complex128).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    used = np.zeros(b.size, dtype=bool)
    worst = 0.0
    for v in a:
        dist = np.abs(b - v)
        dist[used] = np.inf
        k = int(np.argmin(dist))
        used[k] = True
        worst = max(worst, float(dist[k] / max(abs(v), 1e-300)))
    return worst

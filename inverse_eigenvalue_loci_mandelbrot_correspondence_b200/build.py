"""Build liblm_b200.so in-tree with nvcc for sm_100a.

The library is plain CUDA C++ behind a C ABI (include/lm_b200.h); it links the static CUDA
runtime and nothing else, so it loads with ctypes without torch.  nvcc cross-compiles on a
box without a GPU.  The built .so is git-ignored but travels to the GPU box with the tree.

    python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
BUILD_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "liblm_b200.so"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
          "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr"]

# Files whose arithmetic must reproduce the reference's unfused IEEE binary64 operation
# order are compiled with -fmad=false (on top of using __dmul_rn/__dadd_rn explicitly).
# The root solver and the log-potential are tolerance-matched and keep FMA.
PER_FILE = {
    "lm_roots.cu": [],
    "lm_logpot.cu": [],
}
DEFAULT_EXTRA = ["-fmad=false"]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _extra_defs() -> list[str]:
    """Extra -D flags for tuning sweeps, e.g. LM_NVCC_DEFS="-DLM_K1_FB=32 -DLM_K1_MIN_CTAS=4"."""
    return os.environ.get("LM_NVCC_DEFS", "").split()


def _digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(_extra_defs()).encode())
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "lm_b200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def is_current() -> bool:
    stamp = BUILD_DIR / "stamp"
    return LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link liblm_b200.so.  Returns the library path."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    BUILD_DIR.mkdir(exist_ok=True)
    srcs = _sources()
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")

    def compile_one(src: Path) -> Path:
        obj = BUILD_DIR / (src.stem + ".o")
        extra = PER_FILE.get(src.name, DEFAULT_EXTRA)
        cmd = [nvcc, *ARCH_FLAGS, *COMMON, *extra, *_extra_defs(), "-I", str(INCLUDE), "-I", str(CSRC),
               "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))

    link = [nvcc, *ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
            "-o", str(LIB_PATH), *map(str, objs)]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (BUILD_DIR / "stamp").write_text(_digest())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

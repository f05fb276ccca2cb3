"""ctypes binding of liblm_b200.so (C ABI declared in include/lm_b200.h).

There is no CPU fallback: if the library is missing or no Blackwell GPU is present, every
compute call raises.  Calls are serialised with a process-wide lock because the shim keeps
one context (device workspaces, work-queue counters) per process and is not re-entrant.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "liblm_b200.so"

LM_OK, LM_E_INVALID, LM_E_CUDA, LM_E_CAP, LM_E_NOMEM, LM_E_NODEV, LM_E_NOCONV, LM_E_OVERFLOW = (
    0, -1, -2, -3, -4, -5, -6, -7)

FIELD_NONE, FIELD_GREEN, FIELD_POW2_ALWAYS, FIELD_INV_K, FIELD_POW2_FIRST = 0, 1, 2, 3, 4
DE_SCALAR, DE_FIRST_ESCAPE, DE_FINAL_DZ, DE_FINAL_DZ_NUMPY = 0, 1, 2, 3
LOGPOT_SUM_SQRT, LOGPOT_NEG_PERTERM, LOGPOT_SUM_HYPOT, LOGPOT_LOG_INV = 0, 1, 2, 3
PAIR_W_NONE, PAIR_W_VALUE_SQDIFF, PAIR_W_DIST_SQ = 0, 1, 2

# every symbol include/lm_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "lm_abi_version", "lm_last_error", "lm_device_count", "lm_set_device", "lm_get_device_info",
    "lm_device_synchronize", "lm_release_workspace", "lm_host_alloc", "lm_host_free", "lm_dev_alloc",
    "lm_dev_free", "lm_memcpy_h2d", "lm_memcpy_d2h", "lm_memcpy_d2d", "lm_stream_synchronize",
    "lm_escape_grid_f64", "lm_escape_grid_f64_dev", "lm_shard_escape", "lm_escape_grid_f32", "lm_escape_grid_f32_dev", "lm_escape_points_f64",
    "lm_distance_grid_f64",
    "lm_contour_level", "lm_contour_level_dev", "lm_contour_classify_dev", "lm_contour_records_dev", "lm_contour_link", "lm_contour_link_dev", "lm_contour_fetch_last", "lm_boundary_sample", "lm_boundary_sample_potential",
    "lm_roots_batched", "lm_roots_batched_dev", "lm_cloud_compact_dev", "lm_cloud_append_dev", "lm_lucas_cloud_fields", "lm_lucas_cloud_fields_i8", "lm_escape_points_f64_dev",
    "lm_laplacian5_periodic", "lm_laplacian5_periodic_dev", "lm_smooth5_interior", "lm_smooth5_interior_dev",
    "lm_log_potential", "lm_log_potential_sums_dev", "lm_log_potential_finish_dev",
    "lm_nearest_match", "lm_weighted_log_sum", "lm_weighted_cauchy_sum", "lm_curvature_localpoly",
    "lm_pair_histogram", "lm_pair_select_sqdiff", "lm_pair_max_distance", "lm_alpha_shape_edges",
    "lm_histogram2d", "lm_mollified_histogram", "lm_gaussian_filter_nearest", "lm_sum_pairwise", "lm_density_compare", "lm_gi_flow",
    "lm_probe_fp64_peak", "lm_probe_fp64_latency", "lm_probe_k1_loop", "lm_probe_hbm_copy",
)


class LmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"liblm_b200 error {code}: {msg}")
        self.code = code


class DeviceInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_count", C.c_int32), ("clock_khz", C.c_int32), ("l2_bytes", C.c_int32),
                ("total_mem_bytes", C.c_uint64), ("name", C.c_char * 128)]


class Stats(C.Structure):
    _fields_ = [("work_units", C.c_uint64), ("items", C.c_uint64), ("kernel_ms", C.c_float),
                ("launches", C.c_int32)]

    def as_dict(self) -> dict:
        return {"work_units": int(self.work_units), "items": int(self.items),
                "kernel_ms": float(self.kernel_ms), "launches": int(self.launches)}


class CloudStats(C.Structure):
    _fields_ = [("n_roots", C.c_uint64), ("n_points", C.c_uint64), ("potential_work", C.c_uint64), ("pairs", C.c_uint64),
                ("roots_ms", C.c_float), ("compact_ms", C.c_float), ("potential_ms", C.c_float), ("logpot_ms", C.c_float),
                ("stencil_ms", C.c_float), ("launches", C.c_int32)]

    def as_dict(self) -> dict:
        return {k: (int(getattr(self, k)) if "ms" not in k else float(getattr(self, k))) for k, _ in self._fields_}


_lock = threading.RLock()
_lib: C.CDLL | None = None

_vp, _i32, _i64, _f64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t
_pStats = C.POINTER(Stats)
_pi64 = C.POINTER(C.c_int64)

_SIGNATURES = {
    "lm_abi_version": (_i32, []),
    "lm_last_error": (C.c_char_p, []),
    "lm_device_count": (_i32, []),
    "lm_set_device": (_i32, [_i32]),
    "lm_get_device_info": (_i32, [C.POINTER(DeviceInfo)]),
    "lm_device_synchronize": (_i32, []),
    "lm_release_workspace": (_i32, []),
    "lm_host_alloc": (_vp, [_sz]),
    "lm_host_free": (_i32, [_vp]),
    "lm_dev_alloc": (_vp, [_sz]),
    "lm_dev_free": (_i32, [_vp]),
    "lm_memcpy_h2d": (_i32, [_vp, _vp, _sz, _vp]),
    "lm_memcpy_d2h": (_i32, [_vp, _vp, _sz, _vp]),
    "lm_memcpy_d2d": (_i32, [_vp, _vp, _sz, _vp]),
    "lm_stream_synchronize": (_i32, [_vp]),
    "lm_shard_escape": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _pStats]),
    "lm_escape_grid_f64": (_i32, [_vp, _i64, _vp, _i64, _i32, _f64, _i32, _vp, _vp, _vp, _pStats]),
    "lm_escape_grid_f64_dev": (_i32, [_vp, _i64, _vp, _i64, _i32, _f64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lm_escape_grid_f32": (_i32, [_vp, _i64, _vp, _i64, _i32, _f64, _vp, _pStats]),
    "lm_escape_grid_f32_dev": (_i32, [_vp, _i64, _vp, _i64, _i32, _f64, _vp, _vp, _vp]),
    "lm_escape_points_f64": (_i32, [_vp, _vp, _i64, _i32, _f64, _vp, _vp, _vp, _vp, _pStats]),
    "lm_distance_grid_f64": (_i32, [_vp, _i64, _vp, _i64, _i32, _f64, _f64, _i32, _vp, _vp, _pStats]),
    "lm_contour_level": (_i32, [_vp, _vp, _i64, _vp, _i64, _f64, _vp, _i64, _pi64, _vp, _i64, _pi64, _pStats]),
    "lm_contour_level_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _f64, _vp, _i64, _pi64, _vp, _i64, _pi64, _pStats]),
    "lm_contour_fetch_last": (_i32, [_vp, _i64, _pi64, _vp, _i64, _pi64]),
    "lm_boundary_sample": (_i32, [_vp, _i64, _vp, _i64, _i32, _f64, _vp, _vp, _vp, _i64, _pi64, _vp, _i64, _pi64, _pStats]),
    "lm_boundary_sample_potential": (_i32, [_vp, _i64, _vp, _i64, _i32, _f64, _vp, _vp, _vp, _vp, _i64, _pi64, _vp, _i64, _pi64, _pStats]),
    "lm_contour_classify_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _i64, _f64, _vp, _i64, _pi64, _vp]),
    "lm_contour_records_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _i64, _f64, _vp, _i64, _pi64, _vp]),
    "lm_contour_link": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _f64, _vp, _i64, _pi64, _vp, _i64, _pi64]),
    "lm_contour_link_dev": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _f64, _vp, _i64, _pi64, _vp, _i64, _pi64, _vp]),
    "lm_roots_batched": (_i32, [_vp, _vp, _i64, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _pStats]),
    "lm_roots_batched_dev": (_i32, [_vp, _vp, _i64, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lm_cloud_compact_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp]),
    "lm_cloud_append_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp]),
    "lm_lucas_cloud_fields": (_i32, [_vp, _vp, _i64, _i32, _f64, _vp, _vp, _i64, _pi64, _i32, _f64, _vp, _vp,
                                     _vp, _i64, _vp, _i64, _f64, _i32, _f64, _vp, _vp, C.POINTER(CloudStats)]),
    "lm_lucas_cloud_fields_i8": (_i32, [_vp, _vp, _i64, _i32, _f64, _vp, _vp, _i64, _pi64, _i32, _f64, _vp, _vp,
                                     _vp, _i64, _vp, _i64, _f64, _i32, _f64, _vp, _vp, C.POINTER(CloudStats)]),
    "lm_escape_points_f64_dev": (_i32, [_vp, _vp, _i64, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lm_log_potential_sums_dev": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _f64, _i32, _vp, _vp]),
    "lm_log_potential_finish_dev": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp]),
    "lm_laplacian5_periodic": (_i32, [_vp, _i64, _i64, _f64, _vp, _pStats]),
    "lm_laplacian5_periodic_dev": (_i32, [_vp, _i64, _i64, _f64, _vp, _vp]),
    "lm_smooth5_interior": (_i32, [_vp, _i64, _i64, _vp, _pStats]),
    "lm_smooth5_interior_dev": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "lm_log_potential": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _f64, _i32, _vp, _pStats]),
    "lm_nearest_match": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _pStats]),
    "lm_weighted_log_sum": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _f64, _vp, _pStats]),
    "lm_weighted_cauchy_sum": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _f64, _vp, _vp, _pStats]),
    "lm_curvature_localpoly": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pStats]),
    "lm_pair_histogram": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _pStats]),
    "lm_pair_max_distance": (_i32, [_vp, _vp, _i64, C.POINTER(C.c_double), _pStats]),
    "lm_pair_select_sqdiff": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _f64, _f64, _i32, _vp, _i64, _pi64, _pStats]),
    "lm_histogram2d": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _i32, _vp, _pStats]),
    "lm_mollified_histogram": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _i32, _f64, _vp, _i32, _vp, _pStats]),
    "lm_gaussian_filter_nearest": (_i32, [_vp, _i64, _i64, _vp, _i32, _vp, _pStats]),
    "lm_sum_pairwise": (_i32, [_vp, _i64, C.POINTER(C.c_double), _pStats]),
    "lm_density_compare": (_i32, [_vp, _vp, _i64, _f64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), _pStats]),
    "lm_gi_flow": (_i32, [_vp, _vp, _i64, _f64, _f64, _i32, _i32, _f64, _i32, _vp, C.POINTER(C.c_int32),
                          C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _pStats]),
    "lm_alpha_shape_edges": (_i32, [_vp, _vp, _i64, _vp, _i64, _f64, _vp, _vp, _vp, _i64, _pi64, _pStats]),
    "lm_probe_fp64_peak": (_i32, [_i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lm_probe_fp64_latency": (_i32, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lm_probe_k1_loop": (_i32, [_i32, _i32, _i32, C.POINTER(C.c_double)]),
    "lm_probe_hbm_copy": (_i32, [_sz, _i32, C.POINTER(C.c_double)]),
}


def load() -> C.CDLL:
    """dlopen liblm_b200.so (no device needed) and attach signatures.  Raises if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with "
                    "`python -m inverse_eigenvalue_loci_mandelbrot_correspondence_b200.build` "
                    "(there is no CPU fallback)")
            lib = C.CDLL(str(LIB_PATH))
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name, None)
                if fn is None:          # reported by missing_exports(); calling it raises
                    continue
                fn.restype = res
                fn.argtypes = args
            _lib = lib
        return _lib


def missing_exports() -> list[str]:
    """Symbols declared in include/lm_b200.h that the loaded library does not export."""
    lib = load()
    return [n for n in EXPORTS if getattr(lib, n, None) is None]


def last_error() -> str:
    return load().lm_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == LM_OK:
        return
    msg = last_error()
    if rc == LM_E_OVERFLOW:
        raise OverflowError(msg)
    if rc == LM_E_NOMEM:
        raise MemoryError(msg)
    if rc == LM_E_INVALID:
        raise ValueError(msg)
    raise LmError(rc, msg)


def call(name: str, *args) -> None:
    """Call an int32-status entry point under the context lock; raise on failure."""
    lib = load()
    with _lock:
        rc = getattr(lib, name)(*args)
    check(rc)


def ptr(a: np.ndarray | None):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return int(load().lm_device_count())


def set_device(dev: int) -> None:
    call("lm_set_device", int(dev))


def device_info() -> dict:
    info = DeviceInfo()
    call("lm_get_device_info", C.byref(info))
    return {"device": info.device, "cc": (info.cc_major, info.cc_minor), "sm_count": info.sm_count,
            "clock_khz": info.clock_khz, "l2_bytes": info.l2_bytes,
            "total_mem_bytes": int(info.total_mem_bytes), "name": info.name.decode()}


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory (freed when the array is garbage collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    nbytes = max(n * dtype.itemsize, 16)
    lib = load()
    p = lib.lm_host_alloc(nbytes)
    if not p:
        raise MemoryError(last_error())

    class _Owner:
        def __init__(self, addr):
            self.addr = addr

        def __del__(self):
            try:
                lib.lm_host_free(self.addr)
            except Exception:
                pass

    owner = _Owner(p)
    buf = (C.c_char * nbytes).from_address(p)
    buf._owner = owner   # keep the allocation alive as long as the buffer object
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    return arr

"""Device-resident grid: keeps xs, ys and the K1 outputs in HBM between K1 -> K2 / K4 so that
only what the caller asks for crosses PCIe (SURVEY.md section 7 "host I/O dominates at 32768^2").

Device memory comes from the shim's allocator (cudaMalloc); every kernel is launched through the
C ABI's *_dev entry points on the legacy default stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _shim
from . import contour as _contour
from ._shim import FIELD_NONE


class DeviceBuffer:
    """A cudaMalloc'ed buffer owned by this object."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        lib = _shim.load()
        self.ptr = lib.lm_dev_alloc(max(self.nbytes, 16))
        if not self.ptr:
            raise MemoryError(_shim.last_error())

    def free(self) -> None:
        if getattr(self, "ptr", None):
            _shim.load().lm_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def upload(self, host: np.ndarray, stream=None) -> None:
        host = np.ascontiguousarray(host)
        assert host.nbytes <= self.nbytes
        _shim.call("lm_memcpy_h2d", C.c_void_p(self.ptr), _shim.ptr(host), host.nbytes, stream)

    def download(self, host: np.ndarray, nbytes: int | None = None, offset: int = 0, stream=None) -> None:
        n = host.nbytes if nbytes is None else int(nbytes)
        _shim.call("lm_memcpy_d2h", _shim.ptr(host), C.c_void_p(self.ptr + offset), n, stream)


class DeviceGrid:
    """xs[nx], ys[ny] on the device plus lazily allocated dwell / field outputs."""

    def __init__(self, xs, ys):
        self.xs = np.ascontiguousarray(xs, dtype=np.float64).ravel()
        self.ys = np.ascontiguousarray(ys, dtype=np.float64).ravel()
        self.nx, self.ny = self.xs.size, self.ys.size
        self.d_xs = DeviceBuffer(self.xs.nbytes); self.d_xs.upload(self.xs)
        self.d_ys = DeviceBuffer(self.ys.nbytes); self.d_ys.upload(self.ys)
        self.d_dwell: DeviceBuffer | None = None
        self.d_field: DeviceBuffer | None = None
        self.d_work = DeviceBuffer(64)
        self.max_iter: int | None = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self) -> None:
        for b in (self.d_xs, self.d_ys, self.d_dwell, self.d_field, self.d_work):
            if b is not None:
                b.free()

    def escape(self, max_iter: int, bailout: float = 2.0, field_mode: int = FIELD_NONE, stream=None) -> None:
        """Enqueue K1 over the whole grid (results stay in HBM)."""
        npx = self.nx * self.ny
        if self.d_dwell is None:
            self.d_dwell = DeviceBuffer(npx * 4)
        if field_mode != FIELD_NONE and self.d_field is None:
            self.d_field = DeviceBuffer(npx * 8)
        _shim.call("lm_escape_grid_f64_dev", C.c_void_p(self.d_xs.ptr), self.nx, C.c_void_p(self.d_ys.ptr), self.ny,
                   int(max_iter), float(bailout), int(field_mode), C.c_void_p(self.d_dwell.ptr), None,
                   C.c_void_p(self.d_field.ptr) if field_mode != FIELD_NONE else None,
                   C.c_void_p(self.d_work.ptr), stream)
        self.max_iter = int(max_iter)

    def work_units(self) -> int:
        """Exact pixel-iteration count of the last escape() (synchronises)."""
        out = np.zeros(1, dtype=np.uint64)
        self.d_work.download(out)
        _shim.call("lm_stream_synchronize", None)
        return int(out[0])

    def dwell(self, pinned: bool = False) -> np.ndarray:
        out = (_shim.pinned_empty if pinned else np.empty)((self.ny, self.nx), np.int32)
        self.d_dwell.download(out)
        _shim.call("lm_stream_synchronize", None)
        return out

    def field(self, pinned: bool = False) -> np.ndarray:
        out = (_shim.pinned_empty if pinned else np.empty)((self.ny, self.nx), np.float64)
        self.d_field.download(out)
        _shim.call("lm_stream_synchronize", None)
        return out

    def contour(self, level: float):
        """K2 on the resident dwell grid -> list of (N,2) polylines."""
        return _contour.contour_lines_dev(self.d_dwell.ptr, self.xs, self.ys, level)

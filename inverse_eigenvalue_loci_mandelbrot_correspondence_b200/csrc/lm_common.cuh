// lm_common.cuh -- shared plumbing of liblm_b200.so (error reporting, device workspace
// cache, launch geometry).  Internal; the public surface is include/lm_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <vector>

#include "lm_b200.h"

namespace lm {

// ---- error reporting (thread-local message, status code returned to the caller) ----
int32_t fail(int32_t code, const char* fmt, ...);
void    clear_error();

#define LM_CUDA_TRY(expr)                                                              \
    do {                                                                               \
        cudaError_t lm_e__ = (expr);                                                   \
        if (lm_e__ != cudaSuccess)                                                     \
            return lm::fail(lm_e__ == cudaErrorMemoryAllocation ? LM_E_NOMEM : LM_E_CUDA, \
                            "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                            cudaGetErrorString(lm_e__));                               \
    } while (0)

#define LM_REQUIRE(cond, ...)                                                          \
    do {                                                                               \
        if (!(cond)) return lm::fail(LM_E_INVALID, __VA_ARGS__);                       \
    } while (0)

// Makes sure a CUDA device of compute capability >= 10.0 is current; LM_E_NODEV otherwise.
int32_t require_device();
int     sm_count();

// ---- device workspace cache ----------------------------------------------------------
// Host entry points stage caller buffers through cached device allocations so repeated
// calls (bench steps, tiled scripts) do not pay cudaMalloc/cudaFree each time.  Slots are
// per purpose; a slot grows monotonically and is freed by lm_release_workspace().
enum WsSlot {
    WS_XS = 0, WS_YS, WS_OUT_I32, WS_OUT_F64, WS_FIELD, WS_COUNTERS, WS_IN_A, WS_IN_B,
    WS_IN_C, WS_OUT_A, WS_OUT_B, WS_OUT_C, WS_OUT_D, WS_RECORDS, WS_SCRATCH,
    WS_K1_WORK, WS_K2_MASK, WS_K2_COUNT, WS_K2_OFFSET, WS_K2_XS, WS_K2_YS,
    WS_K1_SURVIVORS, WS_LOGPOT_PART, WS_ROOTS_PLAN, WS_ROOTS_INDEX, WS_CLOUD_SCAN, WS_CLOUD_A, WS_CLOUD_B, WS_CLOUD_C, WS_CLOUD_D,
    WS_LINK, WS_LINK_OUT, WS_NSLOTS
};
int32_t ws_get(WsSlot slot, size_t bytes, void** out);
void    ws_release_all();
// Translation units that cache page-locked staging buffers / streams of their own register a hook that
// lm_release_workspace() runs before the device workspaces are freed.
void    register_release_hook(void (*fn)());

// RAII-free helper for event timing on a stream.
struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    int32_t begin(cudaStream_t s);
    int32_t end(cudaStream_t s, float* ms);   // synchronises on the end event
    ~Timer();
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- cross-file internals ---------------------------------------------------------------
// An in-flight host-buffer K1 job (lm_escape.cu): row chunks are computed on s_compute while
// s_copy returns finished chunks to the caller's host buffers.
struct GridHostJob {
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
    void* dwork = nullptr;            // device: [0] work counter, [1] overflow flag
    int32_t* dwell_dev = nullptr;     // device int32 dwell grid (when produced)
    double* field_dev = nullptr;      // device field (when produced)
    int launches = 0;
    size_t npx = 0;
    void release_events();
};
int32_t grid_host_begin(const double* xs, int64_t nx, const double* ys, int64_t ny, int32_t max_iter, double bailout,
                        int32_t field_mode, int32_t* dwell_i32, double* dwell_f64, double* field,
                        bool need_dev_dwell, int64_t extra_rows, GridHostJob* job, const double* row_cost = nullptr);
int32_t grid_host_finish(GridHostJob* job, lm_stats* stats);
// K2 on a device-resident dwell grid -> polylines in host buffers (lm_contour.cu + lm_contour_link.cu); enqueues
// on s and synchronises s for the record count and the line sizes.
int32_t contour_device_to_host(const int32_t* dwell_dev, const double* xs_host, int64_t nx, const double* ys_host,
                               int64_t ny, double level, double* verts, int64_t cap_verts, int64_t* n_verts,
                               int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                               float* kernel_ms, int* launches, cudaStream_t s);

// raster-ordered crossing records on the device -> mpl2014-ordered polylines, kept on the device (lm_contour_link.cu)
int32_t contour_link_device(const long long* rec_dev, long long n, const double* xs_host, long long nx,
                            const double* ys_host, long long ny, double level, long long* nv, long long* nl,
                            float* kernel_ms, cudaStream_t s);
// the lines of the last contour_link_device on this device -> caller buffers (LM_E_CAP with the sizes if too small)
int32_t contour_export(double* verts, long long cap_verts, long long* n_verts, long long* line_offsets,
                       long long cap_lines, long long* n_lines, cudaStream_t s, const char* who);

}  // namespace lm

// lm_common.cuh -- shared plumbing of liblm_b200.so (error reporting, device workspace
// cache, launch geometry).  Internal; the public surface is include/lm_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

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
    WS_IN_C, WS_OUT_A, WS_OUT_B, WS_OUT_C, WS_OUT_D, WS_RECORDS, WS_SCRATCH, WS_NSLOTS
};
int32_t ws_get(WsSlot slot, size_t bytes, void** out);
void    ws_release_all();

// RAII-free helper for event timing on a stream.
struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    int32_t begin(cudaStream_t s);
    int32_t end(cudaStream_t s, float* ms);   // synchronises on the end event
    ~Timer();
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace lm

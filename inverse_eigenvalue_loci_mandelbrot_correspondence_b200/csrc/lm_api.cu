// lm_api.cu -- context, error and memory entry points of the C-ABI (include/lm_b200.h).
#include "lm_common.cuh"

#include <string.h>

namespace lm {

static thread_local char g_err[1024] = "";

int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { g_err[0] = 0; }

static int g_checked_device = -1;   // ordinal for which require_device() already passed
static int g_sm_count = 0;

int32_t require_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LM_E_NODEV, "no usable CUDA device: %s", cudaGetErrorString(e));
    }
    if (dev == g_checked_device) return LM_OK;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LM_E_NODEV, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    }
    if (p.major < 10)
        return fail(LM_E_NODEV,
                    "device %d (%s, sm_%d%d) is not a Blackwell part; liblm_b200 carries sm_100a code only",
                    dev, p.name, p.major, p.minor);
    g_checked_device = dev;
    g_sm_count = p.multiProcessorCount;
    return LM_OK;
}
int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

// ---- workspace cache -----------------------------------------------------------------
struct WsEntry { void* p = nullptr; size_t bytes = 0; int dev = -1; };
static WsEntry g_ws[WS_NSLOTS];

int32_t ws_get(WsSlot slot, size_t bytes, void** out) {
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    WsEntry& w = g_ws[slot];
    if (bytes == 0) bytes = 16;
    if (w.p && (w.bytes < bytes || w.dev != dev)) {
        int cur = dev;
        if (w.dev != dev) cudaSetDevice(w.dev);
        cudaFree(w.p);
        if (w.dev != cur) cudaSetDevice(cur);
        w.p = nullptr; w.bytes = 0;
    }
    if (!w.p) {
        cudaError_t e = cudaMalloc(&w.p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            w.p = nullptr;
            return fail(LM_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        }
        w.bytes = bytes; w.dev = dev;
    }
    *out = w.p;
    return LM_OK;
}

static void (*g_hooks[16])() = {};
static int g_nhooks = 0;
void register_release_hook(void (*fn)()) {
    for (int i = 0; i < g_nhooks; ++i) if (g_hooks[i] == fn) return;
    if (g_nhooks < 16) g_hooks[g_nhooks++] = fn;
}

void ws_release_all() {
    for (int i = 0; i < g_nhooks; ++i) g_hooks[i]();
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
    for (int i = 0; i < WS_NSLOTS; ++i) {
        if (g_ws[i].p) {
            if (g_ws[i].dev != cur) cudaSetDevice(g_ws[i].dev);
            cudaFree(g_ws[i].p);
            if (g_ws[i].dev != cur) cudaSetDevice(cur);
            g_ws[i] = WsEntry();
        }
    }
}

int32_t Timer::begin(cudaStream_t s) {
    LM_CUDA_TRY(cudaEventCreate(&a));
    LM_CUDA_TRY(cudaEventCreate(&b));
    LM_CUDA_TRY(cudaEventRecord(a, s));
    return LM_OK;
}
int32_t Timer::end(cudaStream_t s, float* ms) {
    LM_CUDA_TRY(cudaEventRecord(b, s));
    LM_CUDA_TRY(cudaEventSynchronize(b));
    LM_CUDA_TRY(cudaEventElapsedTime(ms, a, b));
    return LM_OK;
}
Timer::~Timer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
}

}  // namespace lm

using namespace lm;

extern "C" {

int32_t lm_abi_version(void) { return LM_ABI_VERSION; }

const char* lm_last_error(void) { return g_err; }

int32_t lm_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(LM_E_NODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return 0;
    }
    return n;
}

int32_t lm_set_device(int32_t device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(LM_E_NODEV, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    }
    return require_device();
}

int32_t lm_get_device_info(lm_device_info* out) {
    LM_REQUIRE(out != nullptr, "lm_get_device_info: out is NULL");
    int32_t rc = require_device();
    if (rc != LM_OK) return rc;
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    LM_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    memset(out, 0, sizeof(*out));
    out->device = dev;
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->sm_count = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    out->clock_khz = khz;
    out->l2_bytes = p.l2CacheSize;
    out->total_mem_bytes = p.totalGlobalMem;
    strncpy(out->name, p.name, sizeof(out->name) - 1);
    return LM_OK;
}

int32_t lm_device_synchronize(void) {
    LM_CUDA_TRY(cudaDeviceSynchronize());
    return LM_OK;
}

int32_t lm_release_workspace(void) {
    ws_release_all();
    return LM_OK;
}

void* lm_host_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(LM_E_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
int32_t lm_host_free(void* p) {
    if (p) LM_CUDA_TRY(cudaFreeHost(p));
    return LM_OK;
}
void* lm_dev_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(LM_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
int32_t lm_dev_free(void* p) {
    if (p) LM_CUDA_TRY(cudaFree(p));
    return LM_OK;
}
int32_t lm_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream) {
    LM_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return LM_OK;
}
int32_t lm_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream) {
    LM_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return LM_OK;
}
int32_t lm_memcpy_d2d(void* dst_dev, const void* src_dev, size_t bytes, void* stream) {
    LM_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return LM_OK;
}
int32_t lm_stream_synchronize(void* stream) {
    LM_CUDA_TRY(cudaStreamSynchronize(as_stream(stream)));
    return LM_OK;
}

}  // extern "C"

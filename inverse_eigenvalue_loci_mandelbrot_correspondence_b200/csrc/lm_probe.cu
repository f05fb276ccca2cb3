// lm_probe.cu -- measurement probes for the roofline denominators.
//
// MEASURED_PEAKS.json (driver-written) holds HBM copy bandwidth and bf16 GEMM throughput but
// no FP64 figure, and K1/K3 are FP64-pipe bound.  lm_probe_fp64_peak runs a dependent-free
// DFMA loop on every SM (8 independent chains per thread, enough warps to cover the pipe
// latency) and reports TFLOP/s at 2 flops per DFMA; the second figure is the same loop with
// the unfused DMUL/DADD mix K1 uses, in 10^12 FP64 instructions per second.
#include "lm_common.cuh"

namespace {

constexpr int CHAINS = 8;

template <bool FMA_MIX>
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double seed, double* sink) {
    double v[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) v[k] = seed + 1e-3 * (threadIdx.x + k);
    const double m = 1.0 - 1e-9, c = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) {
                if (FMA_MIX) {
                    v[k] = __fma_rn(v[k], m, c);
                } else {
                    // alternate DMUL / DADD like the unfused recurrence
                    v[k] = (k & 1) ? __dmul_rn(v[k], m) : __dadd_rn(v[k], c);
                }
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += v[k];
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains live
}

// one dependent chain per thread, one warp per SMSP: cycles per dependent FP64 instruction
template <int OP>
__global__ void fp64_latency_kernel(int iters, double seed, double* sink, long long* cycles) {
    double v = seed + 1e-3 * threadIdx.x;
    const double m = 1.0 - 1e-9, c = 1e-12;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            if (OP == 0) v = __fma_rn(v, m, c);
            else if (OP == 1) v = __dadd_rn(v, c);
            else v = __dmul_rn(v, m);
        }
    }
    const long long t1 = clock64();
    if (v == 123.456) sink[0] = v;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// the bare K1 recurrence (6 FP64 instructions per iteration, no escape test) on interior
// points: the ceiling for the blind path of lm_escape_kernel at a given occupancy.
// variant 0: recurrence only; 1: + hi-word max per iteration and a vote every 16 iterations.
template <int VARIANT>
__global__ void k1_loop_kernel(int blocks16, double* sink) {
    const double cr = -0.1 + 1e-4 * (threadIdx.x & 31), ci = 0.05 + 1e-5 * (threadIdx.x >> 5);
    double zr = 0.0, zi = 0.0, a = 0.0, b = 0.0;
    unsigned flagged = 0;
    for (int blk = 0; blk < blocks16; ++blk) {
        unsigned acc = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double p = __dmul_rn(zr, zi);
            const double t = __dsub_rn(a, b);
            zr = __dadd_rn(t, cr);
            zi = __fma_rn(2.0, p, ci);
            a = __dmul_rn(zr, zr);
            b = __dmul_rn(zi, zi);
            if (VARIANT == 1) acc = max(acc, max(static_cast<unsigned>(__double2hiint(a)), static_cast<unsigned>(__double2hiint(b))));
        }
        if (VARIANT == 1 && __any_sync(0xffffffffu, acc >= 0x40000000u)) flagged++;
    }
    if (zr == 123.456 || flagged == 77777u) sink[0] = zr + zi;
}

__global__ void __launch_bounds__(256) copy_kernel(const double4* __restrict__ src, double4* __restrict__ dst, size_t n4) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride)
        dst[i] = src[i];
}

}  // namespace

extern "C" {

int32_t lm_probe_fp64_peak(int32_t iters, double* dfma_tflops, double* dmul_dadd_tinstr) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(iters > 0, "lm_probe_fp64_peak: iters must be positive");
    void* sink = nullptr;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &sink)) != LM_OK) return rc;
    const int grid = lm::sm_count() * 8;
    const double instr = static_cast<double>(grid) * 256.0 * CHAINS * 4.0 * iters;
    for (int variant = 0; variant < 2; ++variant) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {   // first rep is the warm-up
            lm::Timer tm;
            if ((rc = tm.begin(nullptr)) != LM_OK) return rc;
            if (variant == 0)
                fp64_probe_kernel<true><<<grid, 256>>>(iters, 1.0, static_cast<double*>(sink));
            else
                fp64_probe_kernel<false><<<grid, 256>>>(iters, 1.0, static_cast<double*>(sink));
            LM_CUDA_TRY(cudaGetLastError());
            float ms = 0.f;
            if ((rc = tm.end(nullptr, &ms)) != LM_OK) return rc;
            if (rep > 0 && ms < best) best = ms;
        }
        const double per_s = instr / (best * 1e-3);
        if (variant == 0 && dfma_tflops) *dfma_tflops = 2.0 * per_s / 1e12;
        if (variant == 1 && dmul_dadd_tinstr) *dmul_dadd_tinstr = per_s / 1e12;
    }
    return LM_OK;
}

int32_t lm_probe_fp64_latency(double* dfma_cycles, double* dadd_cycles, double* dmul_cycles) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    void* buf = nullptr;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &buf)) != LM_OK) return rc;
    double* sink = static_cast<double*>(buf);
    long long* cyc = reinterpret_cast<long long*>(buf) + 4;
    const int iters = 512;
    double out[3] = {0, 0, 0};
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) fp64_latency_kernel<0><<<1, 32>>>(iters, 1.0, sink, cyc);
            else if (op == 1) fp64_latency_kernel<1><<<1, 32>>>(iters, 1.0, sink, cyc);
            else fp64_latency_kernel<2><<<1, 32>>>(iters, 1.0, sink, cyc);
            LM_CUDA_TRY(cudaGetLastError());
            long long h = 0;
            LM_CUDA_TRY(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
            out[op] = static_cast<double>(h) / (iters * 32.0);
        }
    }
    if (dfma_cycles) *dfma_cycles = out[0];
    if (dadd_cycles) *dadd_cycles = out[1];
    if (dmul_cycles) *dmul_cycles = out[2];
    return LM_OK;
}

int32_t lm_probe_k1_loop(int32_t variant, int32_t warps_per_sm, int32_t blocks16, double* gpixel_iters_per_s) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(warps_per_sm >= 4 && warps_per_sm <= 64 && warps_per_sm % 4 == 0 && blocks16 > 0 && gpixel_iters_per_s,
               "lm_probe_k1_loop: bad arguments");
    void* sink = nullptr;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &sink)) != LM_OK) return rc;
    // one CTA per SM with warps_per_sm warps (<= 32 per CTA, else two CTAs)
    const int ctas_per_sm = warps_per_sm > 32 ? 2 : 1;
    const int threads = warps_per_sm / ctas_per_sm * 32;
    const int grid = lm::sm_count() * ctas_per_sm;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        lm::Timer tm;
        if ((rc = tm.begin(nullptr)) != LM_OK) return rc;
        if (variant == 0) k1_loop_kernel<0><<<grid, threads>>>(blocks16, static_cast<double*>(sink));
        else k1_loop_kernel<1><<<grid, threads>>>(blocks16, static_cast<double*>(sink));
        LM_CUDA_TRY(cudaGetLastError());
        float ms = 0.f;
        if ((rc = tm.end(nullptr, &ms)) != LM_OK) return rc;
        if (rep > 0 && ms < best) best = ms;
    }
    *gpixel_iters_per_s = static_cast<double>(grid) * threads * 16.0 * blocks16 / (best * 1e-3) / 1e9;
    return LM_OK;
}

int32_t lm_probe_hbm_copy(size_t bytes, int32_t reps, double* gbs) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(bytes >= 1024 && reps > 0 && gbs, "lm_probe_hbm_copy: bad arguments");
    const size_t n4 = bytes / sizeof(double4);
    void *a = nullptr, *b = nullptr;
    LM_CUDA_TRY(cudaMalloc(&a, n4 * sizeof(double4)));
    if (cudaMalloc(&b, n4 * sizeof(double4)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(a);
        return lm::fail(LM_E_NOMEM, "lm_probe_hbm_copy: cudaMalloc failed");
    }
    cudaMemset(a, 1, n4 * sizeof(double4));
    float best = 1e30f;
    const int grid = lm::sm_count() * 8;
    for (int r = 0; r < reps + 1; ++r) {
        lm::Timer tm;
        if (tm.begin(nullptr) != LM_OK) break;
        copy_kernel<<<grid, 256>>>(static_cast<const double4*>(a), static_cast<double4*>(b), n4);
        float ms = 0.f;
        if (tm.end(nullptr, &ms) != LM_OK) break;
        if (r > 0 && ms < best) best = ms;
    }
    cudaFree(a);
    cudaFree(b);
    LM_CUDA_TRY(cudaGetLastError());
    *gbs = 2.0 * static_cast<double>(n4 * sizeof(double4)) / (best * 1e-3) / 1e9;
    return LM_OK;
}

}  // extern "C"

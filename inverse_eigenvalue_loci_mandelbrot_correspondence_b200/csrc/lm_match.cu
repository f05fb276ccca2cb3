// lm_match.cu -- nearest-neighbour matching of two planar point sets (FP64 pipe, O(n*m)).
//
// The tracker's module matches every Construct point to a Mandelbrot sample with
//   M = cdist(X, Y); M /= M.mean(); K = exp(-M / sinkhorn_eps); match = argmax(K, axis=1)
//   entropic_ot_alignment, tci_construct_mandelbrot_v002_fixed.py:62-71
// exp(-M/const) is strictly decreasing in M, so argmax K is the FIRST index of the smallest Euclidean
// distance (the only deviation: two distances so close that their exponentials round to the same double
// would tie there and not here).  With n = m up to 25 000 the reference spends seconds in the 5 GB distance
// matrix; here a thread owns one X point and streams Y through shared memory.
// Distances are sqrt(dx*dx + dy*dy) with unfused operations, as cdist computes them.
#include "lm_common.cuh"

namespace {

constexpr int NM_THREADS = 128;
constexpr int NM_CHUNK = 1024;

__global__ void __launch_bounds__(NM_THREADS) nearest_kernel(const double* __restrict__ xr, const double* __restrict__ xi, long long n,
                                                             const double* __restrict__ yr, const double* __restrict__ yi, long long m,
                                                             long long* __restrict__ idx, double* __restrict__ dist) {
    __shared__ double2 sy[NM_CHUNK];
    const long long k = static_cast<long long>(blockIdx.x) * NM_THREADS + threadIdx.x;
    const bool live = k < n;
    const double px = live ? xr[k] : 0.0, py = live ? xi[k] : 0.0;
    double best = INFINITY;
    long long arg = 0;
    bool any = false;
    for (long long base = 0; base < m; base += NM_CHUNK) {
        const int c = static_cast<int>(m - base < NM_CHUNK ? m - base : NM_CHUNK);
        __syncthreads();
        for (int t = threadIdx.x; t < c; t += NM_THREADS) sy[t] = make_double2(yr[base + t], yi[base + t]);
        __syncthreads();
        if (!live) continue;
#pragma unroll 4
        for (int t = 0; t < c; ++t) {
            const double dx = __dsub_rn(px, sy[t].x), dy = __dsub_rn(py, sy[t].y);
            const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            // first index of the minimum; a NaN distance never wins unless everything is NaN (argmax of NaNs -> 0)
            if (d < best || (!any && d == d)) { best = d; arg = base + t; any = true; }
        }
    }
    if (live) {
        idx[k] = arg;
        if (dist) dist[k] = any ? best : nan("");
    }
}

}  // namespace

extern "C" {

int32_t lm_nearest_match(const double* x_re, const double* x_im, int64_t n,
                         const double* y_re, const double* y_im, int64_t m,
                         int64_t* index, double* distance, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0 && m >= 0, "lm_nearest_match: negative size");
    LM_REQUIRE(n == 0 || (x_re && x_im && index), "lm_nearest_match: NULL buffer");
    LM_REQUIRE(n == 0 || m > 0, "lm_nearest_match: nothing to match against (m = 0)");
    LM_REQUIRE(m == 0 || (y_re && y_im), "lm_nearest_match: NULL buffer");
    if (stats) *stats = lm_stats{};
    if (n == 0) return LM_OK;
    cudaStream_t s = nullptr;
    void *dxr, *dxi, *dyr, *dyi, *didx, *ddist;
    const size_t nb = static_cast<size_t>(n) * sizeof(double), mb = static_cast<size_t>(m) * sizeof(double);
    if ((rc = lm::ws_get(lm::WS_IN_A, nb, &dxr)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, nb, &dxi)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_C, mb, &dyr)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, mb, &dyi)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, static_cast<size_t>(n) * sizeof(long long), &didx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, nb, &ddist)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dxr, x_re, nb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dxi, x_im, nb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dyr, y_re, mb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dyi, y_im, mb, cudaMemcpyHostToDevice, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    nearest_kernel<<<static_cast<unsigned>((n + NM_THREADS - 1) / NM_THREADS), NM_THREADS, 0, s>>>(
        static_cast<double*>(dxr), static_cast<double*>(dxi), n, static_cast<double*>(dyr), static_cast<double*>(dyi), m,
        static_cast<long long*>(didx), distance ? static_cast<double*>(ddist) : nullptr);
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(index, didx, static_cast<size_t>(n) * sizeof(long long), cudaMemcpyDeviceToHost, s));
    if (distance) LM_CUDA_TRY(cudaMemcpyAsync(distance, ddist, nb, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(n);
        stats->work_units = static_cast<uint64_t>(n) * static_cast<uint64_t>(m);     // distance evaluations
        stats->kernel_ms = ms;
        stats->launches = 1;
    }
    return LM_OK;
}

}  // extern "C"

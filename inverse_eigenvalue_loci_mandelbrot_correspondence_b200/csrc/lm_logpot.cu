// lm_logpot.cu -- K4a: log-potential of a point cloud on a grid (FP64 pipe: sqrt + log per pair).
//
//   log_potential               Potentials.py:19-27                       (LM_LOGPOT_SUM_SQRT)
//   construct_potential         Laplacian_C-M.py:16-25                    (LM_LOGPOT_NEG_PERTERM)
//   log_potential               Iterative_Variogram_Laplacian.py:102-112  (LM_LOGPOT_SUM_HYPOT)
//   log_potential_from_points   variograms_construct_mandelbrot.py:128-146 (LM_LOGPOT_LOG_INV)
//
// One thread per grid cell accumulates the points sequentially in input order (the order the
// reference's loop over points uses), points are staged through shared memory in chunks so
// the whole CTA reads each coordinate once.  Parity is tolerance based (1e-12 relative): the
// reference's log/sqrt come from numpy's SIMD loops, ours from the CUDA math library.
#include "lm_common.cuh"

#include <math.h>

namespace {

constexpr int LP_THREADS = 256;
constexpr int LP_CHUNK = 512;

template <int V>
__global__ void __launch_bounds__(LP_THREADS) logpot_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                                            long long npts, const double* __restrict__ gx, long long nx,
                                                            const double* __restrict__ gy, long long ny, double eps,
                                                            double* __restrict__ U) {
    __shared__ double spx[LP_CHUNK], spy[LP_CHUNK];
    const long long idx = static_cast<long long>(blockIdx.x) * LP_THREADS + threadIdx.x;
    const bool live = idx < nx * ny;
    const long long j = live ? idx / nx : 0, i = live ? idx - j * nx : 0;
    const double x = gx[i], y = gy[j];
    const double N = static_cast<double>(npts);
    double acc = 0.0;
    for (long long base = 0; base < npts; base += LP_CHUNK) {
        const int m = static_cast<int>(npts - base < LP_CHUNK ? npts - base : LP_CHUNK);
        for (int k = threadIdx.x; k < m; k += LP_THREADS) { spx[k] = px[base + k]; spy[k] = py[base + k]; }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int k = 0; k < m; ++k) {
                const double dx = x - spx[k], dy = y - spy[k];
                if (V == LM_LOGPOT_SUM_SQRT) acc = acc + log(sqrt(dx * dx + dy * dy) + eps);
                else if (V == LM_LOGPOT_NEG_PERTERM) acc = acc - log(sqrt(dx * dx + dy * dy) + eps) / N;
                else if (V == LM_LOGPOT_SUM_HYPOT) acc = acc + log(hypot(dx, dy) + eps);
                else acc = acc + log(1.0 / (hypot(dx, dy) + eps));
            }
        }
        __syncthreads();
    }
    if (live) {
        if (V != LM_LOGPOT_NEG_PERTERM && npts > 0) acc = acc / N;
        U[idx] = acc;
    }
}

}  // namespace

extern "C" {

int32_t lm_log_potential(const double* px, const double* py, int64_t npts,
                         const double* gx, int64_t nx, const double* gy, int64_t ny,
                         double eps, int32_t variant, double* U, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npts >= 0 && nx >= 0 && ny >= 0, "lm_log_potential: negative size");
    LM_REQUIRE(variant >= LM_LOGPOT_SUM_SQRT && variant <= LM_LOGPOT_LOG_INV, "lm_log_potential: unknown variant %d", variant);
    LM_REQUIRE((npts == 0 || (px && py)) && (nx * ny == 0 || (gx && gy && U)), "lm_log_potential: NULL buffer");
    if (stats) *stats = lm_stats{};
    if (nx * ny == 0) return LM_OK;
    cudaStream_t s = nullptr;
    void *dpx, *dpy, *dgx, *dgy, *dU;
    const size_t pb = static_cast<size_t>(npts) * sizeof(double), ub = static_cast<size_t>(nx) * ny * sizeof(double);
    if ((rc = lm::ws_get(lm::WS_IN_A, pb, &dpx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, pb, &dpy)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_XS, nx * sizeof(double), &dgx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_YS, ny * sizeof(double), &dgy)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, ub, &dU)) != LM_OK) return rc;
    if (npts) {
        LM_CUDA_TRY(cudaMemcpyAsync(dpx, px, pb, cudaMemcpyHostToDevice, s));
        LM_CUDA_TRY(cudaMemcpyAsync(dpy, py, pb, cudaMemcpyHostToDevice, s));
    }
    LM_CUDA_TRY(cudaMemcpyAsync(dgx, gx, nx * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dgy, gy, ny * sizeof(double), cudaMemcpyHostToDevice, s));
    const long long cells = static_cast<long long>(nx) * ny;
    const unsigned blocks = static_cast<unsigned>((cells + LP_THREADS - 1) / LP_THREADS);
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
#define LM_LP_LAUNCH(V) logpot_kernel<V><<<blocks, LP_THREADS, 0, s>>>(static_cast<double*>(dpx), static_cast<double*>(dpy), npts, \
        static_cast<double*>(dgx), nx, static_cast<double*>(dgy), ny, eps, static_cast<double*>(dU))
    switch (variant) {
        case LM_LOGPOT_SUM_SQRT: LM_LP_LAUNCH(LM_LOGPOT_SUM_SQRT); break;
        case LM_LOGPOT_NEG_PERTERM: LM_LP_LAUNCH(LM_LOGPOT_NEG_PERTERM); break;
        case LM_LOGPOT_SUM_HYPOT: LM_LP_LAUNCH(LM_LOGPOT_SUM_HYPOT); break;
        default: LM_LP_LAUNCH(LM_LOGPOT_LOG_INV); break;
    }
#undef LM_LP_LAUNCH
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(U, dU, ub, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(cells);
        stats->work_units = static_cast<uint64_t>(cells) * static_cast<uint64_t>(npts);   // (cell, point) pairs
        stats->kernel_ms = ms;
        stats->launches = 1;
    }
    return LM_OK;
}

}  // extern "C"

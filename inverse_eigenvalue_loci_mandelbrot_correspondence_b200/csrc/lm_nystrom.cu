// lm_nystrom.cu -- dense boundary-integral (Nystrom) sums of the Lucas-domain Green function (FP64 pipe).
//
//   g_real :  sl[m]  = sum_n w_n * log(|z_m - zeta_n| + eps)          lucas_to_cardioid_v40_reference.py:240-257
//             (the reference builds log(abs(z[:,None] - bdy[None,:]) + 1e-300) @ (sigma*ds) in chunks of 600 rows
//              and adds -log|z - a| + C + g_shift on the host: those are O(M) and stay there)
//   dPhi   :  s[m]   = sum_n w_n / dz_mn,   dz_mn = z_m - zeta_n, replaced by dz_eps + 0j where |dz_mn| < dz_eps
//                                                                      lucas_to_cardioid_v40_reference.py:201-211
// Same O(M*N) shape as the log-potential (K4a) but with a weight per node, so the product trick does not apply:
// one thread per target point, nodes and weights streamed through shared memory, library log.  The reference's
// sums are BLAS / numpy reductions (pairwise, not left-to-right), so parity is tolerance based (1e-12).
#include "lm_common.cuh"

#include <math.h>

namespace {

constexpr int NY_THREADS = 128;
constexpr int NY_CHUNK = 512;

template <int MODE>      // 0: weighted log sum, 1: weighted Cauchy sum
__global__ void __launch_bounds__(NY_THREADS) nystrom_kernel(const double* __restrict__ zr, const double* __restrict__ zi, long long M,
                                                             const double* __restrict__ br, const double* __restrict__ bi,
                                                             const double* __restrict__ w, long long N, double eps,
                                                             double* __restrict__ out_a, double* __restrict__ out_b) {
    __shared__ double sbr[NY_CHUNK], sbi[NY_CHUNK], sw[NY_CHUNK];
    const long long m = static_cast<long long>(blockIdx.x) * NY_THREADS + threadIdx.x;
    const bool live = m < M;
    const double x = live ? zr[m] : 0.0, y = live ? zi[m] : 0.0;
    double acc_a = 0.0, acc_b = 0.0;
    for (long long base = 0; base < N; base += NY_CHUNK) {
        const int c = static_cast<int>(N - base < NY_CHUNK ? N - base : NY_CHUNK);
        __syncthreads();
        for (int t = threadIdx.x; t < c; t += NY_THREADS) { sbr[t] = br[base + t]; sbi[t] = bi[base + t]; sw[t] = w[base + t]; }
        __syncthreads();
        if (!live) continue;
#pragma unroll 2
        for (int t = 0; t < c; ++t) {
            double dx = x - sbr[t], dy = y - sbi[t];
            if (MODE == 0) {
                acc_a = fma(sw[t], log(hypot(dx, dy) + eps), acc_a);
            } else {
                const double r = hypot(dx, dy);
                if (r < eps) { dx = eps; dy = 0.0; }               // DZ = where(|DZ| < DZ_EPS, DZ_EPS + 0j, DZ)
                const double q = dx * dx + dy * dy;
                acc_a = fma(sw[t], dx / q, acc_a);                 // w / (dx + i dy) = w (dx - i dy) / q
                acc_b = fma(-sw[t], dy / q, acc_b);
            }
        }
    }
    if (live) {
        out_a[m] = acc_a;
        if (MODE == 1) out_b[m] = acc_b;
    }
}

int32_t run_nystrom(int mode, const double* z_re, const double* z_im, int64_t M, const double* b_re, const double* b_im,
                    const double* w, int64_t N, double eps, double* out_a, double* out_b, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(M >= 0 && N >= 0, "lm_nystrom: negative size");
    LM_REQUIRE(M == 0 || (z_re && z_im && out_a && (mode == 0 || out_b)), "lm_nystrom: NULL target buffer");
    LM_REQUIRE(N == 0 || (b_re && b_im && w), "lm_nystrom: NULL node buffer");
    LM_REQUIRE(eps >= 0.0, "lm_nystrom: eps must be >= 0");
    if (stats) *stats = lm_stats{};
    if (M == 0) return LM_OK;
    cudaStream_t s = nullptr;
    const size_t mb = static_cast<size_t>(M) * sizeof(double), nb = static_cast<size_t>(N) * sizeof(double);
    void *dzr, *dzi, *dbr, *dbi, *dw, *doa, *dob;
    if ((rc = lm::ws_get(lm::WS_IN_A, mb, &dzr)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, mb, &dzi)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_C, nb, &dbr)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, nb, &dbi)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_D, nb, &dw)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, mb, &doa)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, mb, &dob)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dzr, z_re, mb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dzi, z_im, mb, cudaMemcpyHostToDevice, s));
    if (N) {
        LM_CUDA_TRY(cudaMemcpyAsync(dbr, b_re, nb, cudaMemcpyHostToDevice, s));
        LM_CUDA_TRY(cudaMemcpyAsync(dbi, b_im, nb, cudaMemcpyHostToDevice, s));
        LM_CUDA_TRY(cudaMemcpyAsync(dw, w, nb, cudaMemcpyHostToDevice, s));
    }
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    const unsigned blocks = static_cast<unsigned>((M + NY_THREADS - 1) / NY_THREADS);
    if (mode == 0)
        nystrom_kernel<0><<<blocks, NY_THREADS, 0, s>>>(static_cast<double*>(dzr), static_cast<double*>(dzi), M, static_cast<double*>(dbr),
                                                        static_cast<double*>(dbi), static_cast<double*>(dw), N, eps,
                                                        static_cast<double*>(doa), static_cast<double*>(dob));
    else
        nystrom_kernel<1><<<blocks, NY_THREADS, 0, s>>>(static_cast<double*>(dzr), static_cast<double*>(dzi), M, static_cast<double*>(dbr),
                                                        static_cast<double*>(dbi), static_cast<double*>(dw), N, eps,
                                                        static_cast<double*>(doa), static_cast<double*>(dob));
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(out_a, doa, mb, cudaMemcpyDeviceToHost, s));
    if (mode == 1) LM_CUDA_TRY(cudaMemcpyAsync(out_b, dob, mb, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(M);
        stats->work_units = static_cast<uint64_t>(M) * static_cast<uint64_t>(N);
        stats->kernel_ms = ms;
        stats->launches = 1;
    }
    return LM_OK;
}

}  // namespace

extern "C" {

int32_t lm_weighted_log_sum(const double* z_re, const double* z_im, int64_t M,
                            const double* node_re, const double* node_im, const double* weight, int64_t N,
                            double eps, double* out, lm_stats* stats) {
    return run_nystrom(0, z_re, z_im, M, node_re, node_im, weight, N, eps, out, nullptr, stats);
}

int32_t lm_weighted_cauchy_sum(const double* z_re, const double* z_im, int64_t M,
                               const double* node_re, const double* node_im, const double* weight, int64_t N,
                               double dz_eps, double* out_re, double* out_im, lm_stats* stats) {
    return run_nystrom(1, z_re, z_im, M, node_re, node_im, weight, N, dz_eps, out_re, out_im, stats);
}

}  // extern "C"

// lm_logpot.cu -- K4a: log-potential of a point cloud on a grid (FP64 pipe).
//
//   log_potential               Potentials.py:19-27                       (LM_LOGPOT_SUM_SQRT)
//   construct_potential         Laplacian_C-M.py:16-25                    (LM_LOGPOT_NEG_PERTERM)
//   log_potential               Iterative_Variogram_Laplacian.py:102-112  (LM_LOGPOT_SUM_HYPOT)
//   log_potential_from_points   variograms_construct_mandelbrot.py:128-146 (LM_LOGPOT_LOG_INV)
//
// All four are  U = +-(1/N) sum_p log(|z - p| + eps)  up to the association of the sum, and the
// reference's own variants agree with each other only to rounding, so parity is
// tolerance based (1e-12) and the kernel is free to reorder.  What it exploits:
//
//  * sum of logs = log of product.  A thread multiplies the 16 terms of a point group (the product of
//    16 terms in [eps, ~10] cannot leave the double range; a group whose product is not a normal
//    number is redone term by term with the library functions), folds the group into a running
//    (mantissa, integer exponent) product and takes ONE log at the very end.
//  * eps <= 1e-10 (three of the four variants use 1e-12):  log(s + eps) = log(s) + log1p(eps/s),
//    and for eps/s <= 1e-8 the second term is eps/s to 5e-17.  So the group accumulates the
//    product of r^2 = dx^2 + dy^2 (no square root at all) and the sum of MUFU.RSQ64H(r^2)
//    (a 22-bit 1/s is plenty for a term of relative size <= 1e-8); a group containing a pair
//    with s < eps*1e8 takes the exact path.  6 FP64 instructions per (cell, point) pair become
//    ~4.5 with four cells of one grid row per thread (dy^2 shared).
//  * 2-D decomposition: (tiles of 4 x 256 cells) x (point splits), partial sums reduced in a fixed
//    order by a second kernel, so the grid fills the chip for a 400^2 grid and results are
//    deterministic.  The same partial-sum entry point serves the multi-GPU path, where every rank
//    holds a slice of the cloud and the per-cell sums are all-reduced (NCCL) before `finish`.
#include "lm_common.cuh"

#include <math.h>

namespace {

constexpr int LP_THREADS = 256;
#ifndef LM_K4A_CPT
#define LM_K4A_CPT 4
#endif
constexpr int LP_CPT = LM_K4A_CPT;   // cells per thread, consecutive in x (they share dy^2)
constexpr int LP_CHUNK = 1024;       // points staged in shared memory per round
constexpr int LP_GROUP = 16;         // terms per group product (range check / exponent extraction once per group)

__device__ __forceinline__ double rsqrt_approx(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// exact term, as the reference writes it
template <int V>
__device__ __forceinline__ double exact_term(double dx, double dy, double eps) {
    if (V == LM_LOGPOT_SUM_SQRT || V == LM_LOGPOT_NEG_PERTERM) return log(sqrt(dx * dx + dy * dy) + eps);
    return log(hypot(dx, dy) + eps);
}

// sqrt by Goldschmidt iteration from the 22-bit reciprocal square root (<= 2 ulp); x > 0 normal
__device__ __forceinline__ double sqrt_fast(double x) {
    const double y0 = rsqrt_approx(x);
    double g = x * y0, h = 0.5 * y0;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    return fma(g, r, g);
}

// Raw per-cell sums  S[cell] = sum over this split's points of log(|z - p| + eps).
//   partial[split * ncells + cell]
// The running product of a cell is kept as (mantissa in [1,2), integer exponent): after every group
// the exponent field is moved into an integer accumulator, so ONE log per thread is taken at the very
// end (rounding of the mantissa product: <= 1 ulp per factor, far inside the 1e-12 parity bound).
template <int V, bool FAST>
__global__ void __launch_bounds__(LP_THREADS) logpot_partial_kernel(
    const double* __restrict__ px, const double* __restrict__ py, long long npts, long long pts_per_split,
    const double* __restrict__ gx, long long nx, const double* __restrict__ gy, long long ny, double eps,
    double* __restrict__ partial) {
    __shared__ double2 spt[LP_CHUNK];
    const long long tiles_x = (nx + LP_CPT - 1) / LP_CPT;                 // threads per grid row
    const long long tid = static_cast<long long>(blockIdx.x) * LP_THREADS + threadIdx.x;
    const bool live = tid < tiles_x * ny;
    const long long j = live ? tid / tiles_x : 0;
    const long long i0 = live ? (tid - j * tiles_x) * LP_CPT : 0;
    double x[LP_CPT];
#pragma unroll
    for (int c = 0; c < LP_CPT; ++c) x[c] = gx[(i0 + c < nx) ? i0 + c : nx - 1];
    const double y = gy[j];
    const long long p_begin = static_cast<long long>(blockIdx.y) * pts_per_split;
    const long long p_end = (p_begin + pts_per_split < npts) ? p_begin + pts_per_split : npts;
    // a pair is "near" when s < eps * 1e8, i.e. rsqrt(r^2) > 1e-8 / eps; a group whose rsqrt sum stays
    // below that bound cannot contain one
    const double near_sum = FAST ? 1e-8 / eps : 0.0;

    double P[LP_CPT], E[LP_CPT], slow[LP_CPT], eacc[LP_CPT];
    int esum[LP_CPT];
#pragma unroll
    for (int c = 0; c < LP_CPT; ++c) { P[c] = 1.0; E[c] = 0.0; slow[c] = 0.0; eacc[c] = 0.0; esum[c] = 0; }

    for (long long base = p_begin; base < p_end; base += LP_CHUNK) {
        const int m = static_cast<int>(p_end - base < LP_CHUNK ? p_end - base : LP_CHUNK);
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += LP_THREADS) spt[k] = make_double2(px[base + k], py[base + k]);
        __syncthreads();
        if (!live) continue;
        int k0 = 0;
        for (; k0 + LP_GROUP <= m; k0 += LP_GROUP) {
            double Pg[LP_CPT], Eg[LP_CPT];
#pragma unroll
            for (int c = 0; c < LP_CPT; ++c) { Pg[c] = 1.0; Eg[c] = 0.0; }
#pragma unroll
            for (int k = 0; k < LP_GROUP; ++k) {
                const double2 p = spt[k0 + k];
                const double dy = y - p.y;
                const double dy2 = dy * dy;
#pragma unroll
                for (int c = 0; c < LP_CPT; ++c) {
                    const double dx = x[c] - p.x;
                    const double r2 = fma(dx, dx, dy2);
                    if (FAST) {
                        Pg[c] *= r2;
                        Eg[c] += rsqrt_approx(r2);
                    } else {
                        Pg[c] *= sqrt_fast(r2) + eps;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < LP_CPT; ++c) {
                // the group product must be a comfortably normal number (biased exponent in [100, 1946]; this
                // also rejects 0, Inf and NaN) and (FAST) no pair of the group may be near
                const unsigned ex = (static_cast<unsigned>(__double2hiint(Pg[c])) >> 20) & 0x7ffu;
                const bool ok = (ex - 100u <= 1846u) && (!FAST || Eg[c] < near_sum);
                if (ok) {
                    const double t = P[c] * Pg[c];                  // P in [1,2): cannot leave the range either
                    const int hi = __double2hiint(t);
                    const int e = (hi >> 20) - 1023;
                    esum[c] += e;
                    P[c] = __hiloint2double(hi - (e << 20), __double2loint(t));
                    if (FAST) E[c] += Eg[c];
                } else {
                    for (int k = 0; k < LP_GROUP; ++k) {
                        const double2 p = spt[k0 + k];
                        slow[c] += exact_term<V>(x[c] - p.x, y - p.y, eps);
                    }
                }
            }
        }
        for (; k0 < m; ++k0) {
            const double2 p = spt[k0];
#pragma unroll
            for (int c = 0; c < LP_CPT; ++c) slow[c] += exact_term<V>(x[c] - p.x, y - p.y, eps);
        }
#pragma unroll
        for (int c = 0; c < LP_CPT; ++c) { eacc[c] += static_cast<double>(esum[c]); esum[c] = 0; }
    }
    if (live) {
        double* out = partial + static_cast<long long>(blockIdx.y) * nx * ny + j * nx + i0;
#pragma unroll
        for (int c = 0; c < LP_CPT; ++c) {
            if (i0 + c >= nx) continue;
            const double lp = fma(eacc[c], 0.693147180559945309417232, log(P[c]));    // log of the whole product
            out[c] = FAST ? fma(eps, E[c], fma(0.5, lp, slow[c])) : lp + slow[c];
        }
    }
}

// U[cell] = scale * sum_s partial[s][cell]   (fixed order over the splits: deterministic)
__global__ void __launch_bounds__(256) logpot_reduce_kernel(const double* __restrict__ partial, int nsplit, long long ncells,
                                                            double scale, double* __restrict__ U) {
    const long long c = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (c >= ncells) return;
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += partial[static_cast<long long>(k) * ncells + c];
    U[c] = s * scale;
}

int split_count(long long npts, long long nx, long long ny) {
    const long long tiles = ((nx + LP_CPT - 1) / LP_CPT) * ny;
    const long long cell_blocks = (tiles + LP_THREADS - 1) / LP_THREADS;
    const long long want_blocks = static_cast<long long>(lm::sm_count()) * 8 * 2;        // two waves of 8 CTAs/SM
    long long s = (want_blocks + cell_blocks - 1) / cell_blocks;
    const long long max_by_pts = (npts + LP_CHUNK - 1) / LP_CHUNK;                       // at least one chunk per split
    if (s > max_by_pts) s = max_by_pts;
    if (s > 1024) s = 1024;
    if (s < 1) s = 1;
    return static_cast<int>(s);
}

// sum_dev[cell] = sum_p log(|z_cell - p| + eps) over the npts device-resident points
int32_t logpot_sums_dev(const double* px, const double* py, long long npts, const double* gx, long long nx,
                        const double* gy, long long ny, double eps, int32_t variant, double scale,
                        double* out_dev, int* launches, cudaStream_t s) {
    const long long ncells = nx * ny;
    const int nsplit = split_count(npts, nx, ny);
    void* dpart = nullptr;
    int32_t rc;
    if ((rc = lm::ws_get(lm::WS_LOGPOT_PART, static_cast<size_t>(nsplit) * ncells * sizeof(double), &dpart)) != LM_OK) return rc;
    const long long pts_per_split = ((npts + nsplit - 1) / nsplit + LP_GROUP - 1) / LP_GROUP * LP_GROUP;
    const long long tiles = ((nx + LP_CPT - 1) / LP_CPT) * ny;
    const dim3 grid(static_cast<unsigned>((tiles + LP_THREADS - 1) / LP_THREADS), static_cast<unsigned>(nsplit));
    const bool fast = eps > 0.0 && eps <= 1e-10;
#define LM_LP_LAUNCH(V)                                                                                             \
    do {                                                                                                            \
        if (fast) logpot_partial_kernel<V, true><<<grid, LP_THREADS, 0, s>>>(px, py, npts, pts_per_split, gx, nx, gy, ny, eps, \
                                                                             static_cast<double*>(dpart));            \
        else logpot_partial_kernel<V, false><<<grid, LP_THREADS, 0, s>>>(px, py, npts, pts_per_split, gx, nx, gy, ny, eps,     \
                                                                         static_cast<double*>(dpart));                \
    } while (0)
    switch (variant) {
        case LM_LOGPOT_SUM_SQRT:    LM_LP_LAUNCH(LM_LOGPOT_SUM_SQRT); break;
        case LM_LOGPOT_NEG_PERTERM: LM_LP_LAUNCH(LM_LOGPOT_NEG_PERTERM); break;
        case LM_LOGPOT_SUM_HYPOT:   LM_LP_LAUNCH(LM_LOGPOT_SUM_HYPOT); break;
        default:                    LM_LP_LAUNCH(LM_LOGPOT_LOG_INV); break;
    }
#undef LM_LP_LAUNCH
    LM_CUDA_TRY(cudaGetLastError());
    logpot_reduce_kernel<<<static_cast<unsigned>((ncells + 255) / 256), 256, 0, s>>>(static_cast<double*>(dpart), nsplit, ncells,
                                                                                      scale, out_dev);
    LM_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 2;
    return LM_OK;
}

// the variant's normalisation of the raw sum: +1/N, or -1/N for the two "negative log" forms
double variant_scale(int32_t variant, long long n_total) {
    if (n_total <= 0) return 0.0;
    const double inv = 1.0 / static_cast<double>(n_total);
    return (variant == LM_LOGPOT_NEG_PERTERM || variant == LM_LOGPOT_LOG_INV) ? -inv : inv;
}

}  // namespace

extern "C" {

int32_t lm_log_potential_sums_dev(const double* px_dev, const double* py_dev, int64_t npts,
                                  const double* gx_dev, int64_t nx, const double* gy_dev, int64_t ny,
                                  double eps, int32_t variant, double* sums_dev, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npts >= 0 && nx >= 0 && ny >= 0, "lm_log_potential_sums_dev: negative size");
    LM_REQUIRE(variant >= LM_LOGPOT_SUM_SQRT && variant <= LM_LOGPOT_LOG_INV, "lm_log_potential_sums_dev: unknown variant %d", variant);
    LM_REQUIRE((npts == 0 || (px_dev && py_dev)) && (nx * ny == 0 || (gx_dev && gy_dev && sums_dev)),
               "lm_log_potential_sums_dev: NULL buffer");
    if (nx * ny == 0) return LM_OK;
    cudaStream_t s = lm::as_stream(stream);
    if (npts == 0) {
        LM_CUDA_TRY(cudaMemsetAsync(sums_dev, 0, static_cast<size_t>(nx) * ny * sizeof(double), s));
        return LM_OK;
    }
    return logpot_sums_dev(px_dev, py_dev, npts, gx_dev, nx, gy_dev, ny, eps, variant, 1.0, sums_dev, nullptr, s);
}

int32_t lm_log_potential_finish_dev(const double* sums_dev, int64_t ncells, int64_t n_total_points, int32_t variant,
                                    double* U_dev, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(ncells >= 0 && n_total_points >= 0, "lm_log_potential_finish_dev: negative size");
    LM_REQUIRE(variant >= LM_LOGPOT_SUM_SQRT && variant <= LM_LOGPOT_LOG_INV, "lm_log_potential_finish_dev: unknown variant %d", variant);
    LM_REQUIRE(ncells == 0 || (sums_dev && U_dev), "lm_log_potential_finish_dev: NULL buffer");
    if (ncells == 0) return LM_OK;
    logpot_reduce_kernel<<<static_cast<unsigned>((ncells + 255) / 256), 256, 0, lm::as_stream(stream)>>>(
        sums_dev, 1, ncells, variant_scale(variant, n_total_points), U_dev);
    LM_CUDA_TRY(cudaGetLastError());
    return LM_OK;
}

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
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    int launches = 0;
    if (npts == 0) {
        LM_CUDA_TRY(cudaMemsetAsync(dU, 0, ub, s));
    } else {
        rc = logpot_sums_dev(static_cast<double*>(dpx), static_cast<double*>(dpy), npts, static_cast<double*>(dgx), nx,
                             static_cast<double*>(dgy), ny, eps, variant, variant_scale(variant, npts),
                             static_cast<double*>(dU), &launches, s);
        if (rc != LM_OK) return rc;
    }
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(U, dU, ub, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(nx) * static_cast<uint64_t>(ny);
        stats->work_units = stats->items * static_cast<uint64_t>(npts);   // (cell, point) pairs
        stats->kernel_ms = ms;
        stats->launches = launches;
    }
    return LM_OK;
}

}  // extern "C"

// lm_distance.cu -- K1b: distance-estimator grids, and the optional fp32 dwell variant.
//
//   mandelbrot_distance_estimator   construct_stage1_clean.py:50-58           (LM_DE_SCALAR)
//   mandelbrot_distance_estimator   variograms_construct_mandelbrot.py:61-88  (LM_DE_FIRST_ESCAPE)
//   mandelbrot_distance_estimator   tci_construct_mandelbrot.py:21-39,
//                                   tci_construct_mandelbrot_v002_fixed.py:35-47 (LM_DE_FINAL_DZ: z from the
//                                   first escape, dz from the END of the loop)
//   the same function as numpy really evaluates it                            (LM_DE_FINAL_DZ_NUMPY): numpy's SIMD
//                                   complex multiply is real = fma(ar, br, -(ai*bi)), imag = fma(ar, bi, ai*br)
//                                   (npyv_muladdsub); with that recipe the escape mask and the d == 0 pattern of
//                                   the stock tracker module -- hence sample_mandelbrot_boundary() -- are bit-exact
//
// The reference runs these on small grids (120x80 ... 912^2, max_iter <= 500), so a plain
// one-thread-per-pixel kernel is used; the recurrences are unfused (__dmul_rn/__dadd_rn) in
// CPython's operation order: dz <- (2 z) dz + 1 first, then z <- z z + c.
#include "lm_common.cuh"

#include <math.h>

namespace {

// one sweep dz <- (2 z) dz + 1, z <- z z + c.  NP = false: CPython's unfused complex arithmetic; NP = true: numpy's SIMD
// complex multiply (a * b: real = fma(ar, br, -(ai*bi)), imag = fma(ar, bi, ai*br)); 2 * z is exact for finite z.
template <bool NP>
__device__ __forceinline__ void de_step(double& zr, double& zi, double& dr, double& di, double cr, double ci) {
    const double tr = __dmul_rn(2.0, zr), ti = __dmul_rn(2.0, zi);
    double ndr, ndi, re, im;
    if (NP) {
        ndr = __dadd_rn(__fma_rn(tr, dr, -__dmul_rn(ti, di)), 1.0);
        ndi = __fma_rn(tr, di, __dmul_rn(ti, dr));
        re = __fma_rn(zr, zr, -__dmul_rn(zi, zi));
        im = __fma_rn(zr, zi, __dmul_rn(zi, zr));
    } else {
        ndr = __dadd_rn(__dsub_rn(__dmul_rn(tr, dr), __dmul_rn(ti, di)), 1.0);
        ndi = __dadd_rn(__dmul_rn(tr, di), __dmul_rn(ti, dr));
        re = __dsub_rn(__dmul_rn(zr, zr), __dmul_rn(zi, zi));
        const double pp = __dmul_rn(zr, zi);
        im = __dadd_rn(pp, pp);
    }
    dr = ndr; di = ndi;
    zr = __dadd_rn(re, cr);
    zi = __dadd_rn(im, ci);
}

template <int VARIANT>
__global__ void __launch_bounds__(256) distance_kernel(const double* __restrict__ xs, long long nx,
                                                       const double* __restrict__ ys, long long ny, int max_iter,
                                                       double bailout, double eps, double* __restrict__ dist,
                                                       unsigned char* __restrict__ escaped) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= nx * ny) return;
    const long long j = idx / nx, i = idx - j * nx;
    const double cr = xs[i], ci = ys[j];
    double zr = 0.0, zi = 0.0, dr = (VARIANT == LM_DE_SCALAR) ? 0.0 : 1.0, di = 0.0;
    double d = 0.0;
    bool esc = false;
    const double loop_thr = bailout * bailout * (1.0 - 1e-9);
    constexpr bool NP = (VARIANT == LM_DE_FINAL_DZ_NUMPY);
    for (int n = 0; n < max_iter; ++n) {
        de_step<NP>(zr, zi, dr, di, cr, ci);
        const double m = __dadd_rn(__dmul_rn(zr, zr), __dmul_rn(zi, zi));
        if (!(m <= loop_thr)) {                       // candidate (also catches overflow); hypot decides
            const double az = hypot(zr, zi);
            if (az > bailout) {
                esc = true;
                if (VARIANT == LM_DE_SCALAR) {
                    const double adz = hypot(dr, di);
                    d = __ddiv_rn(__dmul_rn(az, log(az)), adz > 1e-16 ? adz : 1e-16);
                } else if (VARIANT == LM_DE_FIRST_ESCAPE) {
                    const double qr = __dsub_rn(__dmul_rn(__dmul_rn(2.0, zr), dr), __dmul_rn(__dmul_rn(2.0, zi), di));
                    const double qi = __dadd_rn(__dmul_rn(__dmul_rn(2.0, zr), di), __dmul_rn(__dmul_rn(2.0, zi), dr));
                    const double den0 = hypot(qr, qi);
                    const double den = den0 > eps ? den0 : eps;
                    const double num = __dmul_rn(log(az > 1.0 ? az : 1.0), az);
                    d = isnan(den0) ? nan("") : __ddiv_rn(num, den);
                    if (!isfinite(d)) d = 0.0;
                } else {
                    // LM_DE_FINAL_DZ: keep z of this first escape, run dz (and z) on to the end of the loop
                    const double ezr = zr, ezi = zi;
                    for (int m2 = n + 1; m2 < max_iter; ++m2) {
                        de_step<NP>(zr, zi, dr, di, cr, ci);
                        if (!(isfinite(dr) && isfinite(di))) break;     // stays non-finite: d = 0
                    }
                    d = 0.0;
                    if (isfinite(dr) && isfinite(di)) {
                        const double er = __dmul_rn(2.0, ezr), ei = __dmul_rn(2.0, ezi);
                        const double qr = NP ? __fma_rn(er, dr, -__dmul_rn(ei, di)) : __dsub_rn(__dmul_rn(er, dr), __dmul_rn(ei, di));
                        const double qi = NP ? __fma_rn(er, di, __dmul_rn(ei, dr)) : __dadd_rn(__dmul_rn(er, di), __dmul_rn(ei, dr));
                        const double den0 = hypot(qr, qi);
                        if (isfinite(den0)) {
                            d = __ddiv_rn(__dmul_rn(log(az), az), den0 > eps ? den0 : eps);
                            if (!isfinite(d)) d = 0.0;
                        }
                    }
                }
                break;
            }
        }
    }
    dist[idx] = d;
    if (escaped) escaped[idx] = esc ? 1 : 0;
}

}  // namespace

extern "C" {

int32_t lm_distance_grid_f64(const double* xs, int64_t nx, const double* ys, int64_t ny,
                             int32_t max_iter, double bailout, double eps, int32_t variant,
                             double* dist, uint8_t* escaped, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(xs && ys && dist, "lm_distance_grid_f64: NULL buffer");
    LM_REQUIRE(nx >= 0 && ny >= 0 && max_iter >= 0, "lm_distance_grid_f64: negative size");
    LM_REQUIRE(bailout > 0.0 && bailout < 1e150, "lm_distance_grid_f64: bailout out of range");
    LM_REQUIRE(variant >= LM_DE_SCALAR && variant <= LM_DE_FINAL_DZ_NUMPY, "lm_distance_grid_f64: unknown variant %d", variant);
    if (stats) *stats = lm_stats{};
    if (nx * ny == 0) return LM_OK;
    cudaStream_t s = nullptr;
    const size_t npx = static_cast<size_t>(nx) * ny;
    void *dxs, *dys, *dd, *de;
    if ((rc = lm::ws_get(lm::WS_XS, nx * sizeof(double), &dxs)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_YS, ny * sizeof(double), &dys)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_FIELD, npx * sizeof(double), &dd)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, npx, &de)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dxs, xs, nx * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dys, ys, ny * sizeof(double), cudaMemcpyHostToDevice, s));
    const unsigned blocks = static_cast<unsigned>((npx + 255) / 256);
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    if (variant == LM_DE_SCALAR)
        distance_kernel<LM_DE_SCALAR><<<blocks, 256, 0, s>>>(static_cast<double*>(dxs), nx, static_cast<double*>(dys), ny,
                                                             max_iter, bailout, eps, static_cast<double*>(dd),
                                                             static_cast<unsigned char*>(de));
    else if (variant == LM_DE_FIRST_ESCAPE)
        distance_kernel<LM_DE_FIRST_ESCAPE><<<blocks, 256, 0, s>>>(static_cast<double*>(dxs), nx, static_cast<double*>(dys), ny,
                                                                   max_iter, bailout, eps, static_cast<double*>(dd),
                                                                   static_cast<unsigned char*>(de));
    else if (variant == LM_DE_FINAL_DZ_NUMPY)
        distance_kernel<LM_DE_FINAL_DZ_NUMPY><<<blocks, 256, 0, s>>>(static_cast<double*>(dxs), nx, static_cast<double*>(dys), ny,
                                                                     max_iter, bailout, eps, static_cast<double*>(dd),
                                                                     static_cast<unsigned char*>(de));
    else
        distance_kernel<LM_DE_FINAL_DZ><<<blocks, 256, 0, s>>>(static_cast<double*>(dxs), nx, static_cast<double*>(dys), ny,
                                                               max_iter, bailout, eps, static_cast<double*>(dd),
                                                               static_cast<unsigned char*>(de));
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dist, dd, npx * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (escaped) LM_CUDA_TRY(cudaMemcpyAsync(escaped, de, npx, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) { stats->items = npx; stats->kernel_ms = ms; stats->launches = 1; }
    return LM_OK;
}

}  // extern "C"

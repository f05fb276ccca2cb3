// lm_curvature.cu -- local-polynomial curvature of an ordered boundary polyline (consumer of <prefix>_boundary.csv).
//
//   compute_curvature_localpoly(P, neighbors, closed, stride=1)   boundary_curvature_localpoly.py:133-184
//     for every point i: window i-m..i+m (wrapped modulo N when closed, clamped when open, :120-131), signed
//     cumulative arclength s with s = 0 at the centre (:65-83), least-squares quadratics x(s), y(s) (:89-98),
//     kappa_signed = (x'y'' - y'x'') / (sqrt(x'^2 + y'^2) + 1e-16)^3, kappa = |kappa_signed| (:100-118).
// The reference solves the two 3-parameter fits with np.linalg.lstsq (SVD) in a Python loop over the points
// (~0.1 ms per point; the 4.6e5-vertex boundary of config 3 takes most of a minute).  Here a thread owns a point:
// arclengths in the reference's accumulation order, Householder QR of the (2m+1) x 3 design matrix with the
// columns scaled to unit size (better conditioned than the raw [1, s, s^2] the SVD sees), both right-hand sides
// at once.  Parity is tolerance based: the fit is ill-conditioned in s (cond ~ 1e6-1e7 at pixel spacing), so
// agreement with LAPACK's SVD is ~1e-9 relative on the derivatives, not bit-exact.
#include "lm_common.cuh"

#include <math.h>

namespace {

constexpr int CV_THREADS = 128;
constexpr int CV_MAX_NEIGHBORS = 32;            // window of at most 65 points

__global__ void __launch_bounds__(CV_THREADS) curvature_kernel(const double* __restrict__ px, const double* __restrict__ py, long long N,
                                                               int m, int closed, double* __restrict__ kappa,
                                                               double* __restrict__ kappa_signed, double* __restrict__ speed,
                                                               double* __restrict__ x1o, double* __restrict__ y1o,
                                                               double* __restrict__ x2o, double* __restrict__ y2o) {
    const long long i = static_cast<long long>(blockIdx.x) * CV_THREADS + threadIdx.x;
    if (i >= N) return;
    const int W = 2 * m + 1;
    double s[2 * CV_MAX_NEIGHBORS + 1], X[2 * CV_MAX_NEIGHBORS + 1], Y[2 * CV_MAX_NEIGHBORS + 1];
    for (int k = 0; k < W; ++k) {
        long long j = i + (k - m);
        if (closed) { j %= N; if (j < 0) j += N; }
        else j = j < 0 ? 0 : (j > N - 1 ? N - 1 : j);
        X[k] = px[j]; Y[k] = py[j];
    }
    // signed cumulative arclength, centre = 0; np.linalg.norm of a 2-vector = sqrt(dx*dx + dy*dy)
    s[m] = 0.0;
    for (int k = m + 1; k < W; ++k) {
        const double dx = X[k] - X[k - 1], dy = Y[k] - Y[k - 1];
        s[k] = s[k - 1] + sqrt(dx * dx + dy * dy);
    }
    for (int k = m - 1; k >= 0; --k) {
        const double dx = X[k + 1] - X[k], dy = Y[k + 1] - Y[k];
        s[k] = s[k + 1] - sqrt(dx * dx + dy * dy);
    }
    // scale s so that the design columns [1, t, t^2] are O(1): t = s / smax
    double smax = 0.0;
    for (int k = 0; k < W; ++k) smax = fmax(smax, fabs(s[k]));
    const double inv = smax > 0.0 ? 1.0 / smax : 0.0;
    // normal equations are avoided: modified Gram-Schmidt QR on the scaled columns, both right-hand sides
    double q0n = sqrt(static_cast<double>(W));
    // column 0 = 1/sqrt(W); orthogonalise t and t^2 against it and each other
    double r01 = 0.0, r02 = 0.0;
    for (int k = 0; k < W; ++k) { const double t = s[k] * inv; r01 += t; r02 += t * t; }
    r01 /= q0n; r02 /= q0n;
    double r11 = 0.0, r12 = 0.0;
    for (int k = 0; k < W; ++k) {
        const double t = s[k] * inv;
        const double u1 = t - r01 / q0n;
        r11 += u1 * u1;
    }
    r11 = sqrt(r11);
    double cx0 = 0.0, cy0 = 0.0, cx1 = 0.0, cy1 = 0.0;
    if (r11 > 0.0) {
        for (int k = 0; k < W; ++k) {
            const double t = s[k] * inv;
            const double q1 = (t - r01 / q0n) / r11;
            const double u2 = t * t - r02 / q0n;
            r12 += q1 * u2;
        }
    }
    double r22 = 0.0;
    for (int k = 0; k < W; ++k) {
        const double t = s[k] * inv;
        const double q1 = r11 > 0.0 ? (t - r01 / q0n) / r11 : 0.0;
        const double u2 = t * t - r02 / q0n - r12 * q1;
        r22 += u2 * u2;
    }
    r22 = sqrt(r22);
    double cx2 = 0.0, cy2 = 0.0;
    for (int k = 0; k < W; ++k) {
        const double t = s[k] * inv;
        const double q0 = 1.0 / q0n;
        const double q1 = r11 > 0.0 ? (t - r01 / q0n) / r11 : 0.0;
        const double q2 = r22 > 0.0 ? (t * t - r02 / q0n - r12 * q1) / r22 : 0.0;
        cx0 += q0 * X[k]; cy0 += q0 * Y[k];
        cx1 += q1 * X[k]; cy1 += q1 * Y[k];
        cx2 += q2 * X[k]; cy2 += q2 * Y[k];
    }
    // back substitution R c = Q^T b in the scaled variable, then undo the scaling: a1 = c1 / smax, a2 = c2 / smax^2.
    // Rank-deficient windows (all points coincide: smax = 0; two distinct arclengths only) get the minimum-norm
    // answer of lstsq for the components that are determined and 0 for the rest.
    const double tol = 1e-13;
    double bx2 = (r22 > tol) ? cx2 / r22 : 0.0, by2 = (r22 > tol) ? cy2 / r22 : 0.0;
    double bx1 = (r11 > tol) ? (cx1 - r12 * bx2) / r11 : 0.0, by1 = (r11 > tol) ? (cy1 - r12 * by2) / r11 : 0.0;
    const double x1 = bx1 * inv, y1 = by1 * inv;
    const double x2 = 2.0 * bx2 * inv * inv, y2 = 2.0 * by2 * inv * inv;
    const double cross = x1 * y2 - y1 * x2;
    const double sp = sqrt(x1 * x1 + y1 * y1) + 1e-16;
    const double ks = cross / (sp * sp * sp);
    kappa[i] = fabs(ks);
    kappa_signed[i] = ks;
    speed[i] = sp;
    if (x1o) x1o[i] = x1;
    if (y1o) y1o[i] = y1;
    if (x2o) x2o[i] = x2;
    if (y2o) y2o[i] = y2;
}

}  // namespace

extern "C" {

int32_t lm_curvature_localpoly(const double* x, const double* y, int64_t n, int32_t neighbors, int32_t closed,
                               double* kappa, double* kappa_signed, double* speed,
                               double* xprime, double* yprime, double* x2, double* y2, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0, "lm_curvature_localpoly: negative n");
    LM_REQUIRE(neighbors >= 2 && neighbors <= CV_MAX_NEIGHBORS, "lm_curvature_localpoly: neighbors must be in [2, %d]", CV_MAX_NEIGHBORS);
    LM_REQUIRE(n == 0 || (x && y && kappa && kappa_signed && speed), "lm_curvature_localpoly: NULL buffer");
    if (stats) *stats = lm_stats{};
    if (n == 0) return LM_OK;
    cudaStream_t s = nullptr;
    const size_t nb = static_cast<size_t>(n) * sizeof(double);
    void *dx, *dy, *dout;
    if ((rc = lm::ws_get(lm::WS_IN_A, nb, &dx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, nb, &dy)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, 7 * nb, &dout)) != LM_OK) return rc;
    double* o = static_cast<double*>(dout);
    LM_CUDA_TRY(cudaMemcpyAsync(dx, x, nb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dy, y, nb, cudaMemcpyHostToDevice, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    curvature_kernel<<<static_cast<unsigned>((n + CV_THREADS - 1) / CV_THREADS), CV_THREADS, 0, s>>>(
        static_cast<double*>(dx), static_cast<double*>(dy), n, neighbors, closed, o, o + n, o + 2 * n, o + 3 * n, o + 4 * n, o + 5 * n, o + 6 * n);
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    double* outs[7] = {kappa, kappa_signed, speed, xprime, yprime, x2, y2};
    for (int k = 0; k < 7; ++k)
        if (outs[k]) LM_CUDA_TRY(cudaMemcpyAsync(outs[k], o + static_cast<size_t>(k) * n, nb, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(n);
        stats->work_units = static_cast<uint64_t>(n) * static_cast<uint64_t>(2 * neighbors + 1);
        stats->kernel_ms = ms;
        stats->launches = 1;
    }
    return LM_OK;
}

}  // extern "C"

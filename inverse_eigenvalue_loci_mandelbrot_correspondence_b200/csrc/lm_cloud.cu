// lm_cloud.cu -- the Lucas-Loci field stage as one host-buffer call (BASELINE.json config 5).
//
// Chains, without leaving HBM between the stages,
//   K3   roots of every polynomial           compute_inverse_eigenvalues[_family], lucas_equipotential_test_v3.py:93-118
//        -> cloud of 1/lambda, |lambda| > tol, concatenated in polynomial order (the order the
//           reference appends them, :98-99)
//   K1d  escape potential at every cloud point        batch_potential, lucas_equipotential_test_v3.py:153-162
//   K4a  log-potential of the cloud on a grid         log_potential, Potentials.py:19-27 (or another LM_LOGPOT_* variant)
//   K4   5-point periodic Laplacian of that field     laplacian, Laplacian_C-M.py:49-59
// which is what the reference does through files (construct_points.csv -> Potentials.py /
// Laplacian_C-M.py).  Built on the library's own _dev entry points; one stream, one sync for
// the cloud size, event-timed stages.
#include "lm_common.cuh"

extern "C" {

int32_t lm_lucas_cloud_fields(const double* toprows, const int32_t* deg, int64_t npoly, int32_t maxdeg, double tol,
                              double* cloud_re, double* cloud_im, int64_t cap_points, int64_t* n_points,
                              int32_t pot_max_iter, double pot_radius, double* g, int64_t* it,
                              const double* gx, int64_t nx, const double* gy, int64_t ny,
                              double eps, int32_t variant, double h, double* U, double* lapU,
                              lm_cloud_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npoly >= 0 && maxdeg >= 1 && cap_points >= 0, "lm_lucas_cloud_fields: bad sizes");
    LM_REQUIRE(npoly == 0 || (toprows && deg), "lm_lucas_cloud_fields: NULL polynomial buffers");
    LM_REQUIRE(n_points != nullptr, "lm_lucas_cloud_fields: n_points is NULL");
    LM_REQUIRE(nx >= 0 && ny >= 0, "lm_lucas_cloud_fields: negative grid size");
    const bool want_field = (U != nullptr || lapU != nullptr) && nx * ny > 0;
    LM_REQUIRE(!want_field || (gx && gy), "lm_lucas_cloud_fields: grid axes are NULL");
    LM_REQUIRE(lapU == nullptr || h > 0.0, "lm_lucas_cloud_fields: h must be positive for the Laplacian");
    const bool want_pot = (g != nullptr || it != nullptr);
    LM_REQUIRE(!want_pot || (pot_max_iter >= 1 && pot_radius > 0.0), "lm_lucas_cloud_fields: bad potential parameters");
    if (stats) *stats = lm_cloud_stats{};
    *n_points = 0;
    cudaStream_t s = nullptr;
    const size_t ncoef = static_cast<size_t>(npoly) * maxdeg;
    void *dtop, *ddeg, *dre, *dim, *dkept, *dstat, *dpx, *dpy;
    if ((rc = lm::ws_get(lm::WS_IN_A, ncoef * sizeof(double), &dtop)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, static_cast<size_t>(npoly) * sizeof(int), &ddeg)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, ncoef * sizeof(double), &dre)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, ncoef * sizeof(double), &dim)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, static_cast<size_t>(npoly) * sizeof(int), &dkept)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dstat)) != LM_OK) return rc;
    // the cloud has at most sum(deg) <= npoly * maxdeg points; size the packed buffers by the host sum
    uint64_t nroots = 0;
    for (int64_t k = 0; k < npoly; ++k) nroots += static_cast<uint64_t>(deg[k] > 0 ? deg[k] : 0);
    const size_t pb = static_cast<size_t>(nroots) * sizeof(double);
    if ((rc = lm::ws_get(lm::WS_CLOUD_A, pb, &dpx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_CLOUD_B, pb, &dpy)) != LM_OK) return rc;

    cudaEvent_t ev[6];
    for (auto& e : ev) LM_CUDA_TRY(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int k = 0; k < 6; ++k) cudaEventDestroy(e[k]); } } guard{ev};

    if (npoly) {
        LM_CUDA_TRY(cudaMemcpyAsync(dtop, toprows, ncoef * sizeof(double), cudaMemcpyHostToDevice, s));
        LM_CUDA_TRY(cudaMemcpyAsync(ddeg, deg, static_cast<size_t>(npoly) * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    int32_t* dflags = static_cast<int32_t*>(dstat);                       // [0] no-convergence, [1] bad degree
    int64_t* dcount = reinterpret_cast<int64_t*>(static_cast<char*>(dstat) + 16);
    uint64_t* dwork = reinterpret_cast<uint64_t*>(static_cast<char*>(dstat) + 32);
    LM_CUDA_TRY(cudaEventRecord(ev[0], s));
    rc = lm_roots_batched_dev(static_cast<double*>(dtop), static_cast<int32_t*>(ddeg), npoly, maxdeg, 1, tol,
                              static_cast<double*>(dre), static_cast<double*>(dim), static_cast<int32_t*>(dkept), nullptr,
                              dflags, s);
    if (rc != LM_OK) return rc;
    LM_CUDA_TRY(cudaEventRecord(ev[1], s));
    rc = lm_cloud_compact_dev(static_cast<double*>(dre), static_cast<double*>(dim), static_cast<int32_t*>(dkept), npoly, maxdeg,
                              static_cast<double*>(dpx), static_cast<double*>(dpy), static_cast<int64_t>(nroots), dcount, s);
    if (rc != LM_OK) return rc;
    LM_CUDA_TRY(cudaEventRecord(ev[2], s));
    int32_t flags[2] = {0, 0};
    int64_t count = 0;
    LM_CUDA_TRY(cudaMemcpyAsync(flags, dflags, sizeof(flags), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaMemcpyAsync(&count, dcount, sizeof(count), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (flags[1]) return lm::fail(LM_E_INVALID, "lm_lucas_cloud_fields: a degree outside [1, %d]", maxdeg);
    *n_points = count;
    if ((cloud_re || cloud_im || want_pot) && count > cap_points)
        return lm::fail(LM_E_CAP, "lm_lucas_cloud_fields: the cloud has %lld points, capacity %lld",
                        static_cast<long long>(count), static_cast<long long>(cap_points));
    const size_t cb = static_cast<size_t>(count) * sizeof(double);

    int launches = 0;
    void *dg = nullptr, *dit = nullptr;
    if (want_pot && count) {
        if ((rc = lm::ws_get(lm::WS_CLOUD_C, cb, &dg)) != LM_OK) return rc;
        if ((rc = lm::ws_get(lm::WS_CLOUD_D, cb, &dit)) != LM_OK) return rc;
        rc = lm_escape_points_f64_dev(static_cast<double*>(dpx), static_cast<double*>(dpy), count, pot_max_iter, pot_radius,
                                      g ? static_cast<double*>(dg) : nullptr, it ? static_cast<int64_t*>(dit) : nullptr,
                                      nullptr, nullptr, dwork, s);
        if (rc != LM_OK) return rc;
        ++launches;
    }
    LM_CUDA_TRY(cudaEventRecord(ev[3], s));
    void *dgx = nullptr, *dgy = nullptr, *dU = nullptr, *dL = nullptr;
    const size_t ub = static_cast<size_t>(nx) * ny * sizeof(double);
    if (want_field) {
        if ((rc = lm::ws_get(lm::WS_XS, nx * sizeof(double), &dgx)) != LM_OK) return rc;
        if ((rc = lm::ws_get(lm::WS_YS, ny * sizeof(double), &dgy)) != LM_OK) return rc;
        if ((rc = lm::ws_get(lm::WS_FIELD, ub, &dU)) != LM_OK) return rc;
        if ((rc = lm::ws_get(lm::WS_OUT_F64, ub, &dL)) != LM_OK) return rc;
        LM_CUDA_TRY(cudaMemcpyAsync(dgx, gx, nx * sizeof(double), cudaMemcpyHostToDevice, s));
        LM_CUDA_TRY(cudaMemcpyAsync(dgy, gy, ny * sizeof(double), cudaMemcpyHostToDevice, s));
        rc = lm_log_potential_sums_dev(static_cast<double*>(dpx), static_cast<double*>(dpy), count, static_cast<double*>(dgx), nx,
                                       static_cast<double*>(dgy), ny, eps, variant, static_cast<double*>(dL), s);
        if (rc != LM_OK) return rc;
        rc = lm_log_potential_finish_dev(static_cast<double*>(dL), nx * ny, count, variant, static_cast<double*>(dU), s);
        if (rc != LM_OK) return rc;
        launches += 3;
    }
    LM_CUDA_TRY(cudaEventRecord(ev[4], s));
    if (want_field && lapU) {
        rc = lm_laplacian5_periodic_dev(static_cast<double*>(dU), ny, nx, h, static_cast<double*>(dL), s);
        if (rc != LM_OK) return rc;
        ++launches;
    }
    LM_CUDA_TRY(cudaEventRecord(ev[5], s));
    // the cloud goes back on a second stream while the field kernels (already enqueued) run
    cudaStream_t s_copy = nullptr;
    if ((cloud_re || cloud_im) && count) {
        LM_CUDA_TRY(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
        cudaError_t ce = cudaSuccess;
        if (cloud_re) ce = cudaMemcpyAsync(cloud_re, dpx, cb, cudaMemcpyDeviceToHost, s_copy);
        if (ce == cudaSuccess && cloud_im) ce = cudaMemcpyAsync(cloud_im, dpy, cb, cudaMemcpyDeviceToHost, s_copy);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s_copy);
        cudaStreamDestroy(s_copy);
        LM_CUDA_TRY(ce);
    }
    if (want_pot && count) {
        if (g) LM_CUDA_TRY(cudaMemcpyAsync(g, dg, cb, cudaMemcpyDeviceToHost, s));
        if (it) LM_CUDA_TRY(cudaMemcpyAsync(it, dit, cb, cudaMemcpyDeviceToHost, s));
    }
    if (want_field && U) LM_CUDA_TRY(cudaMemcpyAsync(U, dU, ub, cudaMemcpyDeviceToHost, s));
    if (want_field && lapU) LM_CUDA_TRY(cudaMemcpyAsync(lapU, dL, ub, cudaMemcpyDeviceToHost, s));
    uint64_t work = 0;
    if (want_pot && count) LM_CUDA_TRY(cudaMemcpyAsync(&work, dwork, sizeof(work), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->n_roots = nroots;
        stats->n_points = static_cast<uint64_t>(count);
        stats->potential_work = work;
        stats->pairs = want_field ? static_cast<uint64_t>(count) * static_cast<uint64_t>(nx) * static_cast<uint64_t>(ny) : 0;
        cudaEventElapsedTime(&stats->roots_ms, ev[0], ev[1]);
        cudaEventElapsedTime(&stats->compact_ms, ev[1], ev[2]);
        cudaEventElapsedTime(&stats->potential_ms, ev[2], ev[3]);
        cudaEventElapsedTime(&stats->logpot_ms, ev[3], ev[4]);
        cudaEventElapsedTime(&stats->stencil_ms, ev[4], ev[5]);
        stats->launches = launches + 10;      // K3: 3 sort + up to 3 solver launches, compaction: 3
    }
    if (flags[0])
        return lm::fail(LM_E_NOCONV, "lm_lucas_cloud_fields: the Aberth iteration did not converge for some polynomial");
    return LM_OK;
}

}  // extern "C"

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
// Laplacian_C-M.py).  Built on the library's own _dev entry points.  The batch streams through the
// device in chunks of 2^16..2^20 polynomials (at least ~8 per call) on three streams (upload of chunk c+1 | K3 + compaction of
// chunk c | download of chunk c-1's cloud points), so with page-locked host buffers the PCIe
// transfers hide behind the solver; the field stages follow on the compute stream, event-timed.
#include "lm_common.cuh"

#include <vector>

namespace {

// first rows handed over as int8 (generalized-Lucas rows are small integers: family_toprow,
// lucas_equipotential_test_v3.py:76-91) are widened on the device: 1 byte instead of 8 per coefficient over PCIe
__global__ void widen_i8_kernel(const signed char* __restrict__ in, long long n, double* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<double>(in[i]);
}

// toprows (float64) or toprows_i8 (int8): exactly one of them is non-NULL when npoly > 0
int32_t cloud_fields_impl(const double* toprows, const signed char* toprows_i8, const int32_t* deg, int64_t npoly, int32_t maxdeg, double tol,
                          double* cloud_re, double* cloud_im, int64_t cap_points, int64_t* n_points,
                          int32_t pot_max_iter, double pot_radius, double* g, int64_t* it,
                          const double* gx, int64_t nx, const double* gy, int64_t ny,
                          double eps, int32_t variant, double h, double* U, double* lapU,
                          lm_cloud_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npoly >= 0 && maxdeg >= 1 && cap_points >= 0, "lm_lucas_cloud_fields: bad sizes");
    LM_REQUIRE(npoly == 0 || ((toprows || toprows_i8) && deg), "lm_lucas_cloud_fields: NULL polynomial buffers");
    LM_REQUIRE(n_points != nullptr, "lm_lucas_cloud_fields: n_points is NULL");
    LM_REQUIRE(nx >= 0 && ny >= 0, "lm_lucas_cloud_fields: negative grid size");
    const bool want_field = (U != nullptr || lapU != nullptr) && nx * ny > 0;
    LM_REQUIRE(!want_field || (gx && gy), "lm_lucas_cloud_fields: grid axes are NULL");
    LM_REQUIRE(lapU == nullptr || h > 0.0, "lm_lucas_cloud_fields: h must be positive for the Laplacian");
    const bool want_pot = (g != nullptr || it != nullptr);
    LM_REQUIRE(!want_pot || (pot_max_iter >= 1 && pot_radius > 0.0), "lm_lucas_cloud_fields: bad potential parameters");
    if (stats) *stats = lm_cloud_stats{};
    *n_points = 0;
    const size_t ncoef = static_cast<size_t>(npoly) * maxdeg;
    void *dtop, *ddeg, *dre, *dim, *dkept, *dstat, *dpx, *dpy;
    if ((rc = lm::ws_get(lm::WS_IN_A, ncoef * sizeof(double), &dtop)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, static_cast<size_t>(npoly) * sizeof(int), &ddeg)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, ncoef * sizeof(double), &dre)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, ncoef * sizeof(double), &dim)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, static_cast<size_t>(npoly) * sizeof(int), &dkept)) != LM_OK) return rc;
    void* dtop8 = nullptr;
    if (toprows_i8 && (rc = lm::ws_get(lm::WS_IN_C, ncoef, &dtop8)) != LM_OK) return rc;
    // the batch streams through the device in chunks: upload of chunk c+1, K3 + compaction of chunk c and the
    // download of chunk c-1's cloud points run on three streams
    // chunk size: at most 2^20 polynomials, but at least ~8 chunks per call so that a rank's share of a sharded batch
    // (1.25e6 polynomials at N = 8) still overlaps its transfers with the solver -- with 2^20-polynomial chunks such a
    // slice was one big chunk plus a stub, and the download of 84 % of its cloud started only after 84 % of its compute
    int64_t CHUNK = int64_t(1) << 20;
    while (CHUNK > (int64_t(1) << 16) && npoly < 8 * CHUNK) CHUNK >>= 1;
    const int nchunks = static_cast<int>(npoly ? (npoly + CHUNK - 1) / CHUNK : 0);
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64 + 16 * static_cast<size_t>(nchunks + 1), &dstat)) != LM_OK) return rc;
    // the cloud has at most sum(deg) <= npoly * maxdeg points: the packed buffers are sized by that bound, and the
    // exact sum (for the report) is accumulated chunk by chunk below, while the device works on the previous chunk
    uint64_t nroots = 0;
    const uint64_t max_points = static_cast<uint64_t>(ncoef);
    const size_t pb = static_cast<size_t>(max_points) * sizeof(double);
    if ((rc = lm::ws_get(lm::WS_CLOUD_A, pb, &dpx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_CLOUD_B, pb, &dpy)) != LM_OK) return rc;

    static cudaStream_t s = nullptr, s_in = nullptr, s_out = nullptr;
    static int s_dev = -1;
    static long long* h_note = nullptr;          // pinned: [c] running cloud size after chunk c, then 2 flags per chunk
    static int h_note_cap = 0;
    lm::register_release_hook([] {
        if (s) { cudaStreamDestroy(s); cudaStreamDestroy(s_in); cudaStreamDestroy(s_out); s = s_in = s_out = nullptr; s_dev = -1; }
        if (h_note) { cudaFreeHost(h_note); h_note = nullptr; h_note_cap = 0; }
    });
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    if (s_dev != dev) {
        if (s) { cudaStreamDestroy(s); cudaStreamDestroy(s_in); cudaStreamDestroy(s_out); }
        LM_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        LM_CUDA_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        LM_CUDA_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        s_dev = dev;
    }
    if (h_note_cap < 2 * (nchunks + 1)) {
        if (h_note) cudaFreeHost(h_note);
        h_note = nullptr; h_note_cap = 0;
        LM_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h_note), sizeof(long long) * 2 * (nchunks + 8), cudaHostAllocDefault));
        h_note_cap = 2 * (nchunks + 8);
    }
    LM_CUDA_TRY(cudaDeviceSynchronize());        // earlier default-stream work may still use the workspaces

    cudaEvent_t ev[6];
    for (auto& e : ev) LM_CUDA_TRY(cudaEventCreate(&e));
    std::vector<cudaEvent_t> ev_in(nchunks, nullptr), ev_done(nchunks, nullptr);
    struct Guard {
        cudaEvent_t* e; std::vector<cudaEvent_t>*a, *b; cudaStream_t s0, s1, s2;
        ~Guard() {
            cudaStreamSynchronize(s0); cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);   // nothing in flight on exit
            for (int k = 0; k < 6; ++k) cudaEventDestroy(e[k]);
            for (cudaEvent_t x : *a) if (x) cudaEventDestroy(x);
            for (cudaEvent_t x : *b) if (x) cudaEventDestroy(x);
        }
    } guard{ev, &ev_in, &ev_done, s, s_in, s_out};

    int64_t* dcount = reinterpret_cast<int64_t*>(static_cast<char*>(dstat));                 // running cloud size
    uint64_t* dwork = reinterpret_cast<uint64_t*>(static_cast<char*>(dstat) + 16);
    int32_t* dflags = reinterpret_cast<int32_t*>(static_cast<char*>(dstat) + 64);            // 2 per chunk (+ padding)
    int32_t* h_flags = reinterpret_cast<int32_t*>(h_note + (nchunks + 4));
    LM_CUDA_TRY(cudaMemsetAsync(dstat, 0, 64 + 16 * static_cast<size_t>(nchunks + 1), s));
    LM_CUDA_TRY(cudaEventRecord(ev[0], s));
    const bool want_cloud = cloud_re || cloud_im;
    int64_t done_points = 0;                     // cloud points already handed to the download stream
    auto download_upto = [&](int c) -> int32_t {  // chunk c is complete: its points can go back
        LM_CUDA_TRY(cudaEventSynchronize(ev_done[c]));
        const int64_t upto = h_note[c];
        if ((want_cloud || want_pot) && upto > cap_points)
            return lm::fail(LM_E_CAP, "lm_lucas_cloud_fields: the cloud has more than %lld points (capacity %lld)",
                            static_cast<long long>(upto), static_cast<long long>(cap_points));
        if (want_cloud && upto > done_points) {
            const size_t off = static_cast<size_t>(done_points), nb = static_cast<size_t>(upto - done_points) * sizeof(double);
            if (cloud_re) LM_CUDA_TRY(cudaMemcpyAsync(cloud_re + off, static_cast<double*>(dpx) + off, nb, cudaMemcpyDeviceToHost, s_out));
            if (cloud_im) LM_CUDA_TRY(cudaMemcpyAsync(cloud_im + off, static_cast<double*>(dpy) + off, nb, cudaMemcpyDeviceToHost, s_out));
        }
        done_points = upto;
        return LM_OK;
    };
    for (int c = 0; c < nchunks; ++c) {
        const int64_t p0 = static_cast<int64_t>(c) * CHUNK, pn = (p0 + CHUNK <= npoly) ? CHUNK : npoly - p0;
        const size_t co = static_cast<size_t>(p0) * maxdeg;
        if (toprows_i8) {
            const long long nc = static_cast<long long>(pn) * maxdeg;
            LM_CUDA_TRY(cudaMemcpyAsync(static_cast<signed char*>(dtop8) + co, toprows_i8 + co, static_cast<size_t>(nc), cudaMemcpyHostToDevice, s_in));
            widen_i8_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, s_in>>>(static_cast<signed char*>(dtop8) + co, nc,
                                                                                       static_cast<double*>(dtop) + co);
            LM_CUDA_TRY(cudaGetLastError());
        } else {
            LM_CUDA_TRY(cudaMemcpyAsync(static_cast<double*>(dtop) + co, toprows + co, static_cast<size_t>(pn) * maxdeg * sizeof(double),
                                        cudaMemcpyHostToDevice, s_in));
        }
        LM_CUDA_TRY(cudaMemcpyAsync(static_cast<int*>(ddeg) + p0, deg + p0, static_cast<size_t>(pn) * sizeof(int), cudaMemcpyHostToDevice, s_in));
        LM_CUDA_TRY(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        LM_CUDA_TRY(cudaEventRecord(ev_in[c], s_in));
        LM_CUDA_TRY(cudaStreamWaitEvent(s, ev_in[c], 0));
        rc = lm_roots_batched_dev(static_cast<double*>(dtop) + co, static_cast<int32_t*>(ddeg) + p0, pn, maxdeg, 1, tol,
                                  static_cast<double*>(dre) + co, static_cast<double*>(dim) + co, static_cast<int32_t*>(dkept) + p0,
                                  nullptr, dflags + 2 * c, s);
        if (rc != LM_OK) return rc;
        rc = lm_cloud_append_dev(static_cast<double*>(dre) + co, static_cast<double*>(dim) + co, static_cast<int32_t*>(dkept) + p0, pn,
                                 maxdeg, static_cast<double*>(dpx), static_cast<double*>(dpy), static_cast<int64_t>(max_points), dcount, s);
        if (rc != LM_OK) return rc;
        LM_CUDA_TRY(cudaMemcpyAsync(&h_note[c], dcount, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        LM_CUDA_TRY(cudaMemcpyAsync(&h_flags[2 * c], dflags + 2 * c, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        LM_CUDA_TRY(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
        LM_CUDA_TRY(cudaEventRecord(ev_done[c], s));
        for (int64_t k = p0; k < p0 + pn; ++k) nroots += static_cast<uint64_t>(deg[k] > 0 ? deg[k] : 0);
        if (c >= 1 && (rc = download_upto(c - 1)) != LM_OK) return rc;
    }
    LM_CUDA_TRY(cudaEventRecord(ev[1], s));
    LM_CUDA_TRY(cudaEventRecord(ev[2], s));
    if (nchunks && (rc = download_upto(nchunks - 1)) != LM_OK) return rc;
    int32_t flags[2] = {0, 0};
    for (int c = 0; c < nchunks; ++c) { flags[0] |= h_flags[2 * c]; flags[1] |= h_flags[2 * c + 1]; }
    const int64_t count = nchunks ? h_note[nchunks - 1] : 0;
    if (flags[1]) return lm::fail(LM_E_INVALID, "lm_lucas_cloud_fields: a degree outside [1, %d]", maxdeg);
    *n_points = count;
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
        stats->launches = launches + nchunks * (7 + (maxdeg > 32) + (maxdeg > 128));   // per chunk: 3 sort + solver launches + 3 compaction
    }
    if (flags[0])
        return lm::fail(LM_E_NOCONV, "lm_lucas_cloud_fields: the Aberth iteration did not converge for some polynomial");
    return LM_OK;
}

}  // namespace

extern "C" {

int32_t lm_lucas_cloud_fields(const double* toprows, const int32_t* deg, int64_t npoly, int32_t maxdeg, double tol,
                              double* cloud_re, double* cloud_im, int64_t cap_points, int64_t* n_points,
                              int32_t pot_max_iter, double pot_radius, double* g, int64_t* it,
                              const double* gx, int64_t nx, const double* gy, int64_t ny,
                              double eps, int32_t variant, double h, double* U, double* lapU,
                              lm_cloud_stats* stats) {
    return cloud_fields_impl(toprows, nullptr, deg, npoly, maxdeg, tol, cloud_re, cloud_im, cap_points, n_points, pot_max_iter,
                             pot_radius, g, it, gx, nx, gy, ny, eps, variant, h, U, lapU, stats);
}

int32_t lm_lucas_cloud_fields_i8(const int8_t* toprows_i8, const int32_t* deg, int64_t npoly, int32_t maxdeg, double tol,
                                 double* cloud_re, double* cloud_im, int64_t cap_points, int64_t* n_points,
                                 int32_t pot_max_iter, double pot_radius, double* g, int64_t* it,
                                 const double* gx, int64_t nx, const double* gy, int64_t ny,
                                 double eps, int32_t variant, double h, double* U, double* lapU,
                                 lm_cloud_stats* stats) {
    return cloud_fields_impl(nullptr, reinterpret_cast<const signed char*>(toprows_i8), deg, npoly, maxdeg, tol, cloud_re, cloud_im,
                             cap_points, n_points, pot_max_iter, pot_radius, g, it, gx, nx, gy, ny, eps, variant, h, U, lapU, stats);
}

}  // extern "C"

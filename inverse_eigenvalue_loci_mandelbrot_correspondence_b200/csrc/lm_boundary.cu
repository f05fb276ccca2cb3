// lm_boundary.cu -- the fused boundary stage of the reference script:
//   xs, ys, Z = compute_grid(...); contour = extract_contour(xs, ys, Z, max_iter, level)
//   (main(), mandelbrot_boundary_sample.py:66-67)
// as ONE host-buffer call: K1 runs in row chunks, the dwell grid stays in HBM for K2 while a copy
// stream returns it to the caller (if asked for), so nothing is uploaded twice and the PCIe
// transfer overlaps the FP64 work and the contour kernels.
#include "lm_common.cuh"

namespace {

int32_t boundary_sample_impl(const double* xs, int64_t nx, const double* ys, int64_t ny,
                             int32_t max_iter, double level, int32_t field_mode, double* field,
                             int32_t* dwell_i32, double* dwell_f64,
                             double* verts, int64_t cap_verts, int64_t* n_verts,
                             int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                             lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(xs && ys, "lm_boundary_sample: xs/ys is NULL");
    LM_REQUIRE(nx >= 0 && ny >= 0, "lm_boundary_sample: negative grid size");
    LM_REQUIRE(max_iter >= 1, "lm_boundary_sample: max_iter must be >= 1 (got %d)", max_iter);
    LM_REQUIRE(n_verts && n_lines && line_offsets, "lm_boundary_sample: NULL output");
    LM_REQUIRE(cap_verts >= 0 && cap_lines >= 0 && (verts || cap_verts == 0), "lm_boundary_sample: bad capacities");
    if (stats) *stats = lm_stats{};
    *n_verts = 0; *n_lines = 0; line_offsets[0] = 0;
    if (nx == 0 || ny == 0) return LM_OK;
    lm::GridHostJob job;
    rc = lm::grid_host_begin(xs, nx, ys, ny, max_iter, 2.0, field_mode, dwell_i32, dwell_f64, field, true, 0, &job);
    if (rc != LM_OK) return rc;
    float k2_ms = 0.f;
    int k2_launches = 0;
    const int32_t rc2 = lm::contour_device_to_host(job.dwell_dev, xs, nx, ys, ny, level, verts, cap_verts, n_verts,
                                                   line_offsets, cap_lines, n_lines, &k2_ms, &k2_launches, job.s_compute);
    lm_stats st{};
    rc = lm::grid_host_finish(&job, &st);
    if (stats) {
        *stats = st;
        stats->launches += k2_launches;
    }
    return rc != LM_OK ? rc : rc2;
}

}  // namespace

extern "C" {

int32_t lm_boundary_sample(const double* xs, int64_t nx, const double* ys, int64_t ny,
                           int32_t max_iter, double level,
                           int32_t* dwell_i32, double* dwell_f64,
                           double* verts, int64_t cap_verts, int64_t* n_verts,
                           int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                           lm_stats* stats) {
    return boundary_sample_impl(xs, nx, ys, ny, max_iter, level, LM_FIELD_NONE, nullptr, dwell_i32, dwell_f64,
                                verts, cap_verts, n_verts, line_offsets, cap_lines, n_lines, stats);
}

int32_t lm_boundary_sample_potential(const double* xs, int64_t nx, const double* ys, int64_t ny,
                                     int32_t max_iter, double level,
                                     int32_t* dwell_i32, double* dwell_f64, double* potential,
                                     double* verts, int64_t cap_verts, int64_t* n_verts,
                                     int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                                     lm_stats* stats) {
    LM_REQUIRE(potential != nullptr, "lm_boundary_sample_potential: potential is NULL");
    return boundary_sample_impl(xs, nx, ys, ny, max_iter, level, LM_FIELD_GREEN, potential, dwell_i32, dwell_f64,
                                verts, cap_verts, n_verts, line_offsets, cap_lines, n_lines, stats);
}

}  // extern "C"

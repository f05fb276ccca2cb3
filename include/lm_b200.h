/*
 * lm_b200.h -- C-ABI of liblm_b200.so, the B200 (sm_100a) implementation of the
 * data-parallel hot path of aortizt/inverse-eigenvalue-loci-mandelbrot-correspondence.
 *
 * The reference is pure Python and has no FFI layer; its boundary is function-level
 * (SURVEY.md section 8b).  Every entry point below names the reference function
 * (file:line, relative to the reference checkout) whose arithmetic it replaces.  The
 * binding a maintainer adds to the reference scripts is a ctypes stub over numpy
 * buffers.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch types.
 *   - every function returns an int32 status: LM_OK (0) or a negative LM_E_* code;
 *     lm_last_error() returns a thread-local, NUL-terminated description of the last
 *     failure on the calling thread.
 *   - "host" entry points take caller-owned, C-contiguous host buffers sized by the
 *     caller; the library never retains them after return.  "_dev" entry points take
 *     device pointers on the current CUDA device and a CUstream/cudaStream_t passed as
 *     void* (NULL = legacy default stream); they enqueue work and do not synchronise.
 *   - the library keeps one context per process (device workspaces, work-queue
 *     counters): calls are NOT re-entrant; the Python host serialises them with a lock.
 *   - there is no CPU fallback: without a usable CUDA device every compute entry point
 *     returns LM_E_NODEV.
 */
#ifndef LM_B200_H
#define LM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_ABI_VERSION 2

/* ---- status codes -------------------------------------------------------------- */
#define LM_OK           0
#define LM_E_INVALID   -1   /* bad argument (NULL, negative size, unknown mode ...)     */
#define LM_E_CUDA      -2   /* CUDA runtime / launch error, see lm_last_error()          */
#define LM_E_CAP       -3   /* caller-provided output capacity too small                 */
#define LM_E_NOMEM     -4   /* host or device allocation failed                          */
#define LM_E_NODEV     -5   /* no CUDA device / wrong architecture                       */
#define LM_E_NOCONV    -6   /* iterative solver did not converge for some item           */
#define LM_E_OVERFLOW  -7   /* the reference would raise OverflowError (2**k, k > 1023)  */

/* ---- escape-time modes (K1) ---------------------------------------------------- */
/* what `field` receives; the dwell outputs are the same in every mode.              */
#define LM_FIELD_NONE          0
/* g = log|z_k| * 2^-k at the first k (1-based) with |z_k|^2 > bailout^2, clamped to 0
 * when negative / non-finite, 0 when no escape.
 * lucas_equipotential_test_v3.py:124-151 (mandelbrot_parameter_potential).          */
#define LM_FIELD_GREEN         1
/* log|z|/2**k, k 0-based at break, test abs(z) > R; ALSO evaluated for orbits that
 * never escape (k = max_iter-1, |z|>0).  Potentials.py:32-47.                        */
#define LM_FIELD_POW2_ALWAYS   2
/* log|z|/(k+1), k 0-based, test abs(z) > R, 0 when no escape.
 * Laplacian_C-M.py:27-43, Iterative_Variogram_Laplacian.py:114-130.                  */
#define LM_FIELD_INV_K         3
/* log|z|/2.0**n, n 1-based first escape (sticky), test abs(z) > R, 0 when no escape;
 * variograms_construct_mandelbrot.py:148-167 (before its 5-point smoothing).         */
#define LM_FIELD_POW2_FIRST    4

/* ---- distance-estimator variants (K1b) ----------------------------------------- */
/* dz0 = 0, dz <- (2 z) dz + 1 then z <- z^2 + c, stop at abs(z) > bailout;
 * d = |z| log|z| / max(|dz|, 1e-16); 0 when no escape.
 * construct_stage1_clean.py:50-58.                                                   */
#define LM_DE_SCALAR           0
/* dz0 = 1, first escape abs(z) > R captures z and dz;
 * d = log(max(|z|,1)) |z| / max(|2 z dz|, eps), non-finite -> 0.
 * variograms_construct_mandelbrot.py:61-88.                                          */
#define LM_DE_FIRST_ESCAPE     1
/* dz0 = 1, z captured at the first escape abs(z) > R but dz taken AFTER all max_iter
 * iterations (it keeps iterating and usually overflows, which makes d = 0);
 * d = log|z| |z| / max(|2 z dz_final|, eps), non-finite -> 0.  This is what
 * tci_construct_mandelbrot.py:21-39 and tci_construct_mandelbrot_v002_fixed.py:35-47 compute
 * (the module gi_assumption_tracker_v3.py loads).                                       */
#define LM_DE_FINAL_DZ         2
/* the same function with numpy's SIMD complex multiply restated (a*b: real = fma(ar, br, -(ai*bi)),
 * imag = fma(ar, bi, ai*br) -- npyv_muladdsub): escape mask and d == 0 pattern bit-identical to what
 * tci_construct_mandelbrot_v002_fixed.py:35-47 returns on an FMA-capable host, so sample_mandelbrot_boundary()
 * (:49-59) and everything the tracker derives from it reproduce exactly.                                     */
#define LM_DE_FINAL_DZ_NUMPY   3

/* ---- log-potential variants (K4a) ---------------------------------------------- */
/* U = (1/N) sum_p log(sqrt(dx^2+dy^2) + eps)         Potentials.py:19-27             */
#define LM_LOGPOT_SUM_SQRT     0
/* U = - sum_p log(sqrt(dx^2+dy^2) + eps)/N per term  Laplacian_C-M.py:16-25          */
#define LM_LOGPOT_NEG_PERTERM  1
/* U = (1/N) sum_p log(hypot(dx,dy) + eps)            Iterative_Variogram_Laplacian.py:102-112 */
#define LM_LOGPOT_SUM_HYPOT    2
/* U = (1/N) sum_p log(1/(|z-p| + eps))               variograms_construct_mandelbrot.py:128-146 */
#define LM_LOGPOT_LOG_INV      3

/* ---- per-pair weight of the binned pair statistics (lm_pair_histogram) ---------- */
#define LM_PAIR_W_NONE          0   /* counts only (pair_correlation, ripley_K)                     */
#define LM_PAIR_W_VALUE_SQDIFF  1   /* (value_i - value_j)^2 (empirical_variogram_field)            */
#define LM_PAIR_W_DIST_SQ       2   /* d_ij^2 (empirical_variogram_coords)                          */

/* ---- plain-old-data structs ---------------------------------------------------- */
typedef struct lm_device_info {
    int32_t  device;            /* CUDA ordinal                                       */
    int32_t  cc_major, cc_minor;
    int32_t  sm_count;
    int32_t  clock_khz;         /* max SM clock                                       */
    int32_t  l2_bytes;
    uint64_t total_mem_bytes;
    char     name[128];
} lm_device_info;

typedef struct lm_stats {
    uint64_t work_units;        /* K1: pixel-iterations = sum min(dwell+1, max_iter)  */
    uint64_t items;             /* pixels / points / polynomials / cells processed    */
    float    kernel_ms;         /* device time of the dominant kernel (host entry
                                   points only; 0 for _dev entry points)              */
    int32_t  launches;          /* kernels launched by this call                      */
} lm_stats;

/* per-stage report of lm_lucas_cloud_fields */
typedef struct lm_cloud_stats {
    uint64_t n_roots;           /* sum of the degrees                                   */
    uint64_t n_points;          /* cloud size (roots with |lambda| > tol)               */
    uint64_t potential_work;    /* K1d iterations performed                             */
    uint64_t pairs;             /* K4a (cell, point) pairs                              */
    float    roots_ms, compact_ms, potential_ms, logpot_ms, stencil_ms;   /* device time */
    int32_t  launches;
} lm_cloud_stats;

/* ---- context, memory ----------------------------------------------------------- */
int32_t     lm_abi_version(void);
const char* lm_last_error(void);
int32_t     lm_device_count(void);                 /* >=0, or LM_E_* */
int32_t     lm_set_device(int32_t device);
int32_t     lm_get_device_info(lm_device_info* out);
int32_t     lm_device_synchronize(void);
int32_t     lm_release_workspace(void);            /* frees cached device buffers */

void*   lm_host_alloc(size_t bytes);               /* pinned; NULL on failure */
int32_t lm_host_free(void* p);
void*   lm_dev_alloc(size_t bytes);                /* NULL on failure */
int32_t lm_dev_free(void* p);
int32_t lm_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int32_t lm_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int32_t lm_memcpy_d2d(void* dst_dev, const void* src_dev, size_t bytes, void* stream);
int32_t lm_stream_synchronize(void* stream);

/* ---- K1: escape-time grid ------------------------------------------------------ */
/*
 * Replaces compute_grid + mandelbrot_dwell, mandelbrot_boundary_sample.py:22-39:
 *   dwell[j*nx+i] = first n in [0,max_iter) with |z_{n+1}|^2 > bailout^2 for
 *   c = xs[i] + i*ys[j], z_0 = 0, z <- z*z + c, else max_iter.
 * The recurrence is evaluated in IEEE binary64 without contraction in the reference's
 * operation order, so dwell is bit-exact.
 * Any of dwell_i32 / dwell_f64 / field may be NULL (not produced).  `field_mode` is
 * one of LM_FIELD_*; for LM_FIELD_GREEN the escape test is zr^2+zi^2 > bailout^2, for
 * the other field modes it is hypot(zr,zi) > bailout, as in the reference functions.
 * work_units returns the exact pixel-iteration count.
 */
int32_t lm_escape_grid_f64(const double* xs, int64_t nx, const double* ys, int64_t ny,
                           int32_t max_iter, double bailout, int32_t field_mode,
                           int32_t* dwell_i32, double* dwell_f64, double* field,
                           lm_stats* stats);
int32_t lm_escape_grid_f64_dev(const double* xs, int64_t nx, const double* ys, int64_t ny,
                               int32_t max_iter, double bailout, int32_t field_mode,
                               int32_t* dwell_i32, double* dwell_f64, double* field,
                               uint64_t* work_units_dev /* 1 counter, may be NULL */,
                               void* stream);

/*
 * Multi-GPU building block (one process per GPU, each owning a contiguous block of rows): K1 over the
 * ny rows of this shard exactly like lm_escape_grid_f64 (chunked, the dwell block returned to
 * dwell_i32 -- may be NULL -- by a copy stream that overlaps the compute), but the int32 block ALSO
 * stays in HBM with room for `halo_rows` more rows behind it.  *dwell_dev_out is that device block
 * ([ny + halo_rows] x nx, library owned, valid until the next grid call on this device): the caller
 * exchanges shard-edge rows over NCCL straight from / into it (lm_memcpy_d2d) and then runs
 * lm_contour_classify_dev on it, so nothing is uploaded twice.  With potential != NULL the same pass
 * also produces the smooth potential (LM_FIELD_GREEN) of the shard's rows: returned to the host
 * buffer and kept in HBM (*potential_dev_out, [ny] x nx) for the all-gather of the final field.
 * row_cost (may be NULL): ny estimates of the rows' relative cost -- the profile the shards were cut with; the
 * row chunks are then computed cheapest first, so that the chunks finishing last are the ones whose copy to
 * the host hides behind their own compute.
 */
int32_t lm_shard_escape(const double* xs, int64_t nx, const double* ys, int64_t ny, int32_t max_iter,
                        int32_t* dwell_i32, double* potential, int64_t halo_rows, const double* row_cost,
                        int32_t** dwell_dev_out, double** potential_dev_out, lm_stats* stats);

/* Optional single-precision variant of the dwell grid (BASELINE.json north_star, piece 1; no reference
 * counterpart): the SAME persistent kernel -- work stealing, lane refill, blind blocks, staged 128-bit stores --
 * instantiated in binary32 with unfused round-to-nearest operations.  Validated against the fp64 kernel by a stated
 * dwell-mismatch / interior-mask tolerance (tests/test_gpu_escape.py::test_f32_*), never bit-exact.          */
int32_t lm_escape_grid_f32(const double* xs, int64_t nx, const double* ys, int64_t ny,
                           int32_t max_iter, double bailout,
                           int32_t* dwell_i32, lm_stats* stats);
int32_t lm_escape_grid_f32_dev(const double* xs, int64_t nx, const double* ys, int64_t ny,
                               int32_t max_iter, double bailout, int32_t* dwell_i32,
                               uint64_t* work_units_dev /* 1 counter, may be NULL */, void* stream);

/* ---- K1d: escape-time at an arbitrary point list -------------------------------- */
/*
 * Replaces batch_potential / mandelbrot_parameter_potential,
 * lucas_equipotential_test_v3.py:124-162.  it[k] in [1,max_iter]; g, phi as there;
 * non-escaping points give (0, max_iter, nan+nanj).  Any output may be NULL.
 */
int32_t lm_escape_points_f64(const double* c_re, const double* c_im, int64_t n,
                             int32_t max_iter, double escape_radius,
                             double* g, int64_t* it, double* phi_re, double* phi_im,
                             lm_stats* stats);

/* device-resident variant: all pointers are device pointers (outputs may be NULL), work_units_dev
 * (1 counter, may be NULL) receives the iteration count; enqueues on `stream`, never syncs.  */
int32_t lm_escape_points_f64_dev(const double* c_re_dev, const double* c_im_dev, int64_t n,
                                 int32_t max_iter, double escape_radius,
                                 double* g_dev, int64_t* it_dev, double* phi_re_dev, double* phi_im_dev,
                                 uint64_t* work_units_dev, void* stream);

/* ---- K1b: distance-estimator grid ---------------------------------------------- */
/* construct_stage1_clean.py:50-58 (LM_DE_SCALAR), variograms_construct_mandelbrot.py:61-88
 * (LM_DE_FIRST_ESCAPE).  escaped[] (uint8, may be NULL) is the first-escape mask.    */
int32_t lm_distance_grid_f64(const double* xs, int64_t nx, const double* ys, int64_t ny,
                             int32_t max_iter, double bailout, double eps, int32_t variant,
                             double* dist, uint8_t* escaped, lm_stats* stats);

/* ---- K2: level set of the dwell field ------------------------------------------ */
/*
 * Replaces plt.contour(xs, ys, Z, levels=[level]) as used by extract_contour,
 * mandelbrot_boundary_sample.py:41-54 / mandelbrot_boundary_sample_spyder.py:35-43
 * (contourpy "mpl2014" line semantics).
 * Input is the int32 dwell grid (host pointer) or, for _dev, a device pointer.
 * Output: all contour lines in matplotlib's order; line l occupies vertices
 * [line_offsets[l], line_offsets[l+1]) of verts (x,y interleaved); closed loops repeat
 * their first vertex.  Returns LM_E_CAP (with *n_verts / *n_lines set to the required
 * sizes) when a capacity is too small.
 */
int32_t lm_contour_level(const int32_t* dwell, const double* xs, int64_t nx,
                         const double* ys, int64_t ny, double level,
                         double* verts, int64_t cap_verts, int64_t* n_verts,
                         int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                         lm_stats* stats);
int32_t lm_contour_level_dev(const int32_t* dwell_dev, const double* xs_host, int64_t nx,
                             const double* ys_host, int64_t ny, double level,
                             double* verts, int64_t cap_verts, int64_t* n_verts,
                             int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                             lm_stats* stats);

/* After lm_contour_level[_dev] / lm_boundary_sample / lm_contour_link[_dev] returned LM_E_CAP the
 * linked polylines stay on the device; this copies them out (nothing is recomputed).  Valid until the
 * next contour call on the same device.                                                        */
int32_t lm_contour_fetch_last(double* verts, int64_t cap_verts, int64_t* n_verts,
                              int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines);

/*
 * The fused boundary stage of the reference script's main()
 * (xs, ys, Z = compute_grid(...); contour = extract_contour(xs, ys, Z, max_iter, level),
 * mandelbrot_boundary_sample.py:66-67) as one host-buffer call: the dwell grid stays in HBM
 * between K1 and K2 and is returned to dwell_i32 / dwell_f64 (either may be NULL) by a copy
 * stream that overlaps the compute.  `level` is the absolute level (level_frac * max_iter).
 * Lines come back as in lm_contour_level; LM_E_CAP reports the required capacities.
 */
int32_t lm_boundary_sample(const double* xs, int64_t nx, const double* ys, int64_t ny,
                           int32_t max_iter, double level,
                           int32_t* dwell_i32, double* dwell_f64,
                           double* verts, int64_t cap_verts, int64_t* n_verts,
                           int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                           lm_stats* stats);

/* The same with the smooth (Green) potential g = log|z_k| * 2^-k of the escaping iterate (0 inside; the a-7
 * definition of lucas_equipotential_test_v3.py:140-149 on the grid, BASELINE.json config 2) produced by the
 * same K1 pass and returned to potential[ny*nx] by the copy stream.                                       */
int32_t lm_boundary_sample_potential(const double* xs, int64_t nx, const double* ys, int64_t ny,
                                     int32_t max_iter, double level,
                                     int32_t* dwell_i32, double* dwell_f64, double* potential,
                                     double* verts, int64_t cap_verts, int64_t* n_verts,
                                     int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                                     lm_stats* stats);

/*
 * Multi-GPU building block: classify the quads of rows [0, ny-1) of a dwell block on the
 * device (ny rows including one halo row; ys_host holds the block's ny coordinates) and
 * return the compacted crossing-quad records, in raster order, to the host.  Records of
 * consecutive row blocks are concatenated and chained by lm_contour_link[_dev].
 * A record is 8 x int64: {quad = (row_offset + j)*nx + i, SW | SE<<32, NW | NE<<32 (corner
 * dwell values), meta, exit vertex of segment 0 (x, y as binary64), exit vertex of
 * segment 1 (saddle quads)}; meta is described in csrc/lm_contour.cu.  A page-locked `records`
 * buffer (lm_host_alloc) is filled straight from the device, a pageable one through a staging copy.
 */
int32_t lm_contour_classify_dev(const int32_t* dwell_dev, const double* xs_host, int64_t nx,
                                const double* ys_host, int64_t ny, int64_t row_offset, double level,
                                int64_t* records, int64_t cap_records, int64_t* n_records,
                                void* stream);
/* The same with the records left on the device, in the caller's device buffer `records_dev` (for the
 * NCCL gather of the shards' records: nothing bounces through the host).  LM_E_CAP with *n_records set
 * when the buffer is too small.                                                                   */
int32_t lm_contour_records_dev(const int32_t* dwell_dev, const double* xs_host, int64_t nx,
                               const double* ys_host, int64_t ny, int64_t row_offset, double level,
                               int64_t* records_dev, int64_t cap_records, int64_t* n_records,
                               void* stream);
/*
 * The ordered polylines of plt.contour (the line assembly of contourpy's mpl2014: lines(),
 * get_start_edge, follow_interior; extract_contour, mandelbrot_boundary_sample.py:41-54) from
 * raster-ordered records -- of one block or of consecutive row blocks concatenated -- ON THE DEVICE:
 * successor table, pointer-jumping list ranking, scan over the line leaders, scatter of the vertices
 * (csrc/lm_contour_link.cu).  xs, ys: the FULL grid coordinates (host).  lm_contour_link_dev takes the
 * records in device memory, lm_contour_link uploads host records first; both need a device.
 * LM_E_INVALID for records that are not in raster order or miss a neighbour.  Output and LM_E_CAP
 * protocol as in lm_contour_level.
 */
int32_t lm_contour_link_dev(const int64_t* records_dev, int64_t n_records,
                            const double* xs, int64_t nx, const double* ys, int64_t ny,
                            double level,
                            double* verts, int64_t cap_verts, int64_t* n_verts,
                            int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                            void* stream);
int32_t lm_contour_link(const int64_t* records, int64_t n_records,
                        const double* xs, int64_t nx, const double* ys, int64_t ny,
                        double level,
                        double* verts, int64_t cap_verts, int64_t* n_verts,
                        int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines);

/* ---- K3: batched roots of generalized-Lucas characteristic polynomials --------- */
/*
 * Replaces np.linalg.eigvals(companion(top row)) in compute_inverse_eigenvalues[_family],
 * lucas_equipotential_test_v3.py:58-118, construct_points in tci_construct_mandelbrot.py:11-19,
 * tci_construct_mandelbrot_v002_fixed.py:27-33, construct_stage1_clean.py:34-48,
 * variograms_construct_mandelbrot.py:48-56, lucas_to_cardioid_v18...py:83-94.
 * Polynomial k is  x^d - a_1 x^(d-1) - ... - a_d  with d = deg[k] and
 * a_j = toprows[k*maxdeg + j-1] (zero padded to maxdeg).
 * Output slot k*maxdeg + r holds root r of polynomial k (r < deg[k]); with invert != 0
 * the value is 1/lambda and roots with |lambda| <= tol are dropped (slots compacted to
 * the front; n_kept[k] = number of valid slots).  Unused slots are NaN.
 * iters (may be NULL) receives the Aberth sweep count per polynomial.
 * Roots are unordered (the reference's order is LAPACK's); parity is after sorting.
 */
int32_t lm_roots_batched(const double* toprows, const int32_t* deg, int64_t npoly,
                         int32_t maxdeg, int32_t invert, double tol,
                         double* out_re, double* out_im, int32_t* n_kept, int32_t* iters,
                         lm_stats* stats);

/* Device-resident variant: every pointer is a device pointer; the batch is sorted by degree on
 * the device and the solver launches read their ranges from device memory, so the call only
 * enqueues on `stream`.  status_dev (2 x int32, may be NULL): [0] != 0 when some polynomial did
 * not converge, [1] != 0 when some deg[k] is outside [1, maxdeg] (those are skipped).          */
int32_t lm_roots_batched_dev(const double* toprows_dev, const int32_t* deg_dev, int64_t npoly,
                             int32_t maxdeg, int32_t invert, double tol,
                             double* out_re_dev, double* out_im_dev, int32_t* n_kept_dev,
                             int32_t* iters_dev, int32_t* status_dev, void* stream);
/* The cloud of the reference (pts.extend(1/vals) per n, lucas_equipotential_test_v3.py:98-99):
 * the first n_kept[k] slots of every polynomial, concatenated in polynomial order, written to
 * px_dev / py_dev (capacity cap_points; excess is dropped); *n_points_dev = full count.       */
int32_t lm_cloud_compact_dev(const double* re_dev, const double* im_dev, const int32_t* n_kept_dev,
                             int64_t npoly, int32_t maxdeg, double* px_dev, double* py_dev,
                             int64_t cap_points, int64_t* n_points_dev, void* stream);
/* Chunked form: appends this chunk's points behind the *n_points_inout_dev points already in px_dev / py_dev
 * and adds their number to it (lets a batch stream through the device chunk by chunk).                        */
int32_t lm_cloud_append_dev(const double* re_dev, const double* im_dev, const int32_t* n_kept_dev,
                            int64_t npoly, int32_t maxdeg, double* px_dev, double* py_dev,
                            int64_t cap_points, int64_t* n_points_inout_dev, void* stream);

/*
 * The Lucas-Loci field stage (BASELINE.json config 5) as one host-buffer call:
 *   K3 roots -> cloud of 1/lambda (|lambda| > tol, polynomial order)
 *   -> K1d batch_potential at every cloud point (lucas_equipotential_test_v3.py:153-162; skipped
 *      when g and it are NULL) -> K4a log-potential of the cloud on the grid (variant/eps as
 *      lm_log_potential) -> K4 periodic Laplacian of that field (Laplacian_C-M.py:49-59).
 * cloud_re/cloud_im/g/it have capacity cap_points (any may be NULL); U / lapU are [ny*nx]
 * (either may be NULL).  *n_points receives the cloud size; LM_E_CAP if it exceeds
 * cap_points while a per-point output was requested.
 */
int32_t lm_lucas_cloud_fields(const double* toprows, const int32_t* deg, int64_t npoly, int32_t maxdeg,
                              double tol, double* cloud_re, double* cloud_im, int64_t cap_points,
                              int64_t* n_points, int32_t pot_max_iter, double pot_radius,
                              double* g, int64_t* it,
                              const double* gx, int64_t nx, const double* gy, int64_t ny,
                              double eps, int32_t variant, double h, double* U, double* lapU,
                              lm_cloud_stats* stats);
/* Same call with the first rows as int8 (generalized-Lucas rows are small integers -- family_toprow,
 * lucas_equipotential_test_v3.py:76-91 -- and int8 -> binary64 is exact): 1 byte per coefficient over PCIe
 * instead of 8, widened on the device; results are bit-identical to the float64 call.                       */
int32_t lm_lucas_cloud_fields_i8(const int8_t* toprows_i8, const int32_t* deg, int64_t npoly, int32_t maxdeg, double tol,
                                 double* cloud_re, double* cloud_im, int64_t cap_points, int64_t* n_points,
                                 int32_t pot_max_iter, double pot_radius, double* g, int64_t* it,
                                 const double* gx, int64_t nx, const double* gy, int64_t ny,
                                 double eps, int32_t variant, double h, double* U, double* lapU,
                                 lm_cloud_stats* stats);

/* ---- K4: 5-point stencils ------------------------------------------------------ */
/* lap = (((((-4 U) + U[j-1]) + U[j+1]) + U[:,i-1]) + U[:,i+1]) / (h*h), periodic wrap.
 * Laplacian_C-M.py:49-59, Iterative_Variogram_Laplacian.py:132-136.  Bit-exact.      */
int32_t lm_laplacian5_periodic(const double* U, int64_t ny, int64_t nx, double h,
                               double* out, lm_stats* stats);
int32_t lm_laplacian5_periodic_dev(const double* U, int64_t ny, int64_t nx, double h,
                                   double* out, void* stream);
/* interior 5-point average ((((c+up)+down)+left)+right)/5, border copied.
 * variograms_construct_mandelbrot.py:169-173.  Bit-exact.                            */
int32_t lm_smooth5_interior(const double* g, int64_t ny, int64_t nx, double* out,
                            lm_stats* stats);
int32_t lm_smooth5_interior_dev(const double* g, int64_t ny, int64_t nx, double* out,
                                void* stream);

/* ---- K4a: log-potential of a point cloud on a grid ------------------------------ */
/* Potentials.py:19-27, Laplacian_C-M.py:16-25, Iterative_Variogram_Laplacian.py:102-112,
 * variograms_construct_mandelbrot.py:128-146 (variant = LM_LOGPOT_*).
 * All four are +-(1/N) sum_p log(|z - p| + eps); the reference's variants differ from each
 * other in the last ulp, parity is 1e-12 (the sum is re-associated, see csrc/lm_logpot.cu).   */
int32_t lm_log_potential(const double* px, const double* py, int64_t npts,
                         const double* gx, int64_t nx, const double* gy, int64_t ny,
                         double eps, int32_t variant, double* U, lm_stats* stats);
/* Multi-GPU building blocks (device pointers, enqueue only): raw per-cell sums
 * sums[cell] = sum_p log(|z_cell - p| + eps) over THIS rank's slice of the cloud; after an
 * all-reduce(sum) of `sums` across ranks, `finish` applies the variant's +-1/N_total.        */
int32_t lm_log_potential_sums_dev(const double* px_dev, const double* py_dev, int64_t npts,
                                  const double* gx_dev, int64_t nx, const double* gy_dev, int64_t ny,
                                  double eps, int32_t variant, double* sums_dev, void* stream);
int32_t lm_log_potential_finish_dev(const double* sums_dev, int64_t ncells, int64_t n_total_points,
                                    int32_t variant, double* U_dev, void* stream);

/* ---- nearest-neighbour matching of two point sets (tracker module) -- */
/* index[k] = first j minimising |x_k - y_j| (sqrt(dx*dx + dy*dy), unfused): what
 * argmax(exp(-cdist(X, Y)/const), axis=1) selects in entropic_ot_alignment,
 * tci_construct_mandelbrot_v002_fixed.py:62-71.  distance (may be NULL) receives the minimum.   */
int32_t lm_nearest_match(const double* x_re, const double* x_im, int64_t n,
                         const double* y_re, const double* y_im, int64_t m,
                         int64_t* index, double* distance, lm_stats* stats);

/* ---- dense boundary-integral (Nystrom) sums, SURVEY 8f-2 ------------------------ */
/* out[m] = sum_n weight[n] * log(|z_m - node_n| + eps): the O(M*N) part of g_real,
 * lucas_to_cardioid_v40_reference.py:240-257 (log(abs(z[:,None]-bdy[None,:]) + 1e-300) @ (sigma*ds)).  */
int32_t lm_weighted_log_sum(const double* z_re, const double* z_im, int64_t M,
                            const double* node_re, const double* node_im, const double* weight, int64_t N,
                            double eps, double* out, lm_stats* stats);
/* out[m] = sum_n weight[n] / dz_mn with dz_mn = z_m - node_n replaced by dz_eps + 0j where |dz_mn| < dz_eps:
 * the integral term of dPhi, lucas_to_cardioid_v40_reference.py:201-211.                                 */
int32_t lm_weighted_cauchy_sum(const double* z_re, const double* z_im, int64_t M,
                               const double* node_re, const double* node_im, const double* weight, int64_t N,
                               double dz_eps, double* out_re, double* out_im, lm_stats* stats);

/* ---- local-polynomial curvature of an ordered boundary ------------- */
/* compute_curvature_localpoly(P, neighbors, closed, stride=1), boundary_curvature_localpoly.py:133-184:
 * per point a least-squares quadratic in arclength over the window i-neighbors..i+neighbors (wrapped when
 * closed != 0, clamped otherwise); kappa_signed = (x'y'' - y'x'') / (|(x',y')| + 1e-16)^3.  x, y: the n
 * ordered points (the two columns of <prefix>_boundary.csv).  xprime..y2 may be NULL.                   */
int32_t lm_curvature_localpoly(const double* x, const double* y, int64_t n, int32_t neighbors, int32_t closed,
                               double* kappa, double* kappa_signed, double* speed,
                               double* xprime, double* yprime, double* x2, double* y2, lm_stats* stats);

/* ---- binned statistics over all point pairs ----------------------- */
/* counts[k] = #{ i < j : lo[k] <= d_ij < hi[k] },  sums[k] = sum of w_ij over those pairs, with
 * d_ij = sqrt(dx*dx + dy*dy) in unfused binary64 (= scipy pdist / distance_matrix / np.linalg.norm(axis=1)) and
 * w_ij chosen by weight_mode (LM_PAIR_W_*).  One call is the O(N^2) part of
 *   empirical_variogram_field, empirical_variogram_coords   Variogram-Mandelbrot-Construct.py:106-152
 *       (lo = bins[:-1], hi = bins[1:] of np.linspace(0, max_dist, nbins+1); gamma = 0.5*sums/counts)
 *   empirical_variogram_from_field_locs                     Iterative_Variogram_Laplacian.py:53-86
 *   pair_correlation (lo = r_vals, hi = r_vals + dr), ripley_K (cumulated counts)   spatial_stats_phase2.py:9-47
 * without materialising the N(N-1)/2 distances.  Edges: lo strictly increasing, hi[k] <= lo[k+2] (a shell may
 * overlap its neighbour, as r+dr can exceed the next r by an ulp; such a pair is counted in both, like the
 * reference's masks do); hi may be +inf.  1 <= nbins <= 2048.  value is read only for LM_PAIR_W_VALUE_SQDIFF;
 * sums may be NULL for LM_PAIR_W_NONE.  Coordinates must be finite.  n < 2 gives all-zero outputs.            */
int32_t lm_pair_histogram(const double* x, const double* y, const double* value, int64_t n,
                          const double* lo, const double* hi, int32_t nbins, int32_t weight_mode,
                          uint64_t* counts, double* sums, lm_stats* stats);
/* The (v_i - w_j)^2 of the pairs (i in A, j in B) whose distance lies in [lo, hi), in row-major order of (i, j):
 * dV2[np.where((D >= lo) & (D < hi) & ~diag)] of sample_semivariogram / sample_cross_semivariogram,
 * variograms_construct_mandelbrot.py:178-315 -- needed for the one block per bin in which the reference draws a random
 * subset of exactly this list (max_pairs_per_bin); all other blocks only need counts and sums (lm_pair_histogram).
 * skip_diagonal != 0 drops the pairs with i == j (a block of a set against itself).  LM_E_CAP with *n_values set when
 * `values` is too small.                                                                                       */
int32_t lm_pair_select_sqdiff(const double* xa, const double* ya, const double* va, int64_t na,
                              const double* xb, const double* yb, const double* vb, int64_t nb,
                              double lo, double hi, int32_t skip_diagonal,
                              double* values, int64_t cap_values, int64_t* n_values, lm_stats* stats);
/* *dmax = max_{i<j} d_ij: the D.max() behind the default max_dist = 0.5 * D.max()
 * (Variogram-Mandelbrot-Construct.py:118-119, Iterative_Variogram_Laplacian.py:60-61).  0 for n < 2.          */
int32_t lm_pair_max_distance(const double* x, const double* y, int64_t n, double* dmax, lm_stats* stats);

/* ---- the tracker's density stage: histogram, blur, divergences, GI flow -------- */
/* H[ix*nby + iy] = np.histogram2d(x, y, bins=(nbx, nby), range=...)[0]: xedges[nbx+1] / yedges[nby+1] are the caller's
 * np.linspace edges; a sample is in bin k when edges[k] <= v < edges[k+1] (np.searchsorted side="right"), the last
 * edge closed, everything else (NaN included) dropped.  gi_assumption_tracker_v3.py:110-114,
 * tci_construct_mandelbrot_v002_fixed.py:78-82 (to_prob).                                                        */
int32_t lm_histogram2d(const double* x, const double* y, int64_t n, const double* xedges, int32_t nbx,
                       const double* yedges, int32_t nby, double* H, lm_stats* stats);
/* mollified_histogram, gi_assumption_tracker_v3.py:109-127, in one device pass:
 * histogram2d -> max(H, eps) -> [radius > 0: scipy.ndimage.gaussian_filter(mode="nearest") with the given normalised
 * symmetric weights[2*radius+1] (axis 0, then axis 1) -> max(H, eps)] -> H / H.sum() (numpy's pairwise order).
 * Bit-identical to the numpy / scipy chain.                                                                      */
int32_t lm_mollified_histogram(const double* x, const double* y, int64_t n, const double* xedges, int32_t nbx,
                               const double* yedges, int32_t nby, double eps, const double* weights, int32_t radius,
                               double* P, lm_stats* stats);
/* scipy.ndimage.gaussian_filter(in[n0, n1], mode="nearest") given its 1-D weights (correlate1d, symmetric path).  */
int32_t lm_gaussian_filter_nearest(const double* in, int64_t n0, int64_t n1, const double* weights, int32_t radius,
                                   double* out, lm_stats* stats);
/* *out = np.sum(a) for a contiguous float64 array, in numpy's pairwise summation order (bit-identical).            */
int32_t lm_sum_pairwise(const double* a, int64_t n, double* out, lm_stats* stats);
/* One pass over two densities: sum|p-q| (tv_distance = 0.5 * it), sum min(p,q) (overlap_mass) and
 * KL(p, q) = sum p_ (log p_ - log q_) with p_ = max(p, eps), q_ = max(q, eps).
 * gi_assumption_tracker_v3.py:91-96, tci_construct_mandelbrot_v002_fixed.py:84-86.  Outputs may be NULL.          */
int32_t lm_density_compare(const double* p, const double* q, int64_t n, double eps,
                           double* sum_abs_diff, double* sum_min, double* kl_pq, lm_stats* stats);
/* gi_flow_to_threshold (fixed_T == 0) / gi_flow_fixed_T (fixed_T != 0, T = max_steps), gi_assumption_tracker_v3.py:130-151,
 * with the stock module's KL: X <- (1-alpha) X + alpha P (unfused), KL(P, X) after every sweep, stop at the first
 * t >= max(min_steps, 1) with KL <= kl_threshold.  X_out[n] = X_T (bit-identical to numpy), *steps_out = T,
 * *kl_initial = KL(P, X0), *kl_final = KL(P, X_T); kl_history (may be NULL) receives max_steps+1 slots, filled 0..T. */
int32_t lm_gi_flow(const double* P, const double* X0, int64_t n, double alpha, double eps,
                   int32_t max_steps, int32_t min_steps, double kl_threshold, int32_t fixed_T,
                   double* X_out, int32_t* steps_out, double* kl_initial, double* kl_final,
                   double* kl_history, lm_stats* stats);

/* ---- alpha-shape edge filter of the boundary consumers ------------- */
/* alpha_shape_edges(P, alpha), construct_boundary_alpha.py:45-82, given the Delaunay simplices[ntri*3] of the points
 * (the reference calls scipy.spatial.Delaunay(P).simplices): keep[t] = circumradius(t) < 1/alpha
 * (R = abc / (4 sqrt(max(s(s-a)(s-b)(s-c), 0)) + 1e-16), inf for a degenerate triangle), boundary edges = edges used by
 * exactly one kept triangle, as (min, max) vertex pairs in the order the reference's dict yields them (first occurrence
 * walking the kept triangles, edges (t0,t1), (t1,t2), (t2,t0)).  keep / radius may be NULL.  LM_E_CAP (with *n_edges set)
 * when more than cap_edges edges exist.                                                                          */
int32_t lm_alpha_shape_edges(const double* x, const double* y, int64_t npts, const int32_t* simplices, int64_t ntri, double alpha,
                             uint8_t* keep, double* radius, int32_t* edges, int64_t cap_edges, int64_t* n_edges, lm_stats* stats);

/* ---- measurement probes -------------------------------------------------------- */
/* Dependent-free DFMA loop on every SM: FP64 peak (TFLOP/s, 2 flops per DFMA) and a
 * DMUL/DADD-only variant (the unfused mix K1 needs).  Used by bench.py for the
 * roofline denominator because MEASURED_PEAKS.json has no FP64 entry.                */
int32_t lm_probe_fp64_peak(int32_t iters, double* dfma_tflops, double* dmul_dadd_tinstr);
/* Cycles per dependent FP64 instruction (one warp, one chain): DFMA, DADD, DMUL. */
int32_t lm_probe_fp64_latency(double* dfma_cycles, double* dadd_cycles, double* dmul_cycles);
/* The bare K1 recurrence (6 FP64 instr / iteration, interior points, no escape test) at a
 * given occupancy: the practical ceiling of the escape kernel's blind path, in
 * G pixel-iterations/s.  variant 0: recurrence only; 1: + calm tracking and a vote per 16. */
int32_t lm_probe_k1_loop(int32_t variant, int32_t warps_per_sm, int32_t blocks16,
                         double* gpixel_iters_per_s);
/* STREAM-style device copy bandwidth in GB/s (read+write bytes). */
int32_t lm_probe_hbm_copy(size_t bytes, int32_t reps, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* LM_B200_H */

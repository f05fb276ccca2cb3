// lm_stencil.cu -- K4: 5-point stencils on fp64 fields (HBM bound, 16 B/pixel algorithmic).
//
//   laplacian(U, h)   Laplacian_C-M.py:49-59, laplacian_fd Iterative_Variogram_Laplacian.py:132-136
//       (((((-4 U) + U[j-1,i]) + U[j+1,i]) + U[j,i-1]) + U[j,i+1]) / (h*h), periodic wrap (np.roll)
//   5-point interior average   variograms_construct_mandelbrot.py:169-173
//       ((((g[j,i] + g[j-1,i]) + g[j+1,i]) + g[j,i-1]) + g[j,i+1]) / 5.0 on the interior, border copied
//
// Both keep the reference's left-to-right association with __dmul_rn/__dadd_rn/__ddiv_rn, so
// results are bit-exact.  Layout: each thread owns two adjacent columns (one 128-bit load and
// one 128-bit store per row), a CTA of 128 threads covers 256 columns and marches down
// ROWS_PER_CTA rows with the three live rows in registers; left/right neighbours come from
// warp shuffles, only the two edge lanes of a warp issue an extra (L1/L2-resident) scalar
// load.  DRAM traffic is (ROWS_PER_CTA+2)/ROWS_PER_CTA reads + 1 write per pixel.
#include "lm_common.cuh"

namespace {

constexpr int ST_THREADS = 128;
constexpr int ROWS_PER_CTA = 32;
constexpr unsigned FULL = 0xffffffffu;

enum { OP_LAPLACIAN = 0, OP_SMOOTH = 1 };

template <int OP>
__device__ __forceinline__ double combine(double c, double up, double dn, double lf, double rt, double h2) {
    if (OP == OP_LAPLACIAN) {
        double s = __dmul_rn(-4.0, c);
        s = __dadd_rn(s, up);
        s = __dadd_rn(s, dn);
        s = __dadd_rn(s, lf);
        s = __dadd_rn(s, rt);
        return __ddiv_rn(s, h2);
    } else {
        double s = __dadd_rn(c, up);
        s = __dadd_rn(s, dn);
        s = __dadd_rn(s, lf);
        s = __dadd_rn(s, rt);
        return __ddiv_rn(s, 5.0);
    }
}

// vectorised kernel: nx even, pointers 16-byte aligned
#ifndef LM_K4_PF_LAP
#define LM_K4_PF_LAP 2
#endif
#ifndef LM_K4_PF_SMOOTH
#define LM_K4_PF_SMOOTH 2
#endif
template <int OP>
__global__ void __launch_bounds__(ST_THREADS) stencil_vec_kernel(const double* __restrict__ in,
                                                                 double* __restrict__ out,
                                                                 long long ny, long long nx, double h2) {
    const int lane = threadIdx.x & 31;
    const long long c0 = (static_cast<long long>(blockIdx.x) * ST_THREADS + threadIdx.x) * 2;
    const bool live = c0 < nx;                 // nx even => c0+1 < nx as well
    const long long r0 = static_cast<long long>(blockIdx.y) * ROWS_PER_CTA;
    if (r0 >= ny) return;
    const long long r1 = (r0 + ROWS_PER_CTA < ny) ? r0 + ROWS_PER_CTA : ny;
    const long long cl = (c0 == 0) ? nx - 1 : c0 - 1;          // left neighbour column (wrapped)
    const long long cr = (c0 + 2 >= nx) ? 0 : c0 + 2;          // right neighbour column (wrapped)
    const bool edge_r = (lane == 31) || (c0 + 2 >= nx);        // my right neighbour is not in lane+1
    const bool edge_l = (lane == 0);

    auto load_row = [&](long long j) -> double2 {
        if (!live) return make_double2(0.0, 0.0);
        return *reinterpret_cast<const double2*>(in + j * nx + c0);
    };
    // neighbours across a warp edge (lanes 0 / 31, and the row ends): one scalar each, fetched with the rows
    auto load_edges = [&](long long j, double& l, double& r) {
        l = (live && edge_l) ? __ldg(in + j * nx + cl) : 0.0;
        r = (live && edge_r) ? __ldg(in + j * nx + cr) : 0.0;
    };
    const long long jup0 = (r0 == 0) ? ny - 1 : r0 - 1;
    double2 up = load_row(jup0);
    double2 ce = load_row(r0);
    // software pipeline: the PF rows below the current batch (and the batch's edge values) are requested one
    // batch ahead, so PF row loads per thread are in flight while PF rows are being combined
    constexpr int PF = (OP == OP_LAPLACIAN) ? LM_K4_PF_LAP : LM_K4_PF_SMOOTH;   // rows per batch (measured best)
    double2 nxt[PF];
    double el[PF], er[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        nxt[k] = make_double2(0.0, 0.0); el[k] = 0.0; er[k] = 0.0;
        if (r0 + k < r1) {
            const long long jd = r0 + k + 1;
            nxt[k] = load_row(jd == ny ? 0 : jd);
            load_edges(r0 + k, el[k], er[k]);
        }
    }
    for (long long jb = r0; jb < r1; jb += PF) {
        double2 cur[PF];
        double cel[PF], cer[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) { cur[k] = nxt[k]; cel[k] = el[k]; cer[k] = er[k]; }
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const long long j = jb + PF + k;               // row of the next batch
            if (j < r1) {
                nxt[k] = load_row(j + 1 == ny ? 0 : j + 1);
                load_edges(j, el[k], er[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const long long j = jb + k;
            if (j >= r1) break;
            const double2 dn = cur[k];
            double lf = __shfl_up_sync(FULL, ce.y, 1);
            double rt = __shfl_down_sync(FULL, ce.x, 1);
            if (live) {
                if (edge_l) lf = cel[k];
                if (edge_r) rt = cer[k];
                double2 o;
                if (OP == OP_LAPLACIAN) {
                    o.x = combine<OP>(ce.x, up.x, dn.x, lf, ce.y, h2);
                    o.y = combine<OP>(ce.y, up.y, dn.y, ce.x, rt, h2);
                } else {
                    const bool row_border = (j == 0) || (j == ny - 1);
                    o.x = (row_border || c0 == 0) ? ce.x : combine<OP>(ce.x, up.x, dn.x, lf, ce.y, h2);
                    o.y = (row_border || c0 + 1 == nx - 1) ? ce.y : combine<OP>(ce.y, up.y, dn.y, ce.x, rt, h2);
                }
                *reinterpret_cast<double2*>(out + j * nx + c0) = o;
            }
            up = ce;
            ce = dn;
        }
    }
}

// generic kernel: any nx / alignment, one pixel per thread
template <int OP>
__global__ void __launch_bounds__(256) stencil_scalar_kernel(const double* __restrict__ in,
                                                             double* __restrict__ out,
                                                             long long ny, long long nx, double h2) {
    const long long total = ny * nx;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const long long j = idx / nx, i = idx - j * nx;
        const long long jm = (j == 0) ? ny - 1 : j - 1, jp = (j == ny - 1) ? 0 : j + 1;
        const long long im = (i == 0) ? nx - 1 : i - 1, ip = (i == nx - 1) ? 0 : i + 1;
        const double c = in[idx];
        if (OP == OP_SMOOTH && (j == 0 || i == 0 || j == ny - 1 || i == nx - 1)) {
            out[idx] = c;
        } else {
            out[idx] = combine<OP>(c, in[jm * nx + i], in[jp * nx + i], in[j * nx + im], in[j * nx + ip], h2);
        }
    }
}

template <int OP>
int32_t launch_stencil(const double* in, long long ny, long long nx, double h2, double* out, cudaStream_t s) {
    if (ny == 0 || nx == 0) return LM_OK;
    const bool vec = (nx % 2 == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % 16 == 0) && nx >= 4;
    if (vec) {
        const long long gx = (nx / 2 + ST_THREADS - 1) / ST_THREADS;
        const long long gy = (ny + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
        if (gy > 65535) return lm::fail(LM_E_INVALID, "stencil: ny too large (%lld rows)", ny);
        dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
        stencil_vec_kernel<OP><<<grid, ST_THREADS, 0, s>>>(in, out, ny, nx, h2);
    } else {
        const long long total = ny * nx;
        long long blocks = (total + 255) / 256;
        const long long cap = static_cast<long long>(lm::sm_count()) * 32;
        if (blocks > cap) blocks = cap;
        stencil_scalar_kernel<OP><<<static_cast<unsigned>(blocks), 256, 0, s>>>(in, out, ny, nx, h2);
    }
    LM_CUDA_TRY(cudaGetLastError());
    return LM_OK;
}

template <int OP>
int32_t host_stencil(const char* who, const double* in, int64_t ny, int64_t nx, double h2, double* out,
                     lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(ny >= 0 && nx >= 0, "%s: negative size", who);
    LM_REQUIRE((in && out) || ny * nx == 0, "%s: NULL buffer", who);
    if (stats) *stats = lm_stats{};
    if (ny * nx == 0) return LM_OK;
    const size_t nb = static_cast<size_t>(ny) * nx * sizeof(double);
    void *din, *dout;
    if ((rc = lm::ws_get(lm::WS_IN_A, nb, &din)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, nb, &dout)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(din, in, nb, cudaMemcpyHostToDevice, nullptr));
    lm::Timer tm;
    if ((rc = tm.begin(nullptr)) != LM_OK) return rc;
    if ((rc = launch_stencil<OP>(static_cast<double*>(din), ny, nx, h2, static_cast<double*>(dout), nullptr)) != LM_OK)
        return rc;
    float ms = 0.f;
    if ((rc = tm.end(nullptr, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(out, dout, nb, cudaMemcpyDeviceToHost, nullptr));
    LM_CUDA_TRY(cudaStreamSynchronize(nullptr));
    if (stats) {
        stats->items = static_cast<uint64_t>(ny) * nx;
        stats->work_units = stats->items * 16;     // algorithmic bytes
        stats->kernel_ms = ms;
        stats->launches = 1;
    }
    return LM_OK;
}

}  // namespace

extern "C" {

int32_t lm_laplacian5_periodic_dev(const double* U, int64_t ny, int64_t nx, double h, double* out, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(ny >= 0 && nx >= 0, "lm_laplacian5_periodic_dev: negative size");
    LM_REQUIRE((U && out) || ny * nx == 0, "lm_laplacian5_periodic_dev: NULL buffer");
    LM_REQUIRE(U != out, "lm_laplacian5_periodic_dev: in-place operation is not supported");
    return launch_stencil<OP_LAPLACIAN>(U, ny, nx, h * h, out, lm::as_stream(stream));
}

int32_t lm_laplacian5_periodic(const double* U, int64_t ny, int64_t nx, double h, double* out, lm_stats* stats) {
    return host_stencil<OP_LAPLACIAN>("lm_laplacian5_periodic", U, ny, nx, h * h, out, stats);
}

int32_t lm_smooth5_interior_dev(const double* g, int64_t ny, int64_t nx, double* out, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(ny >= 0 && nx >= 0, "lm_smooth5_interior_dev: negative size");
    LM_REQUIRE((g && out) || ny * nx == 0, "lm_smooth5_interior_dev: NULL buffer");
    LM_REQUIRE(g != out, "lm_smooth5_interior_dev: in-place operation is not supported");
    return launch_stencil<OP_SMOOTH>(g, ny, nx, 0.0, out, lm::as_stream(stream));
}

int32_t lm_smooth5_interior(const double* g, int64_t ny, int64_t nx, double* out, lm_stats* stats) {
    return host_stencil<OP_SMOOTH>("lm_smooth5_interior", g, ny, nx, 0.0, out, stats);
}

}  // extern "C"

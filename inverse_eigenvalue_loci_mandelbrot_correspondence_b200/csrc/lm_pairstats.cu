// lm_pairstats.cu -- binned statistics over all N(N-1)/2 point pairs (FP64 pipe).
//
// The reference materialises every pairwise distance on the host (scipy pdist / distance_matrix: 8 B per pair,
// 5.7 GB for the 37 820-point cloud of the tracker) and then bins it:
//   empirical_variogram_field / _coords        Variogram-Mandelbrot-Construct.py:106-152   (np.digitize on linspace bins)
//   empirical_variogram_from_field_locs        Iterative_Variogram_Laplacian.py:53-86      ((D >= b[k]) & (D < b[k+1]))
//   pair_correlation, ripley_K                 spatial_stats_phase2.py:9-47                (shells [r, r+dr), counts d < r)
// Every one of them is "count the pairs i<j with lo[k] <= d_ij < hi[k] and sum a per-pair weight", with
// d_ij = sqrt(dx*dx + dy*dy) in unfused binary64 (probe: scipy pdist, distance_matrix and np.linalg.norm(axis=1) all
// return exactly that).  Here nothing is materialised: a tile loop over the upper triangle, one i per thread, the j
// tile staged in shared memory, and the bin of every pair located EXACTLY against the caller's edges, so the counts
// are bit-identical to the reference's.  The comparison runs in the squared domain: sqrt_rn is monotone, so
//     sqrt_rn(s) >= L   <=>   s >= T(L),   T(L) = the smallest binary64 s whose correctly rounded root reaches L
// (found on the host per edge by stepping around fl(L*L)), which keeps the FP64 square root out of the per-pair path
// (it is only taken for the d*d weight).  The bin is guessed from an FP32 root and checked against T(lo[k]), T(lo[k+1]);
// a wrong guess (a pair within an FP32 ulp of an edge, or non-uniform edges) falls into a linear repair loop.
// Weights are summed per thread in runs of equal bin, per warp in a private shared-memory histogram, per block in
// a partial row, and across blocks in block order by a finishing kernel (the reference's np.mean is a pairwise
// sum: parity of the sums is 1e-12, of the counts exact).
#include "lm_common.cuh"

#include <math.h>

#include <vector>

namespace {

constexpr int PS_THREADS = 128;          // threads per block = i's per tile
constexpr int PS_TILE = 128;             // j's per tile
constexpr int PS_WARPS = PS_THREADS / 32;
constexpr size_t PS_HIST_BYTES = 16384;  // shared-memory budget of the per-block histogram copies
constexpr int PS_MAX_BINS = 2048;
constexpr int PS_BLOCKS_PER_SM = 16;         // upper bound of resident 128-thread blocks per SM (2048 threads)

// tile t of the row-major upper triangle (bi <= bj) of a T x T tile grid
__device__ __forceinline__ void decode_tile(long long t, long long T, int* bi, int* bj) {
    // rows before bi hold bi*T - bi*(bi-1)/2 tiles; solve with a double sqrt and repair
    const double b = 2.0 * static_cast<double>(T) + 1.0;
    long long r = static_cast<long long>((b - sqrt(b * b - 8.0 * static_cast<double>(t))) * 0.5);
    if (r < 0) r = 0;
    if (r > T - 1) r = T - 1;
    while (r > 0 && r * T - r * (r - 1) / 2 > t) --r;
    while (r + 1 < T && (r + 1) * T - (r + 1) * r / 2 <= t) ++r;
    *bi = static_cast<int>(r);
    *bj = static_cast<int>(r + (t - (r * T - r * (r - 1) / 2)));
}

// MUFU.SQRT: the guess only has to be close, the exact check follows
__device__ __forceinline__ float approx_sqrtf(float a) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

struct RunCache {
    int k = -1;
    unsigned cnt = 0;
    double sum = 0.0;
};

template <int WMODE>
__device__ __forceinline__ void flush_run(RunCache& rc, unsigned long long* hc, double* hs) {
    if (rc.k >= 0) {
        atomicAdd(&hc[rc.k], static_cast<unsigned long long>(rc.cnt));
        if (WMODE != LM_PAIR_W_NONE) atomicAdd(&hs[rc.k], rc.sum);
    }
}

// WMODE: LM_PAIR_W_NONE counts only, LM_PAIR_W_VALUE_SQDIFF (v_i - v_j)^2, LM_PAIR_W_DIST_SQ d^2.
// PARTITION: hi[k] == lo[k+1] for every k (np.digitize bins, Ripley's cumulated counts): no upper-edge array needed.
// tlo[nb + 1]: T(lo[k]) with a +inf sentinel; thi[nb]: T(hi[k]) (general shells only).
template <int WMODE, bool PARTITION>
__global__ void __launch_bounds__(PS_THREADS) pair_hist_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                               const double* __restrict__ v, long long n,
                                                               const double* __restrict__ tlo, const double* __restrict__ thi, int nb,
                                                               float lo0, float inv_w, double thi_last, int copies, int stride, long long T,
                                                               long long total_tiles, unsigned long long* __restrict__ part_counts,
                                                               double* __restrict__ part_sums) {
    extern __shared__ __align__(16) unsigned char ps_smem[];
    double* slo = reinterpret_cast<double*>(ps_smem);                          // [nb + 1]
    double* shi = slo + (nb + 1);                                              // [nb] (unused when PARTITION)
    double* ssum = shi + (PARTITION ? 0 : nb);                                 // [copies][stride], stride odd (bank spread)
    unsigned long long* scnt = reinterpret_cast<unsigned long long*>(ssum + static_cast<size_t>(copies) * stride);
    __shared__ double2 sxy[PS_TILE];
    __shared__ double sv[PS_TILE];

    const int tid = threadIdx.x;
    for (int k = tid; k <= nb; k += PS_THREADS) slo[k] = tlo[k];
    if (!PARTITION) for (int k = tid; k < nb; k += PS_THREADS) shi[k] = thi[k];
    for (int k = tid; k < copies * stride; k += PS_THREADS) { ssum[k] = 0.0; scnt[k] = 0ull; }
    // neighbouring lanes look at neighbouring points and tend to flush the same bin at the same time: consecutive
    // lanes own different histogram copies (copies is a power of two), so those atomics do not serialise
    const int copy = tid & (copies - 1);
    unsigned long long* hc = scnt + static_cast<size_t>(copy) * stride;
    double* hs = ssum + static_cast<size_t>(copy) * stride;
    RunCache run;
    const float kmax = static_cast<float>(nb - 1);

    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int bi, bj;
        decode_tile(t, T, &bi, &bj);
        const long long i = static_cast<long long>(bi) * PS_TILE + tid;
        const long long j0 = static_cast<long long>(bj) * PS_TILE;
        const bool live = i < n;
        const double xi = live ? x[i] : 0.0, yi = live ? y[i] : 0.0;
        double vi = 0.0;
        if (WMODE == LM_PAIR_W_VALUE_SQDIFF) vi = live ? v[i] : 0.0;
        __syncthreads();                                   // previous tile consumed (and the zero fill on the first pass)
        if (j0 + tid < n) {
            sxy[tid] = make_double2(x[j0 + tid], y[j0 + tid]);
            if (WMODE == LM_PAIR_W_VALUE_SQDIFF) sv[tid] = v[j0 + tid];
        }
        __syncthreads();
        if (!live) continue;
        const int jn = static_cast<int>(n - j0 < PS_TILE ? n - j0 : PS_TILE);
        const int jfirst = (bi == bj) ? tid + 1 : 0;      // strict upper triangle on diagonal tiles
#pragma unroll 2
        for (int j = jfirst; j < jn; ++j) {
            const double2 pj = sxy[j];
            const double dx = __dsub_rn(xi, pj.x), dy = __dsub_rn(yi, pj.y);
            const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));          // d = sqrt_rn(s)
            // k = largest index with lo[k] <= d, i.e. T(lo[k]) <= s: FP32 guess, exact check, rare repair
            int k = static_cast<int>(fminf(fmaxf((approx_sqrtf(static_cast<float>(s)) - lo0) * inv_w, 0.f), kmax));
            if (s < slo[k] || s >= slo[k + 1]) {
                while (k >= 0 && s < slo[k]) --k;
                while (k + 1 < nb && slo[k + 1] <= s) ++k;
                if (k < 0) continue;
            }
            double w = 0.0;
            if (WMODE == LM_PAIR_W_VALUE_SQDIFF) { const double dv = __dsub_rn(vi, sv[j]); w = __dmul_rn(dv, dv); }
            if (WMODE == LM_PAIR_W_DIST_SQ) { const double d = __dsqrt_rn(s); w = __dmul_rn(d, d); }
            const bool inside = PARTITION ? (k < nb - 1 || s < thi_last) : (s < shi[k]);
            if (inside) {
                if (k == run.k) { ++run.cnt; run.sum += w; }
                else { flush_run<WMODE>(run, hc, hs); run.k = k; run.cnt = 1; run.sum = w; }
            }
            if (!PARTITION && k >= 1 && s < shi[k - 1]) {  // shells whose upper edge overshoots the next lower edge by an ulp
                atomicAdd(&hc[k - 1], 1ull);
                if (WMODE != LM_PAIR_W_NONE) atomicAdd(&hs[k - 1], w);
            }
        }
    }
    flush_run<WMODE>(run, hc, hs);
    __syncthreads();
    for (int k = tid; k < nb; k += PS_THREADS) {
        unsigned long long c = 0ull;
        double sm = 0.0;
        for (int q = 0; q < copies; ++q) { c += scnt[static_cast<size_t>(q) * stride + k]; sm += ssum[static_cast<size_t>(q) * stride + k]; }
        part_counts[static_cast<size_t>(blockIdx.x) * nb + k] = c;
        if (WMODE != LM_PAIR_W_NONE) part_sums[static_cast<size_t>(blockIdx.x) * nb + k] = sm;
    }
}

__global__ void pair_hist_finish_kernel(const unsigned long long* __restrict__ part_counts, const double* __restrict__ part_sums,
                                        int blocks, int nb, int with_sums, unsigned long long* __restrict__ counts,
                                        double* __restrict__ sums) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    unsigned long long c = 0ull;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) {                    // block order: the same sum on every run
        c += part_counts[static_cast<size_t>(b) * nb + k];
        if (with_sums) s += part_sums[static_cast<size_t>(b) * nb + k];
    }
    counts[k] = c;
    sums[k] = s;
}

// max over pairs of dx*dx + dy*dy (sqrt is monotone: sqrt of the maximum is the maximum distance, D.max())
__global__ void __launch_bounds__(PS_THREADS) pair_max_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n,
                                                              long long T, long long total_tiles, unsigned long long* __restrict__ out_bits) {
    __shared__ double sx[PS_TILE], sy[PS_TILE];
    const int tid = threadIdx.x;
    double best = 0.0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int bi, bj;
        decode_tile(t, T, &bi, &bj);
        const long long i = static_cast<long long>(bi) * PS_TILE + tid;
        const long long j0 = static_cast<long long>(bj) * PS_TILE;
        const bool live = i < n;
        const double xi = live ? x[i] : 0.0, yi = live ? y[i] : 0.0;
        __syncthreads();
        if (j0 + tid < n) { sx[tid] = x[j0 + tid]; sy[tid] = y[j0 + tid]; }
        __syncthreads();
        if (!live) continue;
        const int jn = static_cast<int>(n - j0 < PS_TILE ? n - j0 : PS_TILE);
        const int jfirst = (bi == bj) ? tid + 1 : 0;
#pragma unroll 4
        for (int j = jfirst; j < jn; ++j) {
            const double dx = __dsub_rn(xi, sx[j]), dy = __dsub_rn(yi, sy[j]);
            const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            best = s > best ? s : best;                   // NaN pairs never win, like np.max would propagate: inputs must be finite
        }
    }
    unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(best));   // non-negative doubles order like integers
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if ((tid & 31) == 0) atomicMax(out_bits, bits);
}

struct TileGrid {
    long long T = 0, total = 0;
    unsigned blocks = 0;
};

TileGrid tile_grid(int64_t n) {
    TileGrid g;
    g.T = (n + PS_TILE - 1) / PS_TILE;
    g.total = g.T * (g.T + 1) / 2;
    const long long cap = static_cast<long long>(lm::sm_count()) * PS_BLOCKS_PER_SM;
    g.blocks = static_cast<unsigned>(g.total < cap ? g.total : cap);
    return g;
}

int32_t all_finite(const double* a, int64_t n, const char* what) {
    for (int64_t i = 0; i < n; ++i)
        if (!isfinite(a[i])) return lm::fail(LM_E_INVALID, "lm_pair: %s[%lld] is not finite", what, static_cast<long long>(i));
    return LM_OK;
}

// T(L): the smallest s >= 0 with sqrt_rn(s) >= L (sqrt_rn monotone => d >= L <=> s >= T(L)).
double sq_threshold(double L) {
    if (!(L > 0.0)) return 0.0;
    if (isinf(L)) return INFINITY;
    double c = L * L;
    if (isinf(c)) c = 1.7976931348623157e308;
    while (c > 0.0 && sqrt(c) >= L) c = nextafter(c, 0.0);
    while (sqrt(c) < L) {
        const double up = nextafter(c, INFINITY);
        if (isinf(up)) return INFINITY;                 // no finite s has a root that large
        c = up;
    }
    return c;
}

struct HistArgs {
    const double *x, *y, *v, *tlo, *thi;
    int64_t n; int nb; float lo0, inv_w; double thi_last; int copies, stride;
    unsigned long long* pc; double* psum;
};

// persistent grid: exactly the blocks that are resident at once (tiles are dealt round-robin, so a partial second
// wave would leave most SMs idle at the end); returns the grid size used (= rows of the partial arrays)
template <int WMODE, bool PARTITION>
unsigned launch_hist(const TileGrid& g, size_t smem, cudaStream_t s, const HistArgs& a) {
    cudaFuncSetAttribute(pair_hist_kernel<WMODE, PARTITION>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_hist_kernel<WMODE, PARTITION>, PS_THREADS, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    const long long resident = static_cast<long long>(lm::sm_count()) * per_sm;
    const unsigned blocks = static_cast<unsigned>(g.total < resident ? g.total : resident);
    pair_hist_kernel<WMODE, PARTITION><<<blocks, PS_THREADS, smem, s>>>(a.x, a.y, a.v, a.n, a.tlo, a.thi, a.nb, a.lo0, a.inv_w, a.thi_last,
                                                                         a.copies, a.stride, g.T, g.total, a.pc, a.psum);
    return blocks;
}

template <bool PARTITION>
unsigned launch_hist_mode(int weight_mode, const TileGrid& g, size_t smem, cudaStream_t s, const HistArgs& a) {
    if (weight_mode == LM_PAIR_W_NONE) return launch_hist<LM_PAIR_W_NONE, PARTITION>(g, smem, s, a);
    if (weight_mode == LM_PAIR_W_VALUE_SQDIFF) return launch_hist<LM_PAIR_W_VALUE_SQDIFF, PARTITION>(g, smem, s, a);
    return launch_hist<LM_PAIR_W_DIST_SQ, PARTITION>(g, smem, s, a);
}


// ---------------------------------------------------------------------------------------
// ordered selection: the (v_i - w_j)^2 of the pairs (i in A, j in B) whose distance falls into ONE bin, in row-major
// order of (i, j) -- what dV2[np.where(m)] is in sample_semivariogram / sample_cross_semivariogram
// (variograms_construct_mandelbrot.py:178-315) for the one block in which a bin crosses its max_pairs_per_bin cap and
// the reference draws a random subset of exactly this list.  Distances are sqrt_rn(dx*dx + dy*dy) (np.linalg.norm
// over the last axis) compared like the reference's mask (D >= lo) & (D < hi).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_in_bin(double xa, double ya, double xb, double yb, double lo, double hi) {
    const double dx = __dsub_rn(xa, xb), dy = __dsub_rn(ya, yb);
    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    return d >= lo && d < hi;
}

// one warp per row i of A: number of selected j
__global__ void __launch_bounds__(256) pair_select_count_kernel(const double* __restrict__ xa, const double* __restrict__ ya, int na,
                                                                const double* __restrict__ xb, const double* __restrict__ yb, int nb,
                                                                double lo, double hi, int skip_diag, unsigned* __restrict__ row_count) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= na) return;
    const double x = xa[i], y = ya[i];
    unsigned c = 0;
    for (int j = lane; j < nb; j += 32)
        if (!(skip_diag && j == i) && pair_in_bin(x, y, xb[j], yb[j], lo, hi)) ++c;
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) row_count[i] = c;
}

// one warp per row: the selected values at row_offset[i] .. in ascending j (ballot compaction keeps the order)
__global__ void __launch_bounds__(256) pair_select_write_kernel(const double* __restrict__ xa, const double* __restrict__ ya,
                                                                const double* __restrict__ va, int na,
                                                                const double* __restrict__ xb, const double* __restrict__ yb,
                                                                const double* __restrict__ vb, int nb, double lo, double hi,
                                                                int skip_diag, const unsigned long long* __restrict__ row_offset,
                                                                double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= na) return;
    const double x = xa[i], y = ya[i], v = va[i];
    unsigned long long pos = row_offset[i];
    for (int j0 = 0; j0 < nb; j0 += 32) {
        const int j = j0 + lane;
        const bool sel = j < nb && !(skip_diag && j == i) && pair_in_bin(x, y, xb[j], yb[j], lo, hi);
        const unsigned bal = __ballot_sync(0xffffffffu, sel);
        if (sel) {
            const double dv = __dsub_rn(v, vb[j]);
            out[pos + __popc(bal & ((1u << lane) - 1u))] = __dmul_rn(dv, dv);
        }
        pos += __popc(bal);
    }
}

}  // namespace

extern "C" {

int32_t lm_pair_histogram(const double* x, const double* y, const double* value, int64_t n,
                          const double* lo, const double* hi, int32_t nbins, int32_t weight_mode,
                          uint64_t* counts, double* sums, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0, "lm_pair_histogram: negative size");
    LM_REQUIRE(nbins >= 1 && nbins <= PS_MAX_BINS, "lm_pair_histogram: nbins must be in [1, %d]", PS_MAX_BINS);
    LM_REQUIRE(weight_mode == LM_PAIR_W_NONE || weight_mode == LM_PAIR_W_VALUE_SQDIFF || weight_mode == LM_PAIR_W_DIST_SQ,
               "lm_pair_histogram: unknown weight_mode %d", weight_mode);
    LM_REQUIRE(lo && hi && counts, "lm_pair_histogram: NULL edge / count buffer");
    LM_REQUIRE(weight_mode == LM_PAIR_W_NONE || sums, "lm_pair_histogram: sums is NULL");
    LM_REQUIRE(n == 0 || (x && y), "lm_pair_histogram: NULL coordinate buffer");
    LM_REQUIRE(weight_mode != LM_PAIR_W_VALUE_SQDIFF || n == 0 || value, "lm_pair_histogram: value is NULL");
    for (int k = 0; k < nbins; ++k) {
        LM_REQUIRE(isfinite(lo[k]) && !isnan(hi[k]), "lm_pair_histogram: bad edge at bin %d", k);
        LM_REQUIRE(k == 0 || lo[k] > lo[k - 1], "lm_pair_histogram: lower edges must increase (bin %d)", k);
        LM_REQUIRE(k + 2 >= nbins || hi[k] <= lo[k + 2], "lm_pair_histogram: bin %d overlaps bin %d", k, k + 2);
    }
    if ((rc = all_finite(x, n, "x")) != LM_OK) return rc;
    if ((rc = all_finite(y, n, "y")) != LM_OK) return rc;
    if (stats) *stats = lm_stats{};
    for (int k = 0; k < nbins; ++k) { counts[k] = 0; if (sums) sums[k] = 0.0; }
    if (n < 2) return LM_OK;

    cudaStream_t s = nullptr;
    const size_t pb = static_cast<size_t>(n) * sizeof(double), eb = static_cast<size_t>(nbins) * sizeof(double);
    const TileGrid g = tile_grid(n);
    void *dx, *dy, *dv, *dlo, *dhi, *dpc, *dps, *dfin;
    if ((rc = lm::ws_get(lm::WS_IN_A, pb, &dx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, pb, &dy)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_C, pb, &dv)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, eb + sizeof(double), &dlo)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, eb, &dhi)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, eb * g.blocks, &dpc)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_D, eb * g.blocks, &dps)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 2 * eb + 64, &dfin)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dx, x, pb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dy, y, pb, cudaMemcpyHostToDevice, s));
    if (weight_mode == LM_PAIR_W_VALUE_SQDIFF) LM_CUDA_TRY(cudaMemcpyAsync(dv, value, pb, cudaMemcpyHostToDevice, s));
    std::vector<double> tlo(static_cast<size_t>(nbins) + 1), thi(static_cast<size_t>(nbins));
    bool partition = true;
    for (int k = 0; k < nbins; ++k) {
        tlo[k] = sq_threshold(lo[k]);
        thi[k] = sq_threshold(hi[k]);
        if (k + 1 < nbins && hi[k] != lo[k + 1]) partition = false;
    }
    tlo[nbins] = INFINITY;                                  // sentinel: the exact check reads slo[k + 1]
    LM_CUDA_TRY(cudaMemcpyAsync(dlo, tlo.data(), eb + sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dhi, thi.data(), eb, cudaMemcpyHostToDevice, s));

    const int stride = nbins | 1;
    int copies = 1;
    while (copies < PS_THREADS && static_cast<size_t>(2 * copies) * stride * 16 <= PS_HIST_BYTES) copies *= 2;
    const size_t smem = (static_cast<size_t>(nbins) + 1 + (partition ? 0 : nbins)) * 8 + static_cast<size_t>(copies) * stride * 16;
    const double span = lo[nbins - 1] - lo[0];
    const double inv_w = (nbins > 1 && span > 0.0 && isfinite(span)) ? static_cast<double>(nbins - 1) / span : 0.0;
    unsigned long long* pc = static_cast<unsigned long long*>(dpc);
    double* psum = static_cast<double*>(dps);
    unsigned long long* fin_c = static_cast<unsigned long long*>(dfin);
    double* fin_s = reinterpret_cast<double*>(fin_c + nbins);

    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    HistArgs a{static_cast<double*>(dx), static_cast<double*>(dy), static_cast<double*>(dv), static_cast<double*>(dlo),
               static_cast<double*>(dhi), n, nbins, static_cast<float>(lo[0]), static_cast<float>(inv_w), thi[nbins - 1], copies, stride, pc, psum};
    const unsigned blocks = partition ? launch_hist_mode<true>(weight_mode, g, smem, s, a) : launch_hist_mode<false>(weight_mode, g, smem, s, a);
    LM_CUDA_TRY(cudaGetLastError());
    pair_hist_finish_kernel<<<(nbins + 127) / 128, 128, 0, s>>>(pc, psum, static_cast<int>(blocks), nbins,
                                                               weight_mode != LM_PAIR_W_NONE, fin_c, fin_s);
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(counts, fin_c, eb, cudaMemcpyDeviceToHost, s));
    if (sums) LM_CUDA_TRY(cudaMemcpyAsync(sums, fin_s, eb, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(n);
        stats->work_units = static_cast<uint64_t>(n) * static_cast<uint64_t>(n - 1) / 2;
        stats->kernel_ms = ms;
        stats->launches = 2;
    }
    return LM_OK;
}

int32_t lm_pair_max_distance(const double* x, const double* y, int64_t n, double* dmax, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0 && dmax, "lm_pair_max_distance: bad arguments");
    LM_REQUIRE(n == 0 || (x && y), "lm_pair_max_distance: NULL coordinate buffer");
    if ((rc = all_finite(x, n, "x")) != LM_OK) return rc;
    if ((rc = all_finite(y, n, "y")) != LM_OK) return rc;
    if (stats) *stats = lm_stats{};
    *dmax = 0.0;
    if (n < 2) return LM_OK;
    cudaStream_t s = nullptr;
    const size_t pb = static_cast<size_t>(n) * sizeof(double);
    const TileGrid g = tile_grid(n);
    void *dx, *dy, *dout;
    if ((rc = lm::ws_get(lm::WS_IN_A, pb, &dx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, pb, &dy)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dout)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dx, x, pb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dy, y, pb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(dout, 0, 8, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    pair_max_kernel<<<g.blocks, PS_THREADS, 0, s>>>(static_cast<double*>(dx), static_cast<double*>(dy), n, g.T, g.total,
                                                    static_cast<unsigned long long*>(dout));
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    double s2 = 0.0;
    LM_CUDA_TRY(cudaMemcpyAsync(&s2, dout, 8, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    *dmax = sqrt(s2);
    if (stats) {
        stats->items = static_cast<uint64_t>(n);
        stats->work_units = static_cast<uint64_t>(n) * static_cast<uint64_t>(n - 1) / 2;
        stats->kernel_ms = ms;
        stats->launches = 1;
    }
    return LM_OK;
}


int32_t lm_pair_select_sqdiff(const double* xa, const double* ya, const double* va, int64_t na,
                              const double* xb, const double* yb, const double* vb, int64_t nb,
                              double lo, double hi, int32_t skip_diagonal,
                              double* values, int64_t cap_values, int64_t* n_values, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(na >= 0 && nb >= 0 && na < (1 << 30) && nb < (1 << 30) && cap_values >= 0 && n_values, "lm_pair_select_sqdiff: bad sizes");
    LM_REQUIRE((na == 0 || (xa && ya && va)) && (nb == 0 || (xb && yb && vb)), "lm_pair_select_sqdiff: NULL buffer");
    if (stats) *stats = lm_stats{};
    *n_values = 0;
    if (na == 0 || nb == 0) return LM_OK;
    cudaStream_t s = nullptr;
    void *da, *db, *dcnt, *doff, *dout = nullptr;
    if ((rc = lm::ws_get(lm::WS_IN_A, static_cast<size_t>(na) * 3 * sizeof(double), &da)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, static_cast<size_t>(nb) * 3 * sizeof(double), &db)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, static_cast<size_t>(na) * sizeof(unsigned), &dcnt)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, static_cast<size_t>(na) * sizeof(unsigned long long), &doff)) != LM_OK) return rc;
    double* A = static_cast<double*>(da); double* B = static_cast<double*>(db);
    LM_CUDA_TRY(cudaMemcpyAsync(A, xa, na * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(A + na, ya, na * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(A + 2 * na, va, na * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(B, xb, nb * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(B + nb, yb, nb * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(B + 2 * nb, vb, nb * sizeof(double), cudaMemcpyHostToDevice, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    const unsigned blocks = static_cast<unsigned>((na + 7) / 8);
    pair_select_count_kernel<<<blocks, 256, 0, s>>>(A, A + na, static_cast<int>(na), B, B + nb, static_cast<int>(nb), lo, hi,
                                                    skip_diagonal != 0, static_cast<unsigned*>(dcnt));
    LM_CUDA_TRY(cudaGetLastError());
    std::vector<unsigned> cnt(static_cast<size_t>(na));
    LM_CUDA_TRY(cudaMemcpyAsync(cnt.data(), dcnt, static_cast<size_t>(na) * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    std::vector<unsigned long long> off(static_cast<size_t>(na));
    unsigned long long total = 0;
    for (int64_t i = 0; i < na; ++i) { off[static_cast<size_t>(i)] = total; total += cnt[static_cast<size_t>(i)]; }   // 4000-row blocks: a host scan
    *n_values = static_cast<int64_t>(total);
    if (static_cast<int64_t>(total) > cap_values)
        return lm::fail(LM_E_CAP, "lm_pair_select_sqdiff: %llu pairs selected, room for %lld", total, static_cast<long long>(cap_values));
    if (total) {
        LM_REQUIRE(values != nullptr, "lm_pair_select_sqdiff: values is NULL");
        if ((rc = lm::ws_get(lm::WS_OUT_C, static_cast<size_t>(total) * sizeof(double), &dout)) != LM_OK) return rc;
        LM_CUDA_TRY(cudaMemcpyAsync(doff, off.data(), static_cast<size_t>(na) * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        pair_select_write_kernel<<<blocks, 256, 0, s>>>(A, A + na, A + 2 * na, static_cast<int>(na), B, B + nb, B + 2 * nb,
                                                        static_cast<int>(nb), lo, hi, skip_diagonal != 0,
                                                        static_cast<unsigned long long*>(doff), static_cast<double*>(dout));
        LM_CUDA_TRY(cudaGetLastError());
    }
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    if (total) {
        LM_CUDA_TRY(cudaMemcpyAsync(values, dout, static_cast<size_t>(total) * sizeof(double), cudaMemcpyDeviceToHost, s));
        LM_CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (stats) { stats->items = static_cast<uint64_t>(na) * static_cast<uint64_t>(nb); stats->work_units = total; stats->kernel_ms = ms; stats->launches = total ? 2 : 1; }
    return LM_OK;
}

}  // extern "C"

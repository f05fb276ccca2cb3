// lm_density.cu -- the tracker's density stage on the device (second half; HBM / launch bound).
//
//   mollified_histogram(mod, cloud, bins, sigma_bins)          gi_assumption_tracker_v3.py:109-127
//       np.histogram2d -> max(eps) -> scipy.ndimage.gaussian_filter(mode="nearest") -> max(eps) -> H / H.sum()
//   tv_distance, overlap_mass                                  gi_assumption_tracker_v3.py:91-96
//   KL(P, X) of the stock module                               tci_construct_mandelbrot_v002_fixed.py:84-86
//   gi_flow_fixed_T, gi_flow_to_threshold                      gi_assumption_tracker_v3.py:130-151
//       up to --max-steps 800 sweeps of X <- (1-alpha) X + alpha P with a KL evaluation after each, over bins^2
//       (<= 1024^2) cells: the CPU cost of a tracker level once the generators are on the GPU.
//
// Everything except the logarithm is restated operation for operation, so it is BIT-exact against numpy / scipy:
//   * bin lookup = np.searchsorted(edges, v, side="right") - 1 against the caller's np.linspace edges, right edge closed;
//   * the blur is scipy's correlate1d for symmetric kernels: centre term first, then pairs from the far end inwards,
//     (in[l+j] + in[l-j]) * w[j], axis 0 then axis 1, borders clamped (probe: 0 mismatches against gaussian_filter);
//   * every reduction (H.sum(), sum|p-q|, sum min(p,q), KL) runs in NUMPY'S PAIRWISE ORDER: blocks of <= 128
//     elements with 8 strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), halves split at
//     n/2 rounded down to a multiple of 8 (probe: equal to np.sum for every size tried).  Eight lanes own the eight
//     accumulators of a block (the xor butterfly is that parenthesisation), the tree above the blocks is evaluated
//     level by level by one CTA;
//   * the flow update is the unfused fl(fl((1-alpha)*X) + fl(alpha*P)) (file compiled with -fmad=false).
// KL uses CUDA's log (<= 1 ulp from numpy's): KL values agree to ~1e-15 absolute, X_T and T_n exactly.
#include "lm_common.cuh"

#include <math.h>

#include <algorithm>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------------------
// numpy-order sum plan
// ---------------------------------------------------------------------------------------------------------------
constexpr int PW_BLOCK = 128;        // numpy's PW_BLOCKSIZE
constexpr int SUM_SETS = 3;          // independent value sets per plan (tv, overlap, KL in one pass)

struct SumPlan {
    int64_t n = -1;
    int dev = -1;
    int n_leaves = 0, n_internal = 0, n_levels = 0;
    void* block = nullptr;           // one device allocation holding the arrays below
    long long* leaf_off = nullptr;
    int* leaf_len = nullptr;
    int2* children = nullptr;        // per internal node (final numbering), indices into vals
    int* level_start = nullptr;      // [n_levels + 1], in internal-node numbering
    double* vals = nullptr;          // [SUM_SETS][n_leaves + n_internal]
    int nodes() const { return n_leaves + n_internal; }
};
SumPlan g_plan;

void plan_release() {
    if (g_plan.block) {
        int cur = 0;
        cudaGetDevice(&cur);
        if (g_plan.dev != cur) cudaSetDevice(g_plan.dev);
        cudaFree(g_plan.block);
        if (g_plan.dev != cur) cudaSetDevice(cur);
    }
    g_plan = SumPlan();
}

struct HostNode { int left, right, height; };

// DOUBLE_pairwise_sum's recursion (numpy/_core/src/umath/loops_utils.h.src): leaves are runs of <= 128 elements
int build_tree(long long off, long long n, std::vector<long long>& loff, std::vector<int>& llen, std::vector<HostNode>& internal,
               int* height) {
    if (n <= PW_BLOCK) {
        loff.push_back(off); llen.push_back(static_cast<int>(n));
        *height = 0;
        return static_cast<int>(loff.size()) - 1;                 // leaves: ids >= 0
    }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    int hl = 0, hr = 0;
    const int l = build_tree(off, n2, loff, llen, internal, &hl);
    const int r = build_tree(off + n2, n - n2, loff, llen, internal, &hr);
    *height = 1 + std::max(hl, hr);
    internal.push_back(HostNode{l, r, *height});
    return -static_cast<int>(internal.size());                    // internal: -(index + 1)
}

int32_t plan_get(int64_t n, SumPlan** out) {
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    if (g_plan.block && g_plan.n == n && g_plan.dev == dev) { *out = &g_plan; return LM_OK; }
    plan_release();
    lm::register_release_hook(plan_release);
    std::vector<long long> loff; std::vector<int> llen; std::vector<HostNode> internal;
    int h = 0;
    build_tree(0, n, loff, llen, internal, &h);
    const int L = static_cast<int>(loff.size()), I = static_cast<int>(internal.size());
    // renumber the internal nodes by height so that a level only reads finished values
    std::vector<int> order(I);
    for (int i = 0; i < I; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return internal[a].height < internal[b].height; });
    std::vector<int> rank(I);
    for (int i = 0; i < I; ++i) rank[order[i]] = i;
    auto final_id = [&](int id) { return id >= 0 ? id : L + rank[-id - 1]; };
    std::vector<int2> children(I > 0 ? I : 1);
    std::vector<int> level_start;
    int prev_h = 0;
    for (int i = 0; i < I; ++i) {
        const HostNode& nd = internal[order[i]];
        if (nd.height != prev_h) { level_start.push_back(i); prev_h = nd.height; }
        children[i] = make_int2(final_id(nd.left), final_id(nd.right));
    }
    level_start.push_back(I);
    const int n_levels = static_cast<int>(level_start.size()) - 1;
    // one allocation: leaf_off | vals | leaf_len | children | level_start  (8-byte items first)
    const size_t b_off = sizeof(long long) * L, b_vals = sizeof(double) * SUM_SETS * (L + I), b_len = sizeof(int) * L,
                 b_ch = sizeof(int2) * (I > 0 ? I : 1), b_ls = sizeof(int) * level_start.size();
    void* block = nullptr;
    cudaError_t e = cudaMalloc(&block, b_off + b_vals + b_ch + b_len + b_ls + 64);
    if (e != cudaSuccess) { cudaGetLastError(); return lm::fail(LM_E_NOMEM, "lm_density: plan allocation failed: %s", cudaGetErrorString(e)); }
    unsigned char* p = static_cast<unsigned char*>(block);
    g_plan.block = block; g_plan.dev = dev; g_plan.n = n;
    g_plan.n_leaves = L; g_plan.n_internal = I; g_plan.n_levels = n_levels;
    g_plan.leaf_off = reinterpret_cast<long long*>(p); p += b_off;
    g_plan.vals = reinterpret_cast<double*>(p); p += b_vals;
    g_plan.children = reinterpret_cast<int2*>(p); p += b_ch;
    g_plan.leaf_len = reinterpret_cast<int*>(p); p += b_len;
    g_plan.level_start = reinterpret_cast<int*>(p);
    LM_CUDA_TRY(cudaMemcpy(g_plan.leaf_off, loff.data(), b_off, cudaMemcpyHostToDevice));
    LM_CUDA_TRY(cudaMemcpy(g_plan.leaf_len, llen.data(), b_len, cudaMemcpyHostToDevice));
    LM_CUDA_TRY(cudaMemcpy(g_plan.children, children.data(), sizeof(int2) * I, cudaMemcpyHostToDevice));
    LM_CUDA_TRY(cudaMemcpy(g_plan.level_start, level_start.data(), b_ls, cudaMemcpyHostToDevice));
    *out = &g_plan;
    return LM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// element operators of the leaf kernel
// ---------------------------------------------------------------------------------------------------------------
enum { OP_IDENT = 0, OP_COMPARE = 1, OP_KL = 2, OP_FLOW_KL = 3 };

struct OpArgs {
    const double* a;     // IDENT: the array;  COMPARE: p;  KL / FLOW_KL: P
    const double* b;     // COMPARE: q;  KL: X
    double* x;           // FLOW_KL: X, updated in place
    double alpha, one_minus_alpha, eps;
    const int* done;     // FLOW_KL: skip the sweep once the threshold was met
};

__device__ __forceinline__ double kl_term(double p, double x, double eps) {
    const double p_ = p < eps ? eps : p, x_ = x < eps ? eps : x;          // np.clip(., eps, None)
    return __dmul_rn(p_, __dsub_rn(log(p_), log(x_)));
}

// NV values per element: COMPARE produces |p-q|, min(p,q) and the KL(p, q) term together
template <int OP> struct OpWidth { static constexpr int NV = (OP == OP_COMPARE) ? 3 : 1; };

template <int OP>
__device__ __forceinline__ void elem(const OpArgs& g, long long i, double* v) {
    if (OP == OP_IDENT) v[0] = g.a[i];
    if (OP == OP_COMPARE) {
        const double p = g.a[i], q = g.b[i];
        v[0] = fabs(__dsub_rn(p, q));
        v[1] = p < q ? p : q;
        v[OpWidth<OP>::NV - 1] = kl_term(p, q, g.eps);
    }
    if (OP == OP_KL) v[0] = kl_term(g.a[i], g.b[i], g.eps);
    if (OP == OP_FLOW_KL) {
        const double p = g.a[i];
        const double x = __dadd_rn(__dmul_rn(g.one_minus_alpha, g.x[i]), __dmul_rn(g.alpha, p));
        g.x[i] = x;
        v[0] = kl_term(p, x, g.eps);
    }
}

// 8 lanes per leaf: lane k owns numpy's accumulator r[k]
template <int OP>
__global__ void __launch_bounds__(256) pw_leaf_kernel(OpArgs g, const long long* __restrict__ leaf_off, const int* __restrict__ leaf_len,
                                                      int n_leaves, int nodes, double* __restrict__ vals) {
    constexpr int NV = OpWidth<OP>::NV;
    if (OP == OP_FLOW_KL && *g.done) return;
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int leaf = static_cast<int>(gid >> 3), k = static_cast<int>(gid & 7);
    const bool live = leaf < n_leaves;
    const long long off = live ? leaf_off[leaf] : 0;
    const int len = live ? leaf_len[leaf] : 0;
    double r[NV], v[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) r[q] = 0.0;
    const int body = len - (len & 7);
    if (len >= 8) {
        elem<OP>(g, off + k, r);
        for (int i = 8; i < body; i += 8) {
            elem<OP>(g, off + i + k, v);
#pragma unroll
            for (int q = 0; q < NV; ++q) r[q] = __dadd_rn(r[q], v[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        r[q] = __dadd_rn(r[q], __shfl_xor_sync(0xffffffffu, r[q], 1));
        r[q] = __dadd_rn(r[q], __shfl_xor_sync(0xffffffffu, r[q], 2));
        r[q] = __dadd_rn(r[q], __shfl_xor_sync(0xffffffffu, r[q], 4));
    }
    if (live && k == 0) {
        // len < 8: numpy's plain loop from 0.0 (r is 0 here); otherwise the tail after the 8-wide body
        for (int i = (len >= 8 ? body : 0); i < len; ++i) {
            elem<OP>(g, off + i, v);
#pragma unroll
            for (int q = 0; q < NV; ++q) r[q] = __dadd_rn(r[q], v[q]);
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) vals[static_cast<size_t>(q) * nodes + leaf] = r[q];
    }
}

// one CTA walks the levels of the tree; result[q] = root of value set q.  In flow mode (kl_hist != NULL) it also records
// the step's KL and raises the stop flag exactly like gi_flow_to_threshold: t >= min_steps and kl <= threshold.
__global__ void __launch_bounds__(1024) pw_combine_kernel(const int2* __restrict__ children, const int* __restrict__ level_start, int n_levels,
                                                          int n_leaves, int nodes, int nsets, double* __restrict__ vals,
                                                          double* __restrict__ result, double* kl_hist, int step, int min_steps,
                                                          double threshold, int* done, int* steps_done) {
    if (kl_hist && *done) return;
    for (int L = 0; L < n_levels; ++L) {
        const int a = level_start[L], b = level_start[L + 1];
        for (int q = 0; q < nsets; ++q) {
            double* v = vals + static_cast<size_t>(q) * nodes;
            for (int i = a + threadIdx.x; i < b; i += blockDim.x) {
                const int2 c = children[i];
                v[n_leaves + i] = __dadd_rn(v[c.x], v[c.y]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int q = 0; q < nsets; ++q) result[q] = vals[static_cast<size_t>(q) * nodes + nodes - 1];
        if (kl_hist) {
            const double kl = result[0];
            kl_hist[step] = kl;
            *steps_done = step;
            if (step >= 1 && step >= min_steps && kl <= threshold) *done = 1;
        }
    }
}

template <int OP>
int32_t launch_sum(SumPlan* pl, const OpArgs& g, double* result_dev, cudaStream_t s, double* kl_hist = nullptr, int step = 0,
                   int min_steps = 0, double threshold = 0.0, int* done = nullptr, int* steps_done = nullptr) {
    const long long threads = static_cast<long long>(pl->n_leaves) * 8;
    const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
    pw_leaf_kernel<OP><<<blocks, 256, 0, s>>>(g, pl->leaf_off, pl->leaf_len, pl->n_leaves, pl->nodes(), pl->vals);
    pw_combine_kernel<<<1, 1024, 0, s>>>(pl->children, pl->level_start, pl->n_levels, pl->n_leaves, pl->nodes(), OpWidth<OP>::NV,
                                         pl->vals, result_dev, kl_hist, step, min_steps, threshold, done, steps_done);
    LM_CUDA_TRY(cudaGetLastError());
    return LM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// histogram2d and the blur
// ---------------------------------------------------------------------------------------------------------------
// np.searchsorted(e, v, side="right") - 1 with the closed right edge of np.histogramdd; -1 = outlier (or NaN)
__device__ __forceinline__ int locate_bin(double v, const double* __restrict__ e, int nb, double inv_w) {
    if (!(v >= e[0]) || v > e[nb]) return -1;
    if (v == e[nb]) return nb - 1;
    int k = static_cast<int>(fmin(fmax((v - e[0]) * inv_w, 0.0), static_cast<double>(nb - 1)));
    while (k > 0 && v < e[k]) --k;
    while (k + 1 < nb && e[k + 1] <= v) ++k;
    return k;
}

__global__ void hist2d_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n, const double* __restrict__ xe, int nbx,
                              double inv_wx, const double* __restrict__ ye, int nby, double inv_wy, unsigned* __restrict__ counts) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bx = locate_bin(x[i], xe, nbx, inv_wx);
    const int by = locate_bin(y[i], ye, nby, inv_wy);
    if (bx >= 0 && by >= 0) atomicAdd(&counts[static_cast<size_t>(bx) * nby + by], 1u);
}

__global__ void counts_to_f64_kernel(const unsigned* __restrict__ counts, long long n, double floor_eps, double* __restrict__ H) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double h = static_cast<double>(counts[i]);
    H[i] = h < floor_eps ? floor_eps : h;                                   // np.maximum(H, eps); eps <= 0 leaves H as is
}

// scipy correlate1d, symmetric kernel, mode="nearest".  AXIS 0: along the slow index, AXIS 1: along the fast one.
template <int AXIS>
__global__ void blur_kernel(const double* __restrict__ in, int n0, int n1, const double* __restrict__ w, int radius, double floor_eps,
                            int apply_floor, double* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(n0) * n1) return;
    const int i = static_cast<int>(idx / n1), j = static_cast<int>(idx - static_cast<long long>(i) * n1);
    const int len = AXIS == 0 ? n0 : n1, pos = AXIS == 0 ? i : j;
    const long long stride = AXIS == 0 ? n1 : 1;
    const double* line = in + (AXIS == 0 ? static_cast<long long>(j) : static_cast<long long>(i) * n1);
    double t = __dmul_rn(line[pos * stride], w[radius]);
    for (int jj = -radius; jj < 0; ++jj) {
        int lo = pos + jj, hi = pos - jj;
        lo = lo < 0 ? 0 : lo;
        hi = hi > len - 1 ? len - 1 : hi;
        t = __dadd_rn(t, __dmul_rn(__dadd_rn(line[lo * stride], line[hi * stride]), w[radius + jj]));
    }
    if (apply_floor) t = t < floor_eps ? floor_eps : t;
    out[idx] = t;
}

__global__ void divide_kernel(double* __restrict__ H, long long n, const double* __restrict__ total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) H[i] = __ddiv_rn(H[i], *total);
}

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

double inv_width(const double* e, int nb) {
    const double span = e[nb] - e[0];
    return span > 0.0 ? static_cast<double>(nb) / span : 0.0;
}

int32_t check_edges(const double* e, int nb, const char* what) {
    LM_REQUIRE(e != nullptr, "lm_density: %s edges are NULL", what);
    for (int k = 0; k <= nb; ++k) {
        LM_REQUIRE(isfinite(e[k]), "lm_density: %s edge %d is not finite", what, k);
        LM_REQUIRE(k == 0 || e[k] >= e[k - 1], "lm_density: %s edges must be monotonically increasing", what);
    }
    return LM_OK;
}

int32_t check_weights(const double* w, int radius) {
    LM_REQUIRE(radius >= 0 && radius <= 4096, "lm_density: blur radius out of range");
    if (radius == 0) return LM_OK;
    LM_REQUIRE(w != nullptr, "lm_density: blur weights are NULL");
    for (int j = 1; j <= radius; ++j)
        LM_REQUIRE(w[radius + j] == w[radius - j], "lm_density: blur weights must be symmetric (scipy's symmetric correlate1d path)");
    return LM_OK;
}

// device pipeline shared by lm_histogram2d and lm_mollified_histogram; leaves the result in *H_out (device)
int32_t histogram_device(const double* x, const double* y, int64_t n, const double* xe, int32_t nbx, const double* ye, int32_t nby,
                         double eps, bool mollify, const double* w, int32_t radius, cudaStream_t s, double** H_out, int* launches) {
    int32_t rc;
    const long long cells = static_cast<long long>(nbx) * nby;
    const size_t pb = static_cast<size_t>(n) * sizeof(double);
    void *dx, *dy, *dxe, *dye, *dcnt, *dH, *dT, *dw, *dres;
    if ((rc = lm::ws_get(lm::WS_IN_A, pb, &dx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, pb, &dy)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_XS, sizeof(double) * (nbx + 1), &dxe)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_YS, sizeof(double) * (nby + 1), &dye)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_I32, sizeof(unsigned) * cells, &dcnt)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_F64, sizeof(double) * cells, &dH)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_FIELD, sizeof(double) * cells, &dT)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_C, sizeof(double) * (2 * static_cast<size_t>(radius) + 1), &dw)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dres)) != LM_OK) return rc;
    if (n) {
        LM_CUDA_TRY(cudaMemcpyAsync(dx, x, pb, cudaMemcpyHostToDevice, s));
        LM_CUDA_TRY(cudaMemcpyAsync(dy, y, pb, cudaMemcpyHostToDevice, s));
    }
    LM_CUDA_TRY(cudaMemcpyAsync(dxe, xe, sizeof(double) * (nbx + 1), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dye, ye, sizeof(double) * (nby + 1), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(dcnt, 0, sizeof(unsigned) * cells, s));
    double* H = static_cast<double*>(dH);
    double* T = static_cast<double*>(dT);
    if (n) hist2d_kernel<<<blocks_for(n, 256), 256, 0, s>>>(static_cast<double*>(dx), static_cast<double*>(dy), n, static_cast<double*>(dxe), nbx,
                                                            inv_width(xe, nbx), static_cast<double*>(dye), nby, inv_width(ye, nby),
                                                            static_cast<unsigned*>(dcnt));
    counts_to_f64_kernel<<<blocks_for(cells, 256), 256, 0, s>>>(static_cast<unsigned*>(dcnt), cells, mollify ? eps : 0.0, H);
    *launches = n ? 2 : 1;
    if (mollify) {
        if (radius > 0) {
            LM_CUDA_TRY(cudaMemcpyAsync(dw, w, sizeof(double) * (2 * radius + 1), cudaMemcpyHostToDevice, s));
            blur_kernel<0><<<blocks_for(cells, 256), 256, 0, s>>>(H, nbx, nby, static_cast<double*>(dw), radius, eps, 0, T);
            blur_kernel<1><<<blocks_for(cells, 256), 256, 0, s>>>(T, nbx, nby, static_cast<double*>(dw), radius, eps, 1, H);
            *launches += 2;
        }
        SumPlan* pl;
        if ((rc = plan_get(cells, &pl)) != LM_OK) return rc;
        OpArgs g{};
        g.a = H;
        if ((rc = launch_sum<OP_IDENT>(pl, g, static_cast<double*>(dres), s)) != LM_OK) return rc;
        divide_kernel<<<blocks_for(cells, 256), 256, 0, s>>>(H, cells, static_cast<double*>(dres));
        *launches += 3;
    }
    LM_CUDA_TRY(cudaGetLastError());
    *H_out = H;
    return LM_OK;
}

int32_t histogram_entry(const double* x, const double* y, int64_t n, const double* xe, int32_t nbx, const double* ye, int32_t nby,
                        double eps, bool mollify, const double* w, int32_t radius, double* out, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0 && nbx >= 1 && nby >= 1 && out, "lm_histogram2d: bad arguments");
    LM_REQUIRE(static_cast<long long>(nbx) * nby <= (1ll << 28), "lm_histogram2d: more than 2^28 cells");
    LM_REQUIRE(n == 0 || (x && y), "lm_histogram2d: NULL sample buffer");
    LM_REQUIRE(n < (1ll << 32), "lm_histogram2d: more than 2^32 samples");
    if ((rc = check_edges(xe, nbx, "x")) != LM_OK) return rc;
    if ((rc = check_edges(ye, nby, "y")) != LM_OK) return rc;
    if (mollify && (rc = check_weights(w, radius)) != LM_OK) return rc;
    if (stats) *stats = lm_stats{};
    cudaStream_t s = nullptr;
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    double* H = nullptr;
    int launches = 0;
    if ((rc = histogram_device(x, y, n, xe, nbx, ye, nby, eps, mollify, w, mollify ? radius : 0, s, &H, &launches)) != LM_OK) return rc;
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(out, H, sizeof(double) * static_cast<size_t>(nbx) * nby, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(n);
        stats->work_units = static_cast<uint64_t>(nbx) * nby;
        stats->kernel_ms = ms;
        stats->launches = launches;
    }
    return LM_OK;
}

}  // namespace

extern "C" {

int32_t lm_histogram2d(const double* x, const double* y, int64_t n, const double* xedges, int32_t nbx,
                       const double* yedges, int32_t nby, double* H, lm_stats* stats) {
    return histogram_entry(x, y, n, xedges, nbx, yedges, nby, 0.0, false, nullptr, 0, H, stats);
}

int32_t lm_mollified_histogram(const double* x, const double* y, int64_t n, const double* xedges, int32_t nbx,
                               const double* yedges, int32_t nby, double eps, const double* weights, int32_t radius,
                               double* P, lm_stats* stats) {
    return histogram_entry(x, y, n, xedges, nbx, yedges, nby, eps, true, weights, radius, P, stats);
}

int32_t lm_gaussian_filter_nearest(const double* in, int64_t n0, int64_t n1, const double* weights, int32_t radius,
                                   double* out, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n0 >= 0 && n1 >= 0 && n0 * n1 <= (1ll << 30) && n0 < (1ll << 31) && n1 < (1ll << 31), "lm_gaussian_filter_nearest: bad shape");
    LM_REQUIRE(radius >= 1, "lm_gaussian_filter_nearest: radius must be >= 1");
    if ((rc = check_weights(weights, radius)) != LM_OK) return rc;
    if (stats) *stats = lm_stats{};
    const long long cells = n0 * n1;
    if (cells == 0) return LM_OK;
    LM_REQUIRE(in && out, "lm_gaussian_filter_nearest: NULL buffer");
    cudaStream_t s = nullptr;
    void *dA, *dB, *dw;
    if ((rc = lm::ws_get(lm::WS_OUT_F64, sizeof(double) * cells, &dA)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_FIELD, sizeof(double) * cells, &dB)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_C, sizeof(double) * (2 * static_cast<size_t>(radius) + 1), &dw)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dA, in, sizeof(double) * cells, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dw, weights, sizeof(double) * (2 * radius + 1), cudaMemcpyHostToDevice, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    blur_kernel<0><<<blocks_for(cells, 256), 256, 0, s>>>(static_cast<double*>(dA), static_cast<int>(n0), static_cast<int>(n1),
                                                          static_cast<double*>(dw), radius, 0.0, 0, static_cast<double*>(dB));
    blur_kernel<1><<<blocks_for(cells, 256), 256, 0, s>>>(static_cast<double*>(dB), static_cast<int>(n0), static_cast<int>(n1),
                                                          static_cast<double*>(dw), radius, 0.0, 0, static_cast<double*>(dA));
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(out, dA, sizeof(double) * cells, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) { stats->items = static_cast<uint64_t>(cells); stats->work_units = static_cast<uint64_t>(cells) * (2 * radius + 1) * 2;
                 stats->kernel_ms = ms; stats->launches = 2; }
    return LM_OK;
}

int32_t lm_sum_pairwise(const double* a, int64_t n, double* out, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0 && out && (n == 0 || a), "lm_sum_pairwise: bad arguments");
    if (stats) *stats = lm_stats{};
    *out = 0.0;
    if (n == 0) return LM_OK;
    cudaStream_t s = nullptr;
    void *da, *dres;
    if ((rc = lm::ws_get(lm::WS_OUT_F64, sizeof(double) * n, &da)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dres)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(da, a, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    SumPlan* pl;
    if ((rc = plan_get(n, &pl)) != LM_OK) return rc;
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    OpArgs g{};
    g.a = static_cast<double*>(da);
    if ((rc = launch_sum<OP_IDENT>(pl, g, static_cast<double*>(dres), s)) != LM_OK) return rc;
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(out, dres, sizeof(double), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) { stats->items = static_cast<uint64_t>(n); stats->work_units = static_cast<uint64_t>(n); stats->kernel_ms = ms; stats->launches = 2; }
    return LM_OK;
}

int32_t lm_density_compare(const double* p, const double* q, int64_t n, double eps,
                           double* sum_abs_diff, double* sum_min, double* kl_pq, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 1 && p && q, "lm_density_compare: bad arguments");
    if (stats) *stats = lm_stats{};
    cudaStream_t s = nullptr;
    void *dp, *dq, *dres;
    if ((rc = lm::ws_get(lm::WS_OUT_F64, sizeof(double) * n, &dp)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_FIELD, sizeof(double) * n, &dq)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dres)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dp, p, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dq, q, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    SumPlan* pl;
    if ((rc = plan_get(n, &pl)) != LM_OK) return rc;
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    OpArgs g{};
    g.a = static_cast<double*>(dp); g.b = static_cast<double*>(dq); g.eps = eps;
    if ((rc = launch_sum<OP_COMPARE>(pl, g, static_cast<double*>(dres), s)) != LM_OK) return rc;
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    double res[3] = {0, 0, 0};
    LM_CUDA_TRY(cudaMemcpyAsync(res, dres, sizeof(res), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (sum_abs_diff) *sum_abs_diff = res[0];
    if (sum_min) *sum_min = res[1];
    if (kl_pq) *kl_pq = res[2];
    if (stats) { stats->items = static_cast<uint64_t>(n); stats->work_units = static_cast<uint64_t>(n) * 3; stats->kernel_ms = ms; stats->launches = 2; }
    return LM_OK;
}

int32_t lm_gi_flow(const double* P, const double* X0, int64_t n, double alpha, double eps,
                   int32_t max_steps, int32_t min_steps, double kl_threshold, int32_t fixed_T,
                   double* X_out, int32_t* steps_out, double* kl_initial, double* kl_final,
                   double* kl_history, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 1 && P && X0 && X_out && steps_out && kl_initial && kl_final, "lm_gi_flow: bad arguments");
    LM_REQUIRE(max_steps >= 0 && max_steps <= (1 << 24), "lm_gi_flow: max_steps out of range");
    if (stats) *stats = lm_stats{};
    cudaStream_t s = nullptr;
    void *dP, *dX, *dhist, *dres;
    if ((rc = lm::ws_get(lm::WS_OUT_F64, sizeof(double) * n, &dP)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_FIELD, sizeof(double) * n, &dX)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, sizeof(double) * (static_cast<size_t>(max_steps) + 1), &dhist)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dres)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dP, P, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dX, X0, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(dres, 0, 64, s));
    double* res = static_cast<double*>(dres);                 // [0..2] sums, then the flags
    int* done = reinterpret_cast<int*>(res + 4);
    int* steps_done = done + 1;
    SumPlan* pl;
    if ((rc = plan_get(n, &pl)) != LM_OK) return rc;
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    OpArgs g{};
    g.a = static_cast<double*>(dP); g.b = static_cast<double*>(dX); g.x = static_cast<double*>(dX);
    g.alpha = alpha; g.one_minus_alpha = 1.0 - alpha; g.eps = eps; g.done = done;
    // fixed T: the threshold can never be met (kl <= -inf is false; a NaN compares false as well)
    const double thr = fixed_T ? -INFINITY : kl_threshold;
    double* hist = static_cast<double*>(dhist);
    int launches = 0;
    if ((rc = launch_sum<OP_KL>(pl, g, res, s, hist, 0, min_steps, thr, done, steps_done)) != LM_OK) return rc;
    launches += 2;
    int flag = 0;
    for (int t = 1; t <= max_steps && !flag; ++t) {
        if ((rc = launch_sum<OP_FLOW_KL>(pl, g, res, s, hist, t, min_steps, thr, done, steps_done)) != LM_OK) return rc;
        launches += 2;
        if (!fixed_T && (t % 16 == 0)) {                      // look at the stop flag now and then; late sweeps are no-ops
            LM_CUDA_TRY(cudaMemcpyAsync(&flag, done, sizeof(int), cudaMemcpyDeviceToHost, s));
            LM_CUDA_TRY(cudaStreamSynchronize(s));
        }
    }
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    int T = 0;
    LM_CUDA_TRY(cudaMemcpyAsync(&T, steps_done, sizeof(int), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaMemcpyAsync(X_out, dX, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    std::vector<double> h(static_cast<size_t>(T) + 1);
    LM_CUDA_TRY(cudaMemcpy(h.data(), hist, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    *steps_out = T;
    *kl_initial = h[0];
    *kl_final = h[T];
    if (kl_history) for (int t = 0; t <= T; ++t) kl_history[t] = h[t];
    if (stats) { stats->items = static_cast<uint64_t>(n); stats->work_units = static_cast<uint64_t>(n) * (static_cast<uint64_t>(T) + 1);
                 stats->kernel_ms = ms; stats->launches = launches; }
    return LM_OK;
}

}  // extern "C"

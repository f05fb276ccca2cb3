// lm_alpha.cu -- the alpha-shape edge filter of the boundary consumers.
//
//   circumradius(p, q, r), alpha_shape_edges(P, alpha)          construct_boundary_alpha.py:45-82
//       for every Delaunay triangle: R = abc / (4 sqrt(s(s-a)(s-b)(s-c)) + 1e-16), keep it when R < 1/alpha; an edge is a
//       boundary edge when exactly one kept triangle uses it; the list comes out in the order in which the edges are
//       first met walking the kept triangles ((t0,t1), (t1,t2), (t2,t0), each as (min, max)).
// The reference does this with a Python loop over tri.simplices (np.linalg.norm per side) and a dict.  Here: one thread
// per triangle for the radius test, a device hash table (64-bit (min,max) keys, linear probing) for the edge
// multiplicities, then flag / scan / scatter of the occurrences whose edge was counted once -- which, walking in
// triangle order, IS the reference's first-occurrence order.  The triangulation itself (scipy.spatial.Delaunay / Qhull)
// is the caller's: the entry point takes the simplices.
// Arithmetic: np.linalg.norm of a 2-vector is sqrt(ddot(v, v)); OpenBLAS's ddot contracts the second product
// (probe on the build host: sqrt(fma(dy, dy, dx*dx)) reproduces it, the unfused form differs in 8 % of the cases), so
// that is what is restated; everything after it is written operation for operation.
#include "lm_common.cuh"

#include <math.h>

namespace {

constexpr unsigned long long EMPTY_KEY = ~0ull;

__device__ __forceinline__ double side(double ax, double ay, double bx, double by) {
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
    return __dsqrt_rn(__fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

__device__ __forceinline__ double circumradius(double px, double py, double qx, double qy, double rx, double ry) {
    const double a = side(qx, qy, rx, ry), b = side(px, py, rx, ry), c = side(px, py, qx, qy);
    const double s = __ddiv_rn(__dadd_rn(__dadd_rn(a, b), c), 2.0);
    double A = __dmul_rn(__dmul_rn(__dmul_rn(s, __dsub_rn(s, a)), __dsub_rn(s, b)), __dsub_rn(s, c));
    A = (0.0 > A) ? 0.0 : A;                                      // Python's max(A, 0.0): the first argument stays unless 0.0 > A (NaN stays)
    if (A == 0.0) return INFINITY;
    const double area = __dsqrt_rn(A);
    return __ddiv_rn(__dmul_rn(__dmul_rn(a, b), c), __dadd_rn(__dmul_rn(4.0, area), 1e-16));
}

__device__ __forceinline__ unsigned long long edge_key(int i, int j) {
    const unsigned lo = static_cast<unsigned>(i < j ? i : j), hi = static_cast<unsigned>(i < j ? j : i);
    return (static_cast<unsigned long long>(lo) << 32) | hi;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ void table_add(unsigned long long* keys, unsigned* counts, unsigned long long mask, unsigned long long key) {
    unsigned long long slot = mix64(key) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&keys[slot], EMPTY_KEY, key);
        if (prev == EMPTY_KEY || prev == key) { atomicAdd(&counts[slot], 1u); return; }
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ unsigned table_count(const unsigned long long* keys, const unsigned* counts, unsigned long long mask,
                                                unsigned long long key) {
    unsigned long long slot = mix64(key) & mask;
    for (;;) {
        const unsigned long long k = keys[slot];
        if (k == key) return counts[slot];
        if (k == EMPTY_KEY) return 0u;
        slot = (slot + 1) & mask;
    }
}

__global__ void alpha_keep_kernel(const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ tri, long long ntri,
                                  double inv_alpha, unsigned char* __restrict__ keep, double* __restrict__ radius,
                                  unsigned long long* keys, unsigned* counts, unsigned long long mask) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= ntri) return;
    const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
    const double R = circumradius(x[i0], y[i0], x[i1], y[i1], x[i2], y[i2]);
    const bool k = R < inv_alpha;
    keep[t] = k ? 1 : 0;
    if (radius) radius[t] = R;
    if (k) {
        table_add(keys, counts, mask, edge_key(i0, i1));
        table_add(keys, counts, mask, edge_key(i1, i2));
        table_add(keys, counts, mask, edge_key(i2, i0));
    }
}

// pass 0: per-block number of boundary-edge occurrences; pass 1: scatter them behind the block's offset
template <int PASS>
__global__ void __launch_bounds__(256) alpha_edges_kernel(const int* __restrict__ tri, long long ntri, const unsigned char* __restrict__ keep,
                                                          const unsigned long long* __restrict__ keys, const unsigned* __restrict__ counts,
                                                          unsigned long long mask, long long* __restrict__ block_sums,
                                                          int* __restrict__ edges, long long cap_edges) {
    __shared__ int warp_tot[8];
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    unsigned long long k[3] = {0, 0, 0};
    bool once[3] = {false, false, false};
    if (t < ntri && keep[t]) {
        const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
        k[0] = edge_key(i0, i1); k[1] = edge_key(i1, i2); k[2] = edge_key(i2, i0);
#pragma unroll
        for (int j = 0; j < 3; ++j) once[j] = table_count(keys, counts, mask, k[j]) == 1u;
    }
    const int mine = int(once[0]) + int(once[1]) + int(once[2]);
    // exclusive scan of `mine` over the block
    int incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) before += warp_tot[w]; total += warp_tot[w]; }
    if (PASS == 0) {
        if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
        return;
    }
    long long pos = block_sums[blockIdx.x] + before + incl - mine;
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (once[j]) {
            if (pos < cap_edges) { edges[2 * pos] = static_cast<int>(k[j] >> 32); edges[2 * pos + 1] = static_cast<int>(k[j] & 0xffffffffull); }
            ++pos;
        }
}

// exclusive scan of the block sums by one CTA; total -> *n_out
__global__ void __launch_bounds__(1024) alpha_scan_kernel(long long* __restrict__ sums, long long n, long long* __restrict__ n_out) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        const long long v = i < n ? sums[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        long long before = 0, total = 0;
        for (int w = 0; w < 32; ++w) { if (w < warp) before += warp_tot[w]; total += warp_tot[w]; }
        if (i < n) sums[i] = carry + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}

}  // namespace

extern "C" {

int32_t lm_alpha_shape_edges(const double* x, const double* y, int64_t npts, const int32_t* simplices, int64_t ntri, double alpha,
                             uint8_t* keep, double* radius, int32_t* edges, int64_t cap_edges, int64_t* n_edges, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npts >= 0 && ntri >= 0 && cap_edges >= 0 && n_edges, "lm_alpha_shape_edges: bad arguments");
    LM_REQUIRE(npts < (1ll << 31) && ntri < (1ll << 30), "lm_alpha_shape_edges: too many points / triangles");
    LM_REQUIRE(alpha != 0.0 && alpha == alpha, "lm_alpha_shape_edges: alpha must be non-zero (the reference divides by it)");
    LM_REQUIRE(ntri == 0 || (x && y && simplices), "lm_alpha_shape_edges: NULL input buffer");
    LM_REQUIRE(cap_edges == 0 || edges, "lm_alpha_shape_edges: edges is NULL");
    for (int64_t k = 0; k < 3 * ntri; ++k)
        LM_REQUIRE(simplices[k] >= 0 && simplices[k] < npts, "lm_alpha_shape_edges: vertex index %d out of range (triangle %lld)",
                   simplices[k], static_cast<long long>(k / 3));
    if (stats) *stats = lm_stats{};
    *n_edges = 0;
    if (ntri == 0) return LM_OK;
    cudaStream_t s = nullptr;
    unsigned long long slots = 1024;
    while (slots < 8ull * static_cast<unsigned long long>(ntri)) slots <<= 1;       // <= 3 ntri keys: load factor <= 3/8
    const unsigned blocks = static_cast<unsigned>((ntri + 255) / 256);
    const size_t pb = static_cast<size_t>(npts) * sizeof(double);
    void *dx, *dy, *dtri, *dkeep, *drad, *dkeys, *dcounts, *dsums, *dedges;
    if ((rc = lm::ws_get(lm::WS_IN_A, pb, &dx)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, pb, &dy)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_C, sizeof(int) * 3 * static_cast<size_t>(ntri), &dtri)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_I32, static_cast<size_t>(ntri), &dkeep)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_F64, sizeof(double) * static_cast<size_t>(ntri), &drad)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, sizeof(unsigned long long) * slots, &dkeys)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, sizeof(unsigned) * slots, &dcounts)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, sizeof(long long) * (static_cast<size_t>(blocks) + 2), &dsums)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_D, sizeof(int) * 2 * static_cast<size_t>(cap_edges ? cap_edges : 1), &dedges)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dx, x, pb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dy, y, pb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dtri, simplices, sizeof(int) * 3 * static_cast<size_t>(ntri), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(dkeys, 0xff, sizeof(unsigned long long) * slots, s));
    LM_CUDA_TRY(cudaMemsetAsync(dcounts, 0, sizeof(unsigned) * slots, s));
    long long* sums = static_cast<long long*>(dsums);
    long long* total = sums + blocks + 1;
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    const double inv_alpha = 1.0 / alpha;
    alpha_keep_kernel<<<blocks, 256, 0, s>>>(static_cast<double*>(dx), static_cast<double*>(dy), static_cast<int*>(dtri), ntri, inv_alpha,
                                             static_cast<unsigned char*>(dkeep), static_cast<double*>(drad),
                                             static_cast<unsigned long long*>(dkeys), static_cast<unsigned*>(dcounts), slots - 1);
    alpha_edges_kernel<0><<<blocks, 256, 0, s>>>(static_cast<int*>(dtri), ntri, static_cast<unsigned char*>(dkeep),
                                                 static_cast<unsigned long long*>(dkeys), static_cast<unsigned*>(dcounts), slots - 1, sums,
                                                 nullptr, 0);
    alpha_scan_kernel<<<1, 1024, 0, s>>>(sums, blocks, total);
    alpha_edges_kernel<1><<<blocks, 256, 0, s>>>(static_cast<int*>(dtri), ntri, static_cast<unsigned char*>(dkeep),
                                                 static_cast<unsigned long long*>(dkeys), static_cast<unsigned*>(dcounts), slots - 1, sums,
                                                 static_cast<int*>(dedges), cap_edges);
    LM_CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    long long n = 0;
    LM_CUDA_TRY(cudaMemcpyAsync(&n, total, sizeof(n), cudaMemcpyDeviceToHost, s));
    if (keep) LM_CUDA_TRY(cudaMemcpyAsync(keep, dkeep, static_cast<size_t>(ntri), cudaMemcpyDeviceToHost, s));
    if (radius) LM_CUDA_TRY(cudaMemcpyAsync(radius, drad, sizeof(double) * static_cast<size_t>(ntri), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    *n_edges = n;
    if (stats) { stats->items = static_cast<uint64_t>(ntri); stats->work_units = static_cast<uint64_t>(ntri); stats->kernel_ms = ms; stats->launches = 4; }
    if (n > cap_edges)
        return lm::fail(LM_E_CAP, "lm_alpha_shape_edges: %lld boundary edges, capacity %lld", n, static_cast<long long>(cap_edges));
    if (n) LM_CUDA_TRY(cudaMemcpy(edges, dedges, sizeof(int) * 2 * static_cast<size_t>(n), cudaMemcpyDeviceToHost));
    return LM_OK;
}

}  // extern "C"

// lm_contour_link.cu -- K2, second half: the ORDERED boundary polylines, on the device.
//
// Replaces the line assembly of plt.contour(xs, ys, Z, levels=[level]) as used by
//   extract_contour   mandelbrot_boundary_sample.py:41-54
//   extract_contour   mandelbrot_boundary_sample_spyder.py:35-43
// (contourpy "mpl2014": lines(), get_start_edge, follow_interior -- SURVEY.md Appendix B).
//
// Input: the crossing records of lm_contour.cu in raster order (one record per crossed quad, up to two
// segments each: entry edge, exit edge, exit vertex).  mpl2014 walks the quads sequentially; what it
// produces is nevertheless a pure function of the segment graph, and that function is what runs here:
//
//   node          (record k, segment s), id v = 2k + s.  Ids are monotone in mpl2014's scan order
//                 (quads in raster order; inside a quad the start edges are tried S, W, N, E, which is
//                 the order the records list their segments).
//   succ / pred   a segment leaves through its exit edge into the neighbour quad's segment that enters
//                 through the opposite edge (or leaves the grid).  succ is injective, so every node has
//                 at most one predecessor: the graph is a set of open chains and cycles.
//   open lines    chains whose first node enters through a grid-border edge.  mpl2014 emits them first, in
//                 scan order of that node  ->  ordered by the id of their first node ("head").
//   closed lines  cycles.  The scan starts a loop at the first quad that still has an untraversed segment
//                 (saddle quads are revisited until both segments are used)  ->  ordered by their smallest
//                 node id, and started at that node ("leader").
//   vertices      entry vertex of the first node, then the exit vertex of every node in chain order; a loop
//                 whose leader is entered through its N edge drops the entry vertex and repeats its first
//                 exit vertex at the end.  A line with L nodes always has L + 1 vertices.
//
// Kernels:
//   link_build_kernel   thread per record: successor of each segment (E/W neighbours are the adjacent
//                       records, N/S neighbours by binary search on the quad ids), predecessor by scatter.
//   link_rank_kernel    one cooperative (grid-synchronised) kernel:
//                         A. pointer jumping along pred with a running minimum: ceil(log2(nodes)) rounds,
//                            in place on packed 64-bit (pointer, min) words -- every node of a chain ends at
//                            its head, every node of a cycle knows the cycle's smallest id;
//                         B. the same jumping with (pointer, distance) from the leaders: the rank of every
//                            node inside its line (Wyllie list ranking);
//                         C. line lengths (written by each line's last node), a two-level exclusive scan
//                            over the leaders (open lines first, then loops) -> line offsets;
//                         D. every node stores its exit vertex at offset[line] + rank (+1), leaders add the
//                            entry / closing vertex.
// The result stays in HBM; the host copies exactly n_vertices x 16 B + (n_lines + 1) x 8 B back
// (the host linker of round 1 pulled 64 B per record and walked them on one thread).
#include "lm_common.cuh"

#include <cooperative_groups.h>

#include <climits>
#include <cstring>

namespace cg = cooperative_groups;

namespace {

constexpr int REC_WORDS = 8;
constexpr int LINK_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

enum { EDGE_E = 0, EDGE_N = 1, EDGE_W = 2, EDGE_S = 3 };
enum : unsigned { NODE_VALID = 1u, NODE_HEAD = 2u, NODE_LEADER = 16u };      // bits 2-3: entry edge
enum : unsigned { ERR_ORDER = 1u, ERR_NEIGHBOUR = 2u, ERR_PRED = 4u, ERR_RANK = 8u, ERR_META = 16u };

__device__ __forceinline__ bool on_border(int edge, long long i, long long j, long long nx, long long ny) {
    switch (edge) {
        case EDGE_E: return i == nx - 2;
        case EDGE_N: return j == ny - 2;
        case EDGE_W: return i == 0;
        default:     return j == 0;
    }
}

// first record index in [lo, hi) whose quad id is >= q
__device__ __forceinline__ int lower_bound_quad(const long long* __restrict__ rec, int lo, int hi, long long q) {
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (__ldg(rec + static_cast<long long>(mid) * REC_WORDS) < q) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ unsigned long long pack(int lo, unsigned hi) {
    return static_cast<unsigned long long>(static_cast<unsigned>(lo)) | (static_cast<unsigned long long>(hi) << 32);
}
__device__ __forceinline__ int lo32(unsigned long long w) { return static_cast<int>(static_cast<unsigned>(w)); }
__device__ __forceinline__ unsigned hi32(unsigned long long w) { return static_cast<unsigned>(w >> 32); }

// counters: [0] error flags, [1] number of nodes
__global__ void __launch_bounds__(LINK_THREADS) link_build_kernel(
    const long long* __restrict__ rec, int n, long long nx, long long ny,
    int* __restrict__ succ, int* __restrict__ pred, unsigned char* __restrict__ info, unsigned* __restrict__ counters) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned err = 0u, nvalid = 0u;
    if (k < n) {
        const long long q = __ldg(rec + static_cast<long long>(k) * REC_WORDS);
        const unsigned m = static_cast<unsigned>(__ldg(rec + static_cast<long long>(k) * REC_WORDS + 3));
        if (k > 0 && __ldg(rec + static_cast<long long>(k - 1) * REC_WORDS) >= q) err |= ERR_ORDER;     // raster order
        int nseg = static_cast<int>((m >> 16) & 3u);
        if (nseg > 2) { nseg = 2; err |= ERR_META; }
        const long long j = q / nx, i = q - j * nx;
        if (q < 0 || j > ny - 2 || i > nx - 2) { err |= ERR_ORDER; nseg = 0; }
        for (int s = 0; s < 2; ++s) {
            const int v = 2 * k + s;
            if (s >= nseg) { info[v] = 0; succ[v] = -1; continue; }
            const int en = static_cast<int>((m >> (8 + 4 * s)) & 3u), ex = static_cast<int>((m >> (10 + 4 * s)) & 3u);
            const bool head = on_border(en, i, j, nx, ny);
            info[v] = static_cast<unsigned char>(NODE_VALID | (head ? NODE_HEAD : 0u) | (static_cast<unsigned>(en) << 2));
            ++nvalid;
            int t = -1;
            if (!on_border(ex, i, j, nx, ny)) {
                int k2 = -1;
                switch (ex) {
                    case EDGE_E: if (k + 1 < n && __ldg(rec + static_cast<long long>(k + 1) * REC_WORDS) == q + 1) k2 = k + 1; break;
                    case EDGE_W: if (k > 0 && __ldg(rec + static_cast<long long>(k - 1) * REC_WORDS) == q - 1) k2 = k - 1; break;
                    case EDGE_N: { const int c = lower_bound_quad(rec, k + 1, n, q + nx);
                                   if (c < n && __ldg(rec + static_cast<long long>(c) * REC_WORDS) == q + nx) k2 = c; } break;
                    default:     { const int c = lower_bound_quad(rec, 0, k, q - nx);
                                   if (c < k && __ldg(rec + static_cast<long long>(c) * REC_WORDS) == q - nx) k2 = c; } break;
                }
                if (k2 >= 0) {
                    const unsigned m2 = static_cast<unsigned>(__ldg(rec + static_cast<long long>(k2) * REC_WORDS + 3));
                    const int nseg2 = min(static_cast<int>((m2 >> 16) & 3u), 2);
                    const unsigned want = static_cast<unsigned>((ex + 2) & 3);      // entered through the opposite edge
                    for (int s2 = 0; s2 < nseg2; ++s2)
                        if (((m2 >> (8 + 4 * s2)) & 3u) == want) t = 2 * k2 + s2;
                }
                if (t < 0) err |= ERR_NEIGHBOUR; else pred[t] = v;
            }
            succ[v] = t;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) nvalid += __shfl_xor_sync(FULL, nvalid, o);
    if ((threadIdx.x & 31) == 0 && nvalid) atomicAdd(counters + 1, nvalid);
    if (err) atomicOr(counters, err);
}

struct LinkArgs {
    const long long* rec; int n;
    long long nx, ny; double level;
    const double *xs, *ys;                       // full-grid coordinates (device)
    const int* succ; const int* pred; unsigned char* info;
    unsigned long long *st_a, *st_b;             // packed (pointer, min) / (pointer, distance)
    unsigned* len;                               // vertices of the line led by this node (0 elsewhere)
    unsigned long long* lineoff;                 // leaders: vertex offset of the line | N-start flag << 63
    unsigned long long* part;                    // [grid][4] per-CTA sums: open verts, open lines, loop verts, loop lines
    double2* verts; long long* offsets;          // output
    unsigned* counters;                          // [0] error flags, [1] nodes
    unsigned long long* result;                  // [0] n_vertices, [1] n_lines
};

// exclusive scan of one value per thread over the CTA; returns the CTA total in `total`
__device__ __forceinline__ unsigned long long block_exscan(unsigned long long v, unsigned long long* s_warp, unsigned long long& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();                             // s_warp may still be read by the previous call
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    unsigned long long before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < LINK_THREADS / 32; ++w) {
        const unsigned long long sw = s_warp[w];
        if (w < warp) before += sw;
        all += sw;
    }
    total = all;
    return before + x - v;
}

__global__ void __launch_bounds__(LINK_THREADS) link_rank_kernel(const LinkArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned long long s_warp[LINK_THREADS / 32];
    __shared__ unsigned long long s_red[4];
    const int NS = 2 * a.n;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    unsigned err = 0u;
    const unsigned nvalid = __ldcg(a.counters + 1);
    int rounds = 1;
    while (rounds < 31 && (1u << rounds) < nvalid) ++rounds;          // 2^rounds >= the longest possible line

    // ---- A. heads of the open chains, smallest id of every cycle
    for (int v = tid; v < NS; v += nth) {
        unsigned f = a.info[v];
        if (!(f & NODE_VALID)) continue;
        int p = v;
        if (!(f & NODE_HEAD)) {
            p = a.pred[v];
            if (p < 0) { p = v; err |= ERR_PRED; a.info[v] = static_cast<unsigned char>(f | NODE_HEAD); }    // inconsistent input: stay in range
        }
        __stcg(a.st_a + v, pack(p, static_cast<unsigned>(v)));
    }
    grid.sync();
    for (int r = 0; r < rounds; ++r) {
        for (int v = tid; v < NS; v += nth) {
            if ((a.info[v] & (NODE_VALID | NODE_HEAD)) != NODE_VALID) continue;
            const unsigned long long s = __ldcg(a.st_a + v);
            const unsigned long long s2 = __ldcg(a.st_a + lo32(s));       // one 64-bit word: a consistent (pointer, min) pair
            __stcg(a.st_a + v, pack(lo32(s2), min(hi32(s), hi32(s2))));
        }
        grid.sync();
    }
    // ---- B. rank inside the line
    for (int v = tid; v < NS; v += nth) {
        const unsigned f = a.info[v];
        if (!(f & NODE_VALID)) continue;
        bool lead = (f & NODE_HEAD) != 0u;
        if (!lead) {
            const unsigned long long s = __ldcg(a.st_a + v);
            lead = !(a.info[lo32(s)] & NODE_HEAD) && hi32(s) == static_cast<unsigned>(v);
        }
        if (lead) a.info[v] = static_cast<unsigned char>(f | NODE_LEADER);
        __stcg(a.st_b + v, lead ? pack(v, 0u) : pack(a.pred[v], 1u));
    }
    grid.sync();
    for (int r = 0; r < rounds; ++r) {
        for (int v = tid; v < NS; v += nth) {
            if ((a.info[v] & (NODE_VALID | NODE_LEADER)) != NODE_VALID) continue;
            const unsigned long long s = __ldcg(a.st_b + v);
            const unsigned long long s2 = __ldcg(a.st_b + lo32(s));
            __stcg(a.st_b + v, pack(lo32(s2), hi32(s) + hi32(s2)));
        }
        grid.sync();
    }
    // ---- C. line lengths, offsets
    for (int v = tid; v < NS; v += nth) {
        if (!(a.info[v] & NODE_VALID)) continue;
        const unsigned long long s = __ldcg(a.st_b + v);
        const int lead = lo32(s);
        if (!(a.info[lead] & NODE_LEADER)) err |= ERR_RANK;
        const int t = a.succ[v];
        if (t == -1 || t == lead) a.len[lead] = hi32(s) + 2u;           // last node of its line: nodes + 1 vertices
    }
    grid.sync();
    const int lo = static_cast<int>(static_cast<long long>(NS) * blockIdx.x / gridDim.x);
    const int hi = static_cast<int>(static_cast<long long>(NS) * (blockIdx.x + 1) / gridDim.x);
    {
        unsigned long long sums[4] = {0, 0, 0, 0};
        for (int v = lo + threadIdx.x; v < hi; v += blockDim.x) {
            const unsigned f = a.info[v];
            if (!(f & NODE_LEADER)) continue;
            const int c = (f & NODE_HEAD) ? 0 : 2;
            sums[c] += __ldcg(a.len + v);
            sums[c + 1] += 1;
        }
        if (threadIdx.x < 4) s_red[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned long long x = sums[c];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
            if ((threadIdx.x & 31) == 0 && x) atomicAdd(&s_red[c], x);
        }
        __syncthreads();
        if (threadIdx.x < 4) __stcg(a.part + 4 * blockIdx.x + threadIdx.x, s_red[threadIdx.x]);
    }
    grid.sync();
    unsigned long long before[4], total[4];
    {
        unsigned long long b4[4] = {0, 0, 0, 0}, t4[4] = {0, 0, 0, 0};
        for (unsigned g = threadIdx.x; g < gridDim.x; g += blockDim.x) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const unsigned long long x = __ldcg(a.part + 4 * g + c);
                t4[c] += x;
                if (g < blockIdx.x) b4[c] += x;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned long long dummy;
            // block_exscan as a reduction: the total is all we need
            block_exscan(b4[c], s_warp, before[c]);
            block_exscan(t4[c], s_warp, total[c]);
            (void)dummy;
        }
    }
    const unsigned long long n_vertices = total[0] + total[2], n_lines = total[1] + total[3];
    // open lines occupy [0, total[0]) vertices and [0, total[1]) line slots, loops follow
    unsigned long long carry_v[2] = {before[0], total[0] + before[2]};
    unsigned long long carry_l[2] = {before[1], total[1] + before[3]};
    for (int base = lo; base < hi; base += blockDim.x) {
        const int v = base + threadIdx.x;
        unsigned f = 0u; unsigned ln = 0u;
        if (v < hi) { f = a.info[v]; if (f & NODE_LEADER) ln = __ldcg(a.len + v); else f = 0u; }
        const bool open = (f & NODE_HEAD) != 0u, loop = f && !open;
        unsigned long long tv0, tv1, tl;
        const unsigned long long ev0 = block_exscan(open ? ln : 0u, s_warp, tv0);
        const unsigned long long ev1 = block_exscan(loop ? ln : 0u, s_warp, tv1);
        const unsigned long long el = block_exscan((open ? 1ull : 0ull) | (loop ? (1ull << 32) : 0ull), s_warp, tl);
        if (f) {
            const unsigned long long voff = open ? carry_v[0] + ev0 : carry_v[1] + ev1;
            const unsigned long long lidx = open ? carry_l[0] + (el & 0xffffffffull) : carry_l[1] + (el >> 32);
            const bool nstart = loop && ((f >> 2) & 3u) == EDGE_N;
            a.lineoff[v] = voff | (nstart ? (1ull << 63) : 0ull);
            a.offsets[lidx] = static_cast<long long>(voff);
        }
        carry_v[0] += tv0; carry_v[1] += tv1;
        carry_l[0] += tl & 0xffffffffull; carry_l[1] += tl >> 32;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.offsets[n_lines] = static_cast<long long>(n_vertices);
        a.result[0] = n_vertices;
        a.result[1] = n_lines;
    }
    grid.sync();
    // ---- D. vertices.  For consistent records every position is < n_vertices <= 2 * NS (the allocated vertex slots);
    // the bound is checked anyway, so that malformed input (flagged above, LM_E_INVALID) cannot write out of bounds.
    const unsigned long long vcap = 2ull * static_cast<unsigned long long>(NS);
    for (int v = tid; v < NS; v += nth) {
        const unsigned f = a.info[v];
        if (!(f & NODE_VALID)) continue;
        const unsigned long long s = __ldcg(a.st_b + v);
        const int lead = lo32(s);
        const unsigned long long lw = __ldcg(a.lineoff + lead);
        const bool nstart = (lw >> 63) != 0ull;
        const unsigned long long b = lw & ~(1ull << 63);
        const int k = v >> 1, sg = v & 1;
        const longlong2 xyb = __ldg(reinterpret_cast<const longlong2*>(a.rec + static_cast<long long>(k) * REC_WORDS + 4 + 2 * sg));
        const double2 xy = make_double2(__longlong_as_double(xyb.x), __longlong_as_double(xyb.y));
        const unsigned long long pos = b + hi32(s) + (nstart ? 0u : 1u);
        if (pos < vcap) a.verts[pos] = xy;
        if (v == lead) {
            if (nstart) {
                const unsigned long long last = b + __ldcg(a.len + v) - 1u;
                if (last < vcap) a.verts[last] = xy;                    // the loop closes on its first emitted vertex
            } else if (b < vcap) {
                // initial vertex: interp(edge start, edge end) on the ENTRY edge, in the reference's operation order
                const int en = static_cast<int>((f >> 2) & 3u);
                int dj1, di1, dj2, di2;
                switch (en) {
                    case EDGE_E: dj1 = 0; di1 = 1; dj2 = 1; di2 = 1; break;   // SE -> NE
                    case EDGE_N: dj1 = 1; di1 = 1; dj2 = 1; di2 = 0; break;   // NE -> NW
                    case EDGE_W: dj1 = 1; di1 = 0; dj2 = 0; di2 = 0; break;   // NW -> SW
                    default:     dj1 = 0; di1 = 0; dj2 = 0; di2 = 1; break;   // SW -> SE
                }
                const long long q = __ldg(a.rec + static_cast<long long>(k) * REC_WORDS);
                const long long j = q / a.nx, i = q - j * a.nx;
                const unsigned long long c1 = static_cast<unsigned long long>(__ldg(a.rec + static_cast<long long>(k) * REC_WORDS + 1 + dj1));
                const unsigned long long c2 = static_cast<unsigned long long>(__ldg(a.rec + static_cast<long long>(k) * REC_WORDS + 1 + dj2));
                const double z1 = static_cast<double>(static_cast<int>(di1 ? (c1 >> 32) : (c1 & 0xffffffffull)));
                const double z2 = static_cast<double>(static_cast<int>(di2 ? (c2 >> 32) : (c2 & 0xffffffffull)));
                const double fr = __ddiv_rn(__dsub_rn(z2, a.level), __dsub_rn(z2, z1));
                const double g = __dsub_rn(1.0, fr);
                const double vx = __dadd_rn(__dmul_rn(__ldg(a.xs + i + di1), fr), __dmul_rn(__ldg(a.xs + i + di2), g));
                const double vy = __dadd_rn(__dmul_rn(__ldg(a.ys + j + dj1), fr), __dmul_rn(__ldg(a.ys + j + dj2), g));
                a.verts[b] = make_double2(vx, vy);
            }
        }
    }
    if (err) atomicOr(a.counters, err);
}

// ---- the linked lines of the most recent call, resident on the device (exported by lm::contour_export) ----
struct Linked {
    int dev = -1;
    const double* verts = nullptr;      // [nv][2]
    const long long* offsets = nullptr; // [nl + 1]
    long long nv = 0, nl = -1;          // nl < 0: nothing pending
};
Linked g_linked;

size_t align_up(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

}  // namespace

namespace lm {

// records on the device -> polylines on the device (kept until the next contour call); *nv / *nl = sizes
int32_t contour_link_device(const long long* rec_dev, long long n, const double* xs_host, long long nx,
                            const double* ys_host, long long ny, double level, long long* nv, long long* nl,
                            float* kernel_ms, cudaStream_t s) {
    *nv = 0; *nl = 0;
    g_linked = Linked{};
    LM_REQUIRE(n >= 0 && n < (1LL << 30), "lm_contour_link: %lld crossing records (the device linker takes < 2^30)", n);
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    void* dout = nullptr;
    if (n == 0 || nx < 2 || ny < 2) {
        int32_t rc0 = ws_get(WS_LINK_OUT, 256, &dout);
        if (rc0 != LM_OK) return rc0;
        LM_CUDA_TRY(cudaMemsetAsync(dout, 0, 256, s));
        g_linked.dev = dev; g_linked.verts = static_cast<double*>(dout); g_linked.offsets = static_cast<long long*>(dout);
        g_linked.nv = 0; g_linked.nl = 0;
        return LM_OK;
    }
    const size_t NS = static_cast<size_t>(2 * n);
    int32_t rc;
    // persistent CTAs of the cooperative kernel
    static int per_sm[64] = {};
    const int di = dev < 64 ? dev : 63;
    if (per_sm[di] == 0) {
        int v = 0;
        LM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, link_rank_kernel, LINK_THREADS, 0));
        per_sm[di] = v < 1 ? 1 : v;
    }
    long long grid = static_cast<long long>(sm_count()) * per_sm[di];
    const long long want = (static_cast<long long>(NS) + LINK_THREADS - 1) / LINK_THREADS;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;

    // workspace carve-up
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align_up(bytes); return o; };
    const size_t o_succ = take(NS * 4), o_pred = take(NS * 4), o_info = take(NS), o_sta = take(NS * 8), o_stb = take(NS * 8);
    const size_t o_len = take(NS * 4), o_loff = take(NS * 8), o_part = take(static_cast<size_t>(grid) * 4 * 8);
    const size_t o_cnt = take(64), o_xs = take(static_cast<size_t>(nx) * 8), o_ys = take(static_cast<size_t>(ny) * 8);
    void* dws = nullptr;
    if ((rc = ws_get(WS_LINK, off, &dws)) != LM_OK) return rc;
    char* w = static_cast<char*>(dws);
    // output: a line with L nodes has L + 1 vertices, so vertices <= nodes + lines <= 2 * nodes
    const size_t o_verts = 0, o_offs = align_up(NS * 2 * 16), out_bytes = o_offs + (NS + 1) * 8;
    if ((rc = ws_get(WS_LINK_OUT, out_bytes, &dout)) != LM_OK) return rc;
    char* wo = static_cast<char*>(dout);

    LinkArgs a{};
    a.rec = rec_dev; a.n = static_cast<int>(n); a.nx = nx; a.ny = ny; a.level = level;
    a.xs = reinterpret_cast<double*>(w + o_xs); a.ys = reinterpret_cast<double*>(w + o_ys);
    int* succ = reinterpret_cast<int*>(w + o_succ); int* pred = reinterpret_cast<int*>(w + o_pred);
    a.succ = succ; a.pred = pred; a.info = reinterpret_cast<unsigned char*>(w + o_info);
    a.st_a = reinterpret_cast<unsigned long long*>(w + o_sta); a.st_b = reinterpret_cast<unsigned long long*>(w + o_stb);
    a.len = reinterpret_cast<unsigned*>(w + o_len); a.lineoff = reinterpret_cast<unsigned long long*>(w + o_loff);
    a.part = reinterpret_cast<unsigned long long*>(w + o_part);
    a.counters = reinterpret_cast<unsigned*>(w + o_cnt); a.result = reinterpret_cast<unsigned long long*>(w + o_cnt + 16);
    a.verts = reinterpret_cast<double2*>(wo + o_verts); a.offsets = reinterpret_cast<long long*>(wo + o_offs);

    LM_CUDA_TRY(cudaMemcpyAsync(w + o_xs, xs_host, static_cast<size_t>(nx) * 8, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(w + o_ys, ys_host, static_cast<size_t>(ny) * 8, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(pred, 0xff, NS * 4, s));
    LM_CUDA_TRY(cudaMemsetAsync(a.len, 0, NS * 4, s));
    LM_CUDA_TRY(cudaMemsetAsync(w + o_cnt, 0, 64, s));
    Timer tm;
    if (kernel_ms && (rc = tm.begin(s)) != LM_OK) return rc;
    link_build_kernel<<<static_cast<unsigned>((n + LINK_THREADS - 1) / LINK_THREADS), LINK_THREADS, 0, s>>>(
        rec_dev, static_cast<int>(n), nx, ny, succ, pred, a.info, a.counters);
    LM_CUDA_TRY(cudaGetLastError());
    void* kargs[] = {&a};
    LM_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(link_rank_kernel), dim3(static_cast<unsigned>(grid)),
                                            dim3(LINK_THREADS), kargs, 0, s));
    unsigned long long h[4] = {0, 0, 0, 0};            // counters[0..1] | result[0..1] live in one 32-byte block
    unsigned hc[4] = {0, 0, 0, 0};
    LM_CUDA_TRY(cudaMemcpyAsync(hc, a.counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaMemcpyAsync(h, a.result, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    if (kernel_ms) { if ((rc = tm.end(s, kernel_ms)) != LM_OK) return rc; }
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (hc[0])
        return fail(LM_E_INVALID, "lm_contour_link: inconsistent crossing records (flags 0x%x: 1 not in raster order / outside the grid, "
                    "2 a neighbour quad is missing, 4 a segment without predecessor, 8 ranking failed, 16 bad segment count)", hc[0]);
    g_linked.dev = dev; g_linked.verts = reinterpret_cast<const double*>(a.verts); g_linked.offsets = a.offsets;
    g_linked.nv = static_cast<long long>(h[0]); g_linked.nl = static_cast<long long>(h[1]);
    *nv = g_linked.nv; *nl = g_linked.nl;
    return LM_OK;
}

// copy the device-resident lines of the last link to the caller; LM_E_CAP (sizes reported) if they do not fit
int32_t contour_export(double* verts, long long cap_verts, long long* n_verts, long long* line_offsets,
                       long long cap_lines, long long* n_lines, cudaStream_t s, const char* who) {
    LM_REQUIRE(g_linked.nl >= 0, "%s: no linked contour on the device", who);
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    LM_REQUIRE(dev == g_linked.dev, "%s: the pending contour lives on device %d, current device is %d", who, g_linked.dev, dev);
    *n_verts = g_linked.nv; *n_lines = g_linked.nl;
    if (g_linked.nv > cap_verts || g_linked.nl > cap_lines)
        return fail(LM_E_CAP, "%s: need room for %lld vertices and %lld lines (got %lld, %lld); "
                    "call lm_contour_fetch_last with larger buffers", who, g_linked.nv, g_linked.nl, cap_verts, cap_lines);
    LM_REQUIRE(verts || g_linked.nv == 0, "%s: verts is NULL", who);
    if (g_linked.nv)
        LM_CUDA_TRY(cudaMemcpyAsync(verts, g_linked.verts, static_cast<size_t>(g_linked.nv) * 16, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaMemcpyAsync(line_offsets, g_linked.offsets, static_cast<size_t>(g_linked.nl + 1) * 8, cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    return LM_OK;
}

void contour_linked_forget() { g_linked = Linked{}; }

}  // namespace lm

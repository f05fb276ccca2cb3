// lm_contour.cu -- K2: level set of the dwell field (marching squares), HBM bound.
//
// Replaces plt.contour(xs, ys, Z, levels=[level]) as used by
//   extract_contour   mandelbrot_boundary_sample.py:41-54
//   extract_contour   mandelbrot_boundary_sample_spyder.py:35-43
// with contourpy's "mpl2014" line semantics: a point is above iff
// z > level; a vertex on edge p1->p2 is xy1*f + xy2*(1-f) with f = (z2-level)/(z2-z1);
// higher side on the left; saddles decided by the mean of the four corners; boundary lines
// first, then interior loops in raster order of their first quad.
//
// Pipeline (algorithmic bytes: one 4-byte read of every dwell value):
//   1. mark kernel   : every warp takes a 128-quad strip of one row pair, loads the two dwell
//                      rows coalesced, classifies the quads, writes a 1-bit-per-quad crossing
//                      mask (ballot words, raster order) and adds the strip's count to the
//                      row counter.  DRAM: 4 B/pixel read (+1/32 written); the second use of
//                      each row is served by L2.
//   2. scan kernel   : exclusive scan of the row counters (single CTA).
//   3. emit kernel   : warp per row walks the mask words, and for every crossing quad gathers
//                      the four corners, derives the line segment(s) through it and their
//                      exit vertices in the reference's exact arithmetic, and writes a 64-byte
//                      record at its raster-order rank (ordered compaction, no sort needed).
//   4. link          : lm_contour_link.cu chains the records (~0.1-0.4 % of the quads) into ordered
//                      polylines on the device (successor table, pointer-jumping list ranking, scan,
//                      scatter) following mpl2014's start / direction / saddle rules; only the finished
//                      lines are copied to the host.
//
// Record (8 x int64): [0] quad = j*nx + i (global row j), [1] SW | SE<<32, [2] NW | NE<<32
// (corner dwell, uint32 each), [3] meta, [4..5] exit vertex of segment 0 (x, y as doubles),
// [6..7] exit vertex of segment 1 (saddle quads only).
// meta: bits 0-3 config (NW<<3|NE<<2|SW<<1|SE), bit 4 saddle turns right (mean > level),
// bits 8-9 entry edge of segment 0, 10-11 its exit edge, 12-13 / 14-15 the same for
// segment 1, bits 16-17 number of segments.  Edges: E=0, N=1, W=2, S=3.
#include "lm_common.cuh"

#include <cuda/ptx>

#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STRIP = 128;                 // quads per warp step (4 ballot words)
constexpr int MARK_WARPS = 8;
constexpr int REC_WORDS = 8;               // int64 words per record

enum { EDGE_E = 0, EDGE_N = 1, EDGE_W = 2, EDGE_S = 3, EDGE_NONE = -1 };

__host__ __device__ inline bool above_level(int z, double level) { return static_cast<double>(z) > level; }

// exit edge when entering through `edge` (see lm_oracle_contour.c / mpl2014 follow_interior):
// dir: +1 left, 0 straight, -1 right
__host__ __device__ inline int exit_edge_of(int edge, int dir) {
    // left turn = next edge clockwise seen from inside:  E->S, N->E, W->N, S->W
    // straight = opposite edge;  right turn: E->N, N->W, W->S, S->E
    if (dir == 0) return (edge + 2) & 3;
    if (dir > 0) return (edge + 3) & 3;
    return (edge + 1) & 3;
}

// Segments through a quad.  Returns the number of segments (0, 1 or 2) and fills
// entry[]/exit[]; saddle_right = (mean of corners > level).
__host__ __device__ inline int quad_segments(bool sw, bool se, bool nw, bool ne, bool saddle_right,
                                             int entry[2], int exit_[2]) {
    int n = 0;
    // an entry edge has its start point above and its end point not above (edges are CCW:
    // E: SE->NE, N: NE->NW, W: NW->SW, S: SW->SE)
    const bool ent[4] = {se && !ne, ne && !nw, nw && !sw, sw && !se};
    for (int e = 3; e >= 0; --e) {           // S first so that segment 0 of a saddle is the S / W entry
        const int edge = (e == 3) ? EDGE_S : (e == 2) ? EDGE_W : (e == 1) ? EDGE_N : EDGE_E;
        if (!ent[edge]) continue;
        // far-left / far-right corners seen from the entry edge
        bool pl, pr;
        switch (edge) {
            case EDGE_E: pl = sw; pr = nw; break;
            case EDGE_N: pl = se; pr = sw; break;
            case EDGE_W: pl = ne; pr = se; break;
            default:     pl = nw; pr = ne; break;
        }
        int dir;
        if (!pl && pr) dir = saddle_right ? -1 : +1;        // saddle
        else if (!pl && !pr) dir = +1;
        else if (pl && pr) dir = -1;
        else dir = 0;
        if (n < 2) { entry[n] = edge; exit_[n] = exit_edge_of(edge, dir); }
        ++n;
    }
    return n < 2 ? n : 2;
}

// points of an edge of quad (j, i): returns (j1,i1) start and (j2,i2) end as offsets 0/1
__host__ __device__ inline void edge_corners(int edge, int& dj1, int& di1, int& dj2, int& di2) {
    switch (edge) {
        case EDGE_E: dj1 = 0; di1 = 1; dj2 = 1; di2 = 1; break;   // SE -> NE
        case EDGE_N: dj1 = 1; di1 = 1; dj2 = 1; di2 = 0; break;   // NE -> NW
        case EDGE_W: dj1 = 1; di1 = 0; dj2 = 0; di2 = 0; break;   // NW -> SW
        default:     dj1 = 0; di1 = 0; dj2 = 0; di2 = 1; break;   // SW -> SE
    }
}

// ------------------------------------------------------------------------------------
// 1. mark: crossing mask + row counts.  A CTA takes MARK_ROWS quad rows at a time
// (MARK_ROWS+1 dwell rows), its warps stride over 128-column strips.  "z > level" for an
// integer z is "z > floor(level)"; every 32-column chunk of a dwell row becomes one ballot
// word (coalesced 128-byte loads), and the crossing words follow from warp-uniform bit
// operations on those words:  quad (r, b) is crossed iff its four corner bits are not equal,
//   X = (L ^ L') | (H ^ H') | (L ^ H),  L' / H' = the row words shifted by one column
// (the funnel shift pulls in the first bit of the next chunk / the strip's edge column).
// ------------------------------------------------------------------------------------
// quad rows per tile: 8 rows re-read every 9th dwell row (measured on the 32768^2 grid, ncu: 4.50 GB read for 4.29 GB
// of dwell values, 0.735 ms = 0.90 of the measured HBM peak; 4 rows: 4.84 GB, 0.822 ms -- profiles/r02_k2_mark_rows_ab.txt)
#ifndef LM_K2_MARK_ROWS
#define LM_K2_MARK_ROWS 8
#endif
constexpr int MARK_ROWS = LM_K2_MARK_ROWS;

// one dwell value: through the read-only path from global memory, or from the staged tile in shared memory
template <bool SMEM>
__device__ __forceinline__ int dwell_at(const int* q) { return SMEM ? *q : __ldg(q); }

// `strip0` points at column c0 of dwell row j0 (in global memory with row stride nx, or in the staged
// shared-memory tile with its own row stride)
template <bool SAFE, bool SMEM>
__device__ __forceinline__ void mark_strip(const int* __restrict__ strip0, const long long stride,
                                           const long long nx, const long long ny,
                                           const long long j0, const long long c0, const int ilevel, const int lane,
                                           unsigned* __restrict__ mrow0, const long long words_per_row,
                                           unsigned (&cnt)[MARK_ROWS]) {
    unsigned M[MARK_ROWS + 1][4];
    unsigned E;                                       // bit r: corner right of the strip in dwell row j0 + r
    const int* p = strip0 + lane;
    if (SAFE) {
        int z[MARK_ROWS + 1][4];
#pragma unroll
        for (int r = 0; r <= MARK_ROWS; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) z[r][k] = dwell_at<SMEM>(p + r * stride + 32 * k);
        const int ze = (lane <= MARK_ROWS) ? dwell_at<SMEM>(strip0 + lane * stride + STRIP) : INT_MIN;
#pragma unroll
        for (int r = 0; r <= MARK_ROWS; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) M[r][k] = __ballot_sync(FULL, z[r][k] > ilevel);
        E = __ballot_sync(FULL, ze > ilevel);
    } else {
#pragma unroll
        for (int r = 0; r <= MARK_ROWS; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool in = (j0 + r < ny) && (c0 + 32 * k + lane < nx);
                const int z = in ? dwell_at<SMEM>(p + r * stride + 32 * k) : INT_MIN;
                M[r][k] = __ballot_sync(FULL, z > ilevel);
            }
        const bool ein = (lane <= MARK_ROWS) && (j0 + lane < ny) && (c0 + STRIP < nx);
        const int ze = ein ? dwell_at<SMEM>(strip0 + lane * stride + STRIP) : INT_MIN;
        E = __ballot_sync(FULL, ze > ilevel);
    }
    // row words shifted by one column
    unsigned S[MARK_ROWS + 1][4];
#pragma unroll
    for (int r = 0; r <= MARK_ROWS; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) S[r][k] = __funnelshift_r(M[r][k], M[r][k + 1], 1);
        S[r][3] = __funnelshift_r(M[r][3], (E >> r) & 1u, 1);
    }
    unsigned* mp = mrow0 + (c0 >> 5);
#pragma unroll
    for (int r = 0; r < MARK_ROWS; ++r) {
        if (!SAFE && j0 + r >= ny - 1) break;
        uint4 w;
        unsigned* wp = &w.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned x = (M[r][k] ^ S[r][k]) | (M[r + 1][k] ^ S[r + 1][k]) | (M[r][k] ^ M[r + 1][k]);
            if (!SAFE) {
                // quads exist for columns < nx - 1
                const long long left = (nx - 1) - (c0 + 32 * k);
                x &= (left >= 32) ? 0xffffffffu : ((left <= 0) ? 0u : ((1u << left) - 1u));
            }
            wp[k] = x;
        }
        if (lane == 0) *reinterpret_cast<uint4*>(mp + r * words_per_row) = w;
        cnt[r] += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
}

// Work order: consecutive CTAs take consecutive 8-strip chunks of the SAME row group, so at any
// moment the whole chip streams through a few neighbouring dwell rows (long contiguous DRAM
// runs) instead of thousands of separate rows.
__global__ void __launch_bounds__(MARK_WARPS * 32) contour_mark_kernel(
    const int* __restrict__ dwell, long long nx, long long ny, int ilevel,
    unsigned* __restrict__ mask, long long words_per_row, unsigned* __restrict__ row_count) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long strips_per_row = words_per_row / 4;
    const long long ngroups = (ny - 1 + MARK_ROWS - 1) / MARK_ROWS;
    const long long chunks_per_group = (strips_per_row + MARK_WARPS - 1) / MARK_WARPS;
    const long long nitems = ngroups * chunks_per_group;
    const long long safe_strips = (nx - 1) / STRIP;          // strips with c0 + STRIP <= nx - 1
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        long long grp, chunk;
        if ((nitems >> 31) == 0) {
            const unsigned g32 = static_cast<unsigned>(item) / static_cast<unsigned>(chunks_per_group);
            grp = g32;
            chunk = static_cast<unsigned>(item) - g32 * static_cast<unsigned>(chunks_per_group);
        } else {
            grp = item / chunks_per_group;
            chunk = item - grp * chunks_per_group;
        }
        const long long sidx = chunk * MARK_WARPS + warp;
        if (sidx >= strips_per_row) continue;
        const long long j0 = grp * MARK_ROWS;
        const int* row0 = dwell + j0 * nx;
        unsigned* mrow0 = mask + j0 * words_per_row;
        unsigned cnt[MARK_ROWS];
#pragma unroll
        for (int r = 0; r < MARK_ROWS; ++r) cnt[r] = 0u;
        if (j0 + MARK_ROWS < ny && sidx < safe_strips)
            mark_strip<true, false>(row0 + sidx * STRIP, nx, nx, ny, j0, sidx * STRIP, ilevel, lane, mrow0, words_per_row, cnt);
        else
            mark_strip<false, false>(row0 + sidx * STRIP, nx, nx, ny, j0, sidx * STRIP, ilevel, lane, mrow0, words_per_row, cnt);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < MARK_ROWS; ++r)
                if (cnt[r]) atomicAdd(row_count + j0 + r, cnt[r]);
        }
    }
}

// The same pass with the dwell rows brought in by the bulk-copy engine (cp.async.bulk -> UBLKCP) instead of
// per-thread loads, and with the bit arithmetic spread over the lanes.  ncu showed the register-load kernel to be
// instruction-issue bound (~330 warp instructions per 4 x 128-quad strip, every lane repeating the same
// warp-uniform XOR/OR/popc work), not memory bound.  Here
//   * one elected thread arms a "full" mbarrier with the tile's byte count and issues MARK_ROWS+1 row copies of
//     (MARK_WARPS*STRIP + 4) ints into one of two shared-memory buffers; a warp releases a buffer through an
//     "empty" mbarrier as soon as its ballots are done, so there is no CTA-wide barrier in the loop;
//   * the 20 ballot words (+ the edge bits) of a strip go to a 5 x 5 word scratch in shared memory and lane
//     l < 16 alone derives crossing word (row l/4, word l%4): one coalesced 64-byte-per-row store and a
//     4-lane shuffle reduction for the row counts.
// Needs 16-byte aligned rows: nx % 4 == 0 and a 16-byte aligned grid (the driver falls back otherwise).
constexpr int BULK_COLS = MARK_WARPS * STRIP + 4;           // quad columns of a tile + the right edge (padded to 16 B)
constexpr int BULK_STAGE_INTS = (MARK_ROWS + 1) * BULK_COLS;
constexpr int BULK_SCRATCH = (MARK_ROWS + 1) * 5 + 3;       // per warp: M[r][0..3] and the edge bit as a fifth word

template <bool SAFE>
__device__ __forceinline__ void mark_strip_tile(const int* __restrict__ strip0, unsigned* __restrict__ sm,
                                                uint64_t* empty_bar,
                                                const long long nx, const long long ny, const long long j0, const long long c0,
                                                const int ilevel, const int lane,
                                                unsigned* __restrict__ mrow0, const long long words_per_row,
                                                unsigned* __restrict__ row_count) {
    const int* p = strip0 + lane;
#pragma unroll
    for (int r = 0; r <= MARK_ROWS; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool in = SAFE || ((j0 + r < ny) && (c0 + 32 * k + lane < nx));
            const int z = in ? p[r * BULK_COLS + 32 * k] : INT_MIN;
            const unsigned b = __ballot_sync(FULL, z > ilevel);
            if (lane == ((r * 4 + k) & 31)) sm[r * 5 + k] = b;    // any lane holds the (warp-uniform) ballot
        }
    }
    {
        const bool ein = (lane <= MARK_ROWS) && (SAFE || ((j0 + lane < ny) && (c0 + STRIP < nx)));
        const int ze = ein ? strip0[lane * BULK_COLS + STRIP] : INT_MIN;
        if (lane <= MARK_ROWS) sm[lane * 5 + 4] = (ze > ilevel) ? 1u : 0u;
    }
    __syncwarp();
    if (lane == 0) cuda::ptx::mbarrier_arrive(empty_bar);         // this warp no longer reads the tile
    unsigned x = 0u;
    const int r = lane >> 2, k = lane & 3;
    if (lane < 4 * MARK_ROWS && (SAFE || j0 + r < ny - 1)) {
        const unsigned a = sm[r * 5 + k], an = sm[r * 5 + k + 1];
        const unsigned b = sm[(r + 1) * 5 + k], bn = sm[(r + 1) * 5 + k + 1];
        const unsigned sa = __funnelshift_r(a, an, 1), sb = __funnelshift_r(b, bn, 1);   // the rows shifted by one column
        x = (a ^ sa) | (b ^ sb) | (a ^ b);
        if (!SAFE) {
            const long long left = (nx - 1) - (c0 + 32 * k);       // quads exist for columns < nx - 1
            x &= (left >= 32) ? 0xffffffffu : ((left <= 0) ? 0u : ((1u << left) - 1u));
        }
        mrow0[r * words_per_row + (c0 >> 5) + k] = x;
    }
    unsigned cnt = __popc(x);
    cnt += __shfl_xor_sync(FULL, cnt, 1);
    cnt += __shfl_xor_sync(FULL, cnt, 2);
    if (k == 0 && lane < 4 * MARK_ROWS && cnt) atomicAdd(row_count + j0 + r, cnt);
    __syncwarp();                                                  // scratch is reused by the next strip
}

__global__ void __launch_bounds__(MARK_WARPS * 32) contour_mark_bulk_kernel(
    const int* __restrict__ dwell, long long nx, long long ny, int ilevel,
    unsigned* __restrict__ mask, long long words_per_row, unsigned* __restrict__ row_count) {
    static_assert(4 * MARK_ROWS <= 32, "one lane per crossing word");
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    int* tile = reinterpret_cast<int*>(bulk_smem);                                   // 2 stages
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bulk_smem + 2 * BULK_STAGE_INTS * sizeof(int));
    uint64_t* empty_bar = full_bar + 2;
    unsigned* scratch = reinterpret_cast<unsigned*>(empty_bar + 2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* sm = scratch + warp * BULK_SCRATCH;
    const long long strips_per_row = words_per_row / 4;
    const long long ngroups = (ny - 1 + MARK_ROWS - 1) / MARK_ROWS;
    const long long chunks_per_group = (strips_per_row + MARK_WARPS - 1) / MARK_WARPS;
    const long long nitems = ngroups * chunks_per_group;
    const long long safe_strips = (nx - 1) / STRIP;

    if (threadIdx.x == 0) {
        cuda::ptx::mbarrier_init(&full_bar[0], 1);
        cuda::ptx::mbarrier_init(&full_bar[1], 1);
        cuda::ptx::mbarrier_init(&empty_bar[0], MARK_WARPS);
        cuda::ptx::mbarrier_init(&empty_bar[1], MARK_WARPS);
        cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);       // make the initialised barriers visible to the copy engine
    }
    __syncthreads();

    // item -> (row group, column chunk); 32-bit division whenever the item count allows it
    const bool small = (nitems >> 31) == 0;
    auto split = [&](long long item, long long& grp, long long& chunk) {
        if (small) {
            const unsigned g32 = static_cast<unsigned>(item) / static_cast<unsigned>(chunks_per_group);
            grp = g32;
            chunk = static_cast<unsigned>(item) - g32 * static_cast<unsigned>(chunks_per_group);
        } else {
            grp = item / chunks_per_group;
            chunk = item - grp * chunks_per_group;
        }
    };
    unsigned empty_phase[2] = {1u, 1u};                   // a fresh barrier passes a wait on parity 1: both buffers start free
    auto issue = [&](long long item, int stage) {        // elected thread only
        while (!cuda::ptx::mbarrier_try_wait_parity(&empty_bar[stage], empty_phase[stage])) {}   // all warps released it
        empty_phase[stage] ^= 1u;
        long long grp, chunk;
        split(item, grp, chunk);
        const long long j0 = grp * MARK_ROWS, c0 = chunk * (MARK_WARPS * STRIP);
        const long long cols = (nx - c0 < BULK_COLS) ? nx - c0 : BULK_COLS;          // multiple of 4 (nx % 4 == 0)
        const long long rows = (ny - j0 < MARK_ROWS + 1) ? ny - j0 : MARK_ROWS + 1;
        const uint32_t row_bytes = static_cast<uint32_t>(cols * sizeof(int));
        cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared, &full_bar[stage],
                                             row_bytes * static_cast<uint32_t>(rows));
        int* dst = tile + stage * BULK_STAGE_INTS;
        for (long long r = 0; r < rows; ++r)
            cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, dst + r * BULK_COLS,
                                     dwell + (j0 + r) * nx + c0, row_bytes, &full_bar[stage]);
    };

    unsigned full_phase[2] = {0u, 0u};
    int stage = 0;
    long long item = blockIdx.x;
    if (item < nitems && threadIdx.x == 0) issue(item, 0);
    for (; item < nitems; item += gridDim.x, stage ^= 1) {
        const long long next = item + gridDim.x;
        if (next < nitems && threadIdx.x == 0) issue(next, stage ^ 1);
        while (!cuda::ptx::mbarrier_try_wait_parity(&full_bar[stage], full_phase[stage])) {}
        full_phase[stage] ^= 1u;
        long long grp, chunk;
        split(item, grp, chunk);
        const long long sidx = chunk * MARK_WARPS + warp;
        if (sidx < strips_per_row) {
            const long long j0 = grp * MARK_ROWS;
            unsigned* mrow0 = mask + j0 * words_per_row;
            const int* strip0 = tile + stage * BULK_STAGE_INTS + warp * STRIP;
            if (j0 + MARK_ROWS < ny && sidx < safe_strips)
                mark_strip_tile<true>(strip0, sm, &empty_bar[stage], nx, ny, j0, sidx * STRIP, ilevel, lane, mrow0, words_per_row, row_count);
            else
                mark_strip_tile<false>(strip0, sm, &empty_bar[stage], nx, ny, j0, sidx * STRIP, ilevel, lane, mrow0, words_per_row, row_count);
        } else if (lane == 0) {
            cuda::ptx::mbarrier_arrive(&empty_bar[stage]);         // nothing to read for this warp: release at once
        }
    }
}

// ------------------------------------------------------------------------------------
// 2. exclusive scan of the row counters (one CTA); total -> row_offset[nrows]
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) contour_scan_kernel(const unsigned* __restrict__ row_count, long long nrows,
                                                            unsigned long long* __restrict__ row_offset) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < nrows; base += 1024) {
        const long long idx = base + threadIdx.x;
        const unsigned long long v = idx < nrows ? row_count[idx] : 0ULL;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;           // inclusive over warps
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        const unsigned long long before = (warp ? warp_sums[warp - 1] : 0ULL);
        if (idx < nrows) row_offset[idx] = carry + before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_offset[nrows] = carry_s;
}

// vertex on `edge` of quad (j, i): interp(point1 = edge start, point2 = edge end)
__device__ __forceinline__ void edge_vertex_dev(int edge, const int z[2][2], const double* __restrict__ xs,
                                                const double* __restrict__ ys, long long j, long long i,
                                                double level, double& vx, double& vy) {
    int dj1, di1, dj2, di2;
    edge_corners(edge, dj1, di1, dj2, di2);
    const double z1 = static_cast<double>(z[dj1][di1]), z2 = static_cast<double>(z[dj2][di2]);
    const double f = __ddiv_rn(__dsub_rn(z2, level), __dsub_rn(z2, z1));
    const double g = __dsub_rn(1.0, f);
    vx = __dadd_rn(__dmul_rn(__ldg(xs + i + di1), f), __dmul_rn(__ldg(xs + i + di2), g));
    vy = __dadd_rn(__dmul_rn(__ldg(ys + j + dj1), f), __dmul_rn(__ldg(ys + j + dj2), g));
}

// ------------------------------------------------------------------------------------
// 3. emit: ordered records
// ------------------------------------------------------------------------------------
// one crossing quad (j, i) -> its record at rank `pos`
__device__ __forceinline__ void emit_record(const int* __restrict__ dwell, long long nx, long long row_offset_global, double level,
                                            const double* __restrict__ xs, const double* __restrict__ ys, long long j, long long i,
                                            unsigned long long pos, long long* __restrict__ records) {
    int z[2][2];
    z[0][0] = __ldg(dwell + j * nx + i);
    z[0][1] = __ldg(dwell + j * nx + i + 1);
    z[1][0] = __ldg(dwell + (j + 1) * nx + i);
    z[1][1] = __ldg(dwell + (j + 1) * nx + i + 1);
    const bool sw = above_level(z[0][0], level), se = above_level(z[0][1], level);
    const bool nw = above_level(z[1][0], level), ne = above_level(z[1][1], level);
    // mean of the four corners, summed in the reference's order SW+SE+NW+NE
    const double zmid = __dmul_rn(0.25, __dadd_rn(__dadd_rn(__dadd_rn(static_cast<double>(z[0][0]),
                        static_cast<double>(z[0][1])), static_cast<double>(z[1][0])), static_cast<double>(z[1][1])));
    const bool saddle_right = zmid > level;
    int entry[2] = {0, 0}, exit_[2] = {0, 0};
    const int nseg = quad_segments(sw, se, nw, ne, saddle_right, entry, exit_);
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (nseg >= 1) edge_vertex_dev(exit_[0], z, xs, ys, j, i, level, v[0], v[1]);
    if (nseg >= 2) edge_vertex_dev(exit_[1], z, xs, ys, j, i, level, v[2], v[3]);
    const unsigned config = (nw ? 8u : 0u) | (ne ? 4u : 0u) | (sw ? 2u : 0u) | (se ? 1u : 0u);
    const long long meta = static_cast<long long>(config | (saddle_right ? 16u : 0u) |
                           (static_cast<unsigned>(entry[0]) << 8) | (static_cast<unsigned>(exit_[0]) << 10) |
                           (static_cast<unsigned>(entry[1]) << 12) | (static_cast<unsigned>(exit_[1]) << 14) |
                           (static_cast<unsigned>(nseg) << 16));
    longlong2* r2 = reinterpret_cast<longlong2*>(records + pos * REC_WORDS);
    r2[0] = make_longlong2((row_offset_global + j) * nx + i,
                           static_cast<long long>(static_cast<unsigned>(z[0][0])) |
                           (static_cast<long long>(static_cast<unsigned>(z[0][1])) << 32));
    r2[1] = make_longlong2(static_cast<long long>(static_cast<unsigned>(z[1][0])) |
                           (static_cast<long long>(static_cast<unsigned>(z[1][1])) << 32), meta);
    r2[2] = make_longlong2(__double_as_longlong(v[0]), __double_as_longlong(v[1]));
    r2[3] = make_longlong2(__double_as_longlong(v[2]), __double_as_longlong(v[3]));
}

// A warp owns a quad row.  The row's mask words are fetched 32 x 32 at a time (32 independent coalesced
// loads in flight), the set bits are turned into a list of crossing columns in shared memory (raster order
// by ballot-free prefix sums), and the list is worked off 32 crossings at a time -- every lane gathers and
// emits one record in parallel, instead of one lane walking its word while 31 wait for its DRAM round trips.
constexpr int EMIT_LIST = 256;                 // list capacity per warp; flushed when more than EMIT_LIST - 64 are pending
__global__ void __launch_bounds__(MARK_WARPS * 32) contour_emit_kernel(
    const int* __restrict__ dwell, long long nx, long long ny, long long row_offset_global, double level,
    const double* __restrict__ xs, const double* __restrict__ ys,   // ys indexed by local row
    const unsigned* __restrict__ mask, long long words_per_row, const unsigned* __restrict__ row_count,
    const unsigned long long* __restrict__ row_offset, long long* __restrict__ records) {
    __shared__ unsigned s_words[MARK_WARPS][32 * 32];
    __shared__ int s_list[MARK_WARPS][EMIT_LIST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nwarps = static_cast<long long>(gridDim.x) * MARK_WARPS;
    for (long long j = static_cast<long long>(blockIdx.x) * MARK_WARPS + warp; j < ny - 1; j += nwarps) {
        if (row_count[j] == 0u) continue;
        unsigned long long list_pos = row_offset[j];      // rank of s_list[0]
        int nlist = 0;                                    // pending crossings (warp-uniform)
        auto flush = [&]() {
            __syncwarp();
            for (int b = 0; b < nlist; b += 32)
                if (b + lane < nlist)
                    emit_record(dwell, nx, row_offset_global, level, xs, ys, j, s_list[warp][b + lane], list_pos + b + lane, records);
            list_pos += static_cast<unsigned long long>(nlist);
            nlist = 0;
            __syncwarp();
        };
        for (long long s0 = 0; s0 < words_per_row; s0 += 32 * 32) {
            {
                unsigned wv[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const long long w = s0 + t * 32 + lane;
                    wv[t] = (w < words_per_row) ? __ldg(mask + j * words_per_row + w) : 0u;
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 32; ++t) s_words[warp][t * 32 + lane] = wv[t];
                __syncwarp();
            }
            for (int t = 0; t < 32; ++t) {
                if (s0 + t * 32 >= words_per_row) break;
                const long long w = s0 + t * 32 + lane;
                unsigned word = s_words[warp][t * 32 + lane];
                if (__ballot_sync(FULL, word != 0u) == 0u) continue;
                const int cnt = __popc(word);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += y;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                if (total > 64) {
                    // a dense stretch (up to 1024 crossings in 32 words): keep the order by flushing the list, then
                    // let every lane walk its own word
                    flush();
                    unsigned long long pos = list_pos + static_cast<unsigned long long>(incl - cnt);
                    while (word) {
                        const int bit = __ffs(word) - 1;
                        word &= word - 1;
                        emit_record(dwell, nx, row_offset_global, level, xs, ys, j, w * 32 + bit, pos, records);
                        ++pos;
                    }
                    list_pos += static_cast<unsigned long long>(total);
                    continue;
                }
                int k = nlist + incl - cnt;
                while (word) {
                    const int bit = __ffs(word) - 1;
                    word &= word - 1;
                    s_list[warp][k++] = static_cast<int>(w * 32 + bit);
                }
                nlist += total;
                if (nlist > EMIT_LIST - 64) flush();
            }
        }
        flush();
    }
}

// ------------------------------------------------------------------------------------
// device side driver: dwell block on the device -> raster-ordered records on the device
// ------------------------------------------------------------------------------------
// The records land in dev_dst when the caller gives a device buffer that is large enough, otherwise in the
// library's own workspace; *records_out points at them.  Synchronises s once (the record count sizes the buffer).
int32_t classify_device(const int* dwell_dev, const double* xs_host, long long nx,
                        const double* ys_host, long long ny, long long row_offset, double level,
                        const long long** records_out, long long* n_out, float* kernel_ms, int* launches, cudaStream_t s,
                        long long* dev_dst = nullptr, long long dev_cap = 0) {
    *records_out = nullptr; *n_out = 0;
    if (launches) *launches = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    if (nx < 2 || ny < 2) return LM_OK;
    const long long nrows = ny - 1;
    const long long words_per_row = ((nx - 1 + STRIP - 1) / STRIP) * 4;
    void *dmask, *dcount, *doff, *dxs, *dys;
    int32_t rc;
    if ((rc = lm::ws_get(lm::WS_K2_MASK, static_cast<size_t>(nrows) * words_per_row * sizeof(unsigned), &dmask)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_K2_COUNT, static_cast<size_t>(nrows) * sizeof(unsigned), &dcount)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_K2_OFFSET, static_cast<size_t>(nrows + 1) * sizeof(unsigned long long), &doff)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_K2_XS, static_cast<size_t>(nx) * sizeof(double), &dxs)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_K2_YS, static_cast<size_t>(ny) * sizeof(double), &dys)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dxs, xs_host, nx * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dys, ys_host, ny * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(dcount, 0, static_cast<size_t>(nrows) * sizeof(unsigned), s));

    lm::Timer tm;
    if (kernel_ms && (rc = tm.begin(s)) != LM_OK) return rc;
    // z > level  <=>  z > floor(level) for integer z (clamped to the int32 range)
    int ilevel;
    if (!(level >= -2147483648.0)) ilevel = INT_MIN;            // also NaN: nothing is above a NaN level
    else if (level >= 2147483647.0) ilevel = INT_MAX;
    else ilevel = static_cast<int>(floor(level));
    if (level != level) ilevel = INT_MAX;
    const long long cap = static_cast<long long>(lm::sm_count()) * 64;
    const long long strips_per_row = words_per_row / 4;
    long long blocks = ((nrows + MARK_ROWS - 1) / MARK_ROWS) * ((strips_per_row + MARK_WARPS - 1) / MARK_WARPS);
    if (blocks > cap) blocks = cap;
    static const bool no_bulk = getenv("LM_K2_NO_BULK") != nullptr;      // tuning / A-B switch
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    const int di = (dev >= 0 && dev < 64) ? dev : 63;
    if (!no_bulk && nx % 4 == 0 && (reinterpret_cast<uintptr_t>(dwell_dev) & 15u) == 0) {
        const size_t smem = 2 * BULK_STAGE_INTS * sizeof(int) + 4 * sizeof(uint64_t) + MARK_WARPS * BULK_SCRATCH * sizeof(unsigned);
        static int per_sm[64] = {};                                                  // per device: attribute set, CTAs that fit
        if (per_sm[di] == 0) {
            LM_CUDA_TRY(cudaFuncSetAttribute(contour_mark_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            int v = 0;
            LM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, contour_mark_bulk_kernel, MARK_WARPS * 32, smem));
            per_sm[di] = v < 1 ? 1 : v;
        }
        long long bblocks = blocks;
        const long long bcap = static_cast<long long>(lm::sm_count()) * per_sm[di];       // persistent: as many CTAs as fit
        if (bblocks > bcap) bblocks = bcap;
        contour_mark_bulk_kernel<<<static_cast<unsigned>(bblocks), MARK_WARPS * 32, smem, s>>>(
            dwell_dev, nx, ny, ilevel, static_cast<unsigned*>(dmask), words_per_row, static_cast<unsigned*>(dcount));
    } else {
        contour_mark_kernel<<<static_cast<unsigned>(blocks), MARK_WARPS * 32, 0, s>>>(
            dwell_dev, nx, ny, ilevel, static_cast<unsigned*>(dmask), words_per_row, static_cast<unsigned*>(dcount));
    }
    LM_CUDA_TRY(cudaGetLastError());
    contour_scan_kernel<<<1, 1024, 0, s>>>(static_cast<unsigned*>(dcount), nrows, static_cast<unsigned long long*>(doff));
    LM_CUDA_TRY(cudaGetLastError());
    if (launches) *launches = 2;
    unsigned long long total = 0;
    LM_CUDA_TRY(cudaMemcpyAsync(&total, static_cast<unsigned long long*>(doff) + nrows, sizeof(total),
                                cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    *n_out = static_cast<long long>(total);
    void* drec = nullptr;
    if (total) {
        if (dev_dst && static_cast<long long>(total) <= dev_cap) drec = dev_dst;
        else if ((rc = lm::ws_get(lm::WS_RECORDS, static_cast<size_t>(total) * REC_WORDS * sizeof(long long), &drec)) != LM_OK) return rc;
        long long eblocks = (nrows + MARK_WARPS - 1) / MARK_WARPS;
        if (eblocks > cap) eblocks = cap;
        contour_emit_kernel<<<static_cast<unsigned>(eblocks), MARK_WARPS * 32, 0, s>>>(
            dwell_dev, nx, ny, row_offset, level, static_cast<double*>(dxs), static_cast<double*>(dys),
            static_cast<unsigned*>(dmask), words_per_row, static_cast<unsigned*>(dcount),
            static_cast<unsigned long long*>(doff), static_cast<long long*>(drec));
        LM_CUDA_TRY(cudaGetLastError());
        if (launches) *launches = 3;
    }
    if (kernel_ms && (rc = tm.end(s, kernel_ms)) != LM_OK) return rc;
    *records_out = static_cast<const long long*>(drec);
    return LM_OK;
}

int32_t check_contour_args(const char* who, const void* dwell, const void* xs, int64_t nx, const void* ys, int64_t ny,
                           const void* verts, int64_t cap_verts, const void* n_verts, const void* offs,
                           int64_t cap_lines, const void* n_lines) {
    LM_REQUIRE(dwell && xs && ys, "%s: NULL input", who);
    LM_REQUIRE(nx >= 0 && ny >= 0, "%s: negative grid size", who);
    LM_REQUIRE(n_verts && n_lines, "%s: n_verts / n_lines is NULL", who);
    LM_REQUIRE(cap_verts >= 0 && cap_lines >= 0, "%s: negative capacity", who);
    LM_REQUIRE((verts || cap_verts == 0) && offs, "%s: NULL output buffer", who);
    return LM_OK;
}

}  // namespace

namespace lm {
// K2 on a device-resident dwell grid: mark / scan / emit (records), build / rank (ordered lines), then the
// lines -- and only the lines -- go back to the host.
int32_t contour_device_to_host(const int32_t* dwell_dev, const double* xs_host, int64_t nx, const double* ys_host,
                               int64_t ny, double level, double* verts, int64_t cap_verts, int64_t* n_verts,
                               int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                               float* kernel_ms, int* launches, cudaStream_t s) {
    const long long* recs = nullptr;
    long long nrec = 0;
    float ms_a = 0.f, ms_b = 0.f;
    int la = 0;
    int32_t rc = classify_device(dwell_dev, xs_host, nx, ys_host, ny, 0, level, &recs, &nrec, kernel_ms ? &ms_a : nullptr, &la, s);
    if (rc != LM_OK) return rc;
    long long nv = 0, nl = 0;
    rc = contour_link_device(recs, nrec, xs_host, nx, ys_host, ny, level, &nv, &nl, kernel_ms ? &ms_b : nullptr, s);
    if (rc != LM_OK) return rc;
    if (kernel_ms) *kernel_ms = ms_a + ms_b;
    if (launches) *launches = la + (nrec ? 2 : 0);
    return contour_export(verts, cap_verts, reinterpret_cast<long long*>(n_verts), reinterpret_cast<long long*>(line_offsets),
                          cap_lines, reinterpret_cast<long long*>(n_lines), s, "lm_contour");
}
}  // namespace lm

extern "C" {

int32_t lm_contour_fetch_last(double* verts, int64_t cap_verts, int64_t* n_verts,
                              int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n_verts && n_lines && line_offsets, "lm_contour_fetch_last: NULL argument");
    LM_REQUIRE(cap_verts >= 0 && cap_lines >= 0, "lm_contour_fetch_last: negative capacity");
    return lm::contour_export(verts, cap_verts, reinterpret_cast<long long*>(n_verts), reinterpret_cast<long long*>(line_offsets),
                              cap_lines, reinterpret_cast<long long*>(n_lines), nullptr, "lm_contour_fetch_last");
}

int32_t lm_contour_classify_dev(const int32_t* dwell_dev, const double* xs_host, int64_t nx,
                                const double* ys_host, int64_t ny, int64_t row_offset, double level,
                                int64_t* records, int64_t cap_records, int64_t* n_records, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(dwell_dev && xs_host && ys_host && n_records, "lm_contour_classify_dev: NULL argument");
    LM_REQUIRE(nx >= 0 && ny >= 0 && cap_records >= 0, "lm_contour_classify_dev: negative size");
    const long long* recs = nullptr;
    long long n = 0;
    cudaStream_t s = lm::as_stream(stream);
    rc = classify_device(dwell_dev, xs_host, nx, ys_host, ny, row_offset, level, &recs, &n, nullptr, nullptr, s);
    if (rc != LM_OK) return rc;
    *n_records = n;
    if (n > cap_records)
        return lm::fail(LM_E_CAP, "lm_contour_classify_dev: need room for %lld records (got %lld)", n,
                        static_cast<long long>(cap_records));
    LM_REQUIRE(records || n == 0, "lm_contour_classify_dev: records is NULL");
    if (n) {
        LM_CUDA_TRY(cudaMemcpyAsync(records, recs, static_cast<size_t>(n) * REC_WORDS * sizeof(long long), cudaMemcpyDeviceToHost, s));
        LM_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return LM_OK;
}

int32_t lm_contour_records_dev(const int32_t* dwell_dev, const double* xs_host, int64_t nx,
                               const double* ys_host, int64_t ny, int64_t row_offset, double level,
                               int64_t* records_dev, int64_t cap_records, int64_t* n_records, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(dwell_dev && xs_host && ys_host && n_records, "lm_contour_records_dev: NULL argument");
    LM_REQUIRE(nx >= 0 && ny >= 0 && cap_records >= 0 && (records_dev || cap_records == 0), "lm_contour_records_dev: bad size / buffer");
    const long long* recs = nullptr;
    long long n = 0;
    cudaStream_t s = lm::as_stream(stream);
    rc = classify_device(dwell_dev, xs_host, nx, ys_host, ny, row_offset, level, &recs, &n, nullptr, nullptr, s,
                         reinterpret_cast<long long*>(records_dev), cap_records);
    if (rc != LM_OK) return rc;
    *n_records = n;
    if (n > cap_records)
        return lm::fail(LM_E_CAP, "lm_contour_records_dev: need room for %lld records (got %lld)", n,
                        static_cast<long long>(cap_records));
    return LM_OK;
}

int32_t lm_contour_link_dev(const int64_t* records_dev, int64_t n_records, const double* xs, int64_t nx,
                            const double* ys, int64_t ny, double level,
                            double* verts, int64_t cap_verts, int64_t* n_verts,
                            int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE((records_dev || n_records == 0) && xs && ys && n_verts && n_lines && line_offsets, "lm_contour_link_dev: NULL argument");
    LM_REQUIRE(n_records >= 0 && nx >= 0 && ny >= 0 && cap_verts >= 0 && cap_lines >= 0, "lm_contour_link_dev: negative size");
    cudaStream_t s = lm::as_stream(stream);
    long long nv = 0, nl = 0;
    rc = lm::contour_link_device(reinterpret_cast<const long long*>(records_dev), n_records, xs, nx, ys, ny, level, &nv, &nl, nullptr, s);
    if (rc != LM_OK) return rc;
    return lm::contour_export(verts, cap_verts, reinterpret_cast<long long*>(n_verts), reinterpret_cast<long long*>(line_offsets),
                              cap_lines, reinterpret_cast<long long*>(n_lines), s, "lm_contour_link_dev");
}

int32_t lm_contour_link(const int64_t* records, int64_t n_records, const double* xs, int64_t nx,
                        const double* ys, int64_t ny, double level,
                        double* verts, int64_t cap_verts, int64_t* n_verts,
                        int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE((records || n_records == 0) && xs && ys && n_verts && n_lines && line_offsets,
               "lm_contour_link: NULL argument");
    LM_REQUIRE(n_records >= 0 && nx >= 0 && ny >= 0, "lm_contour_link: negative size");
    void* drec = nullptr;
    if (n_records) {
        const size_t nb = static_cast<size_t>(n_records) * REC_WORDS * sizeof(long long);
        if ((rc = lm::ws_get(lm::WS_RECORDS, nb, &drec)) != LM_OK) return rc;
        LM_CUDA_TRY(cudaMemcpyAsync(drec, records, nb, cudaMemcpyHostToDevice, nullptr));
    }
    return lm_contour_link_dev(static_cast<const int64_t*>(drec), n_records, xs, nx, ys, ny, level, verts, cap_verts, n_verts,
                               line_offsets, cap_lines, n_lines, nullptr);
}

int32_t lm_contour_level_dev(const int32_t* dwell_dev, const double* xs_host, int64_t nx,
                             const double* ys_host, int64_t ny, double level,
                             double* verts, int64_t cap_verts, int64_t* n_verts,
                             int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines,
                             lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    rc = check_contour_args("lm_contour_level_dev", dwell_dev, xs_host, nx, ys_host, ny, verts, cap_verts, n_verts,
                            line_offsets, cap_lines, n_lines);
    if (rc != LM_OK) return rc;
    if (stats) *stats = lm_stats{};
    float ms = 0.f;
    int launches = 0;
    rc = lm::contour_device_to_host(dwell_dev, xs_host, nx, ys_host, ny, level, verts, cap_verts, n_verts, line_offsets,
                                    cap_lines, n_lines, &ms, &launches, nullptr);
    if (stats) {
        stats->items = static_cast<uint64_t>(nx) * static_cast<uint64_t>(ny);
        stats->work_units = stats->items * 4;           // algorithmic bytes: one int32 read per pixel
        stats->kernel_ms = ms;
        stats->launches = launches;
    }
    return rc;
}

int32_t lm_contour_level(const int32_t* dwell, const double* xs, int64_t nx, const double* ys, int64_t ny,
                         double level, double* verts, int64_t cap_verts, int64_t* n_verts,
                         int64_t* line_offsets, int64_t cap_lines, int64_t* n_lines, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    rc = check_contour_args("lm_contour_level", dwell, xs, nx, ys, ny, verts, cap_verts, n_verts, line_offsets,
                            cap_lines, n_lines);
    if (rc != LM_OK) return rc;
    if (nx == 0 || ny == 0) {
        *n_verts = 0; *n_lines = 0; line_offsets[0] = 0;
        if (stats) *stats = lm_stats{};
        return LM_OK;
    }
    void* dd = nullptr;
    const size_t nb = static_cast<size_t>(nx) * ny * sizeof(int32_t);
    if ((rc = lm::ws_get(lm::WS_OUT_I32, nb, &dd)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dd, dwell, nb, cudaMemcpyHostToDevice, nullptr));
    return lm_contour_level_dev(static_cast<const int32_t*>(dd), xs, nx, ys, ny, level, verts, cap_verts, n_verts,
                                line_offsets, cap_lines, n_lines, stats);
}

}  // extern "C"

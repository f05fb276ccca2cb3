// lm_escape.cu -- K1: fp64 escape-time kernels (grid and point list) for sm_100a.
//
// Reference semantics (bit-exact dwell):
//   mandelbrot_dwell / compute_grid      mandelbrot_boundary_sample.py:22-39
//   mandelbrot_parameter_potential       lucas_equipotential_test_v3.py:124-162
//   escape_potential                     Potentials.py:32-47
//   mandelbrot_potential                 Laplacian_C-M.py:27-43
//   escape_potential                     Iterative_Variogram_Laplacian.py:114-130
//   mandelbrot_escape_potential          variograms_construct_mandelbrot.py:148-167
//
// Design (B200-first, FP64-pipe bound -- nothing here is a contraction, so no tensor cores):
//   * persistent CTAs (one wave, grid = SMs x resident CTAs); every WARP pulls 128-pixel
//     row segments ("tiles") from a global atomic counter (work stealing) with the next
//     tile id prefetched one tile ahead;
//   * lane refill: a lane that finishes its pixel immediately takes the next pixel of
//     the warp's tile (ballot + popc rank), so divergence near the set boundary never
//     idles lanes -- all 32 lanes carry live orbits until the global queue drains;
//   * the recurrence uses __dmul_rn/__dadd_rn/__fma_rn only (and the file is built with
//     -fmad=false): a=zr*zr, b=zi*zi, p=zr*zi, zr'=(a-b)+cr, zi'=fma(2,p,ci) [== (p+p)+ci
//     exactly], test fl(a'+b') > bailout^2.  That is 6 FP64-pipe instructions for the
//     update + 1 DADD for the test;
//   * blind fast path: for |c| <= cmax (= R^2 - R - 0.05, i.e. 1.95 for R = 2) escape is
//     absorbing with a margin that dwarfs rounding error: |z_j|^2 = |z_{j+1} - c| <=
//     |z_{j+1}| + |c|, so if the LAST iterate of a block has fl(a+b) <= R^2 then by backward
//     induction every earlier iterate of the block had |z_j|^2 <= R + cmax < R^2 - 0.05 and
//     the reference's test could not have fired.  Blocks of FB iterations therefore run
//     with NO per-iteration test (6 FP64 instructions per iteration, the issue-port
//     minimum); one DADD + compare per block decides, and a block in which some lane did
//     escape is rolled back and re-run with the exact per-iteration test to find the
//     first-escape index.  Warps holding a pixel with |c| > cmax never go blind;
//   * results are staged per warp in a shared-memory ring of tiles and leave the SM as
//     one 128-bit coalesced store per lane per tile; a pixel that is still iterating
//     when its tile is evicted from the ring patches its own word later.
//
// Work unit: pixel_iters = sum over pixels of min(dwell+1, max_iter), counted exactly in
// the kernel (lm_stats.work_units).
#include "lm_common.cuh"

#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include <vector>

namespace {

// tunables (overridable with -D for sweeps; the defaults are the measured best)
#ifndef LM_K1_WARPS
#define LM_K1_WARPS 8
#endif
#ifndef LM_K1_MIN_CTAS
#define LM_K1_MIN_CTAS 3
#endif
#ifndef LM_K1_FB
#define LM_K1_FB 64
#endif
#ifndef LM_K1_COOL_MIN
#define LM_K1_COOL_MIN 4
#endif
#ifndef LM_K1_REDO_ESCAPED_ONLY
#define LM_K1_REDO_ESCAPED_ONLY 1
#endif
// What happens when some lane escaped inside a blind block (LM_K1_REDO_ESCAPED_ONLY = 1):
//   1  the escaped lanes repeat the block alone with the exact test while the other lanes of the warp wait
//   2  the escaped lanes go back to the block's start and the WHOLE warp continues with careful blocks until they
//      have retired: nobody waits, the other lanes pay 8 instead of 6 FP64 instructions for those iterations
#ifndef LM_K1_ESC_MODE
#define LM_K1_ESC_MODE 2
#endif
// cool-down (COOL_MIN careful iterations before the next blind block) after a refill only when a retiring pixel had
// ESCAPED: lanes that retire at max_iter sit inside the set and their next pixels almost always do too
#ifndef LM_K1_ADAPTIVE_COOL
#define LM_K1_ADAPTIVE_COOL 1
#endif

constexpr int TILE = 128;          // pixels per tile (one int4 per lane)
constexpr int WARPS = LM_K1_WARPS; // warps per CTA
constexpr int CTA_THREADS = WARPS * 32;
constexpr int RING = 4;            // resident tiles per warp
constexpr int FB = LM_K1_FB;       // iterations per blind (fast) block
constexpr int COOL_MIN = LM_K1_COOL_MIN;   // careful iterations after a refill / rollback before going blind
constexpr unsigned FULL = 0xffffffffu;

struct EscapeArgs {
    const double* xs;              // grid: x coordinates [nx]; points: c_re [nx]
    const double* ys;              // grid: y coordinates [ny]; points: c_im [nx]
    long long nx, ny;
    unsigned long long chunks_per_row, ntiles;
    int max_iter;
    double thr2;                   // loop threshold on a+b (bailout^2, slightly lowered for hypot tests)
    double bailout;                // R
    double cfar2;                  // lanes with |c|^2 > cfar2 forbid the blind path (<0: never blind)
    int* dwell;                    // [ny*nx] or NULL
    double* dwell_f64;             // [ny*nx] or NULL
    double* field;                 // grid: [ny*nx] or NULL; points: g
    long long* it64;               // points only
    double* phi_re;                // points only
    double* phi_im;                // points only
    unsigned long long* tile_counter;
    unsigned long long* work_counter;   // may be NULL
    int* overflow_flag;            // set when the reference would raise OverflowError
    int vec_i32, vec_f64, vec_field;
    // points only: two-pass schedule (see lm_escape_points_f64_dev)
    const long long* index;        // pass 2: point ids to process (NULL: 0..nx-1)
    const unsigned long long* count_dev;   // pass 2: number of ids, read on the device (NULL: nx)
    long long* survivors;          // pass 1: ids of points still bounded after max_iter (NULL: none kept)
    unsigned long long* survivor_count;
};

// Unfused, round-to-nearest arithmetic in the working precision R: double for the reference's semantics (bit-exact
// dwell), float for the optional single-precision variant (same kernel, same scheduling; lm_escape_grid_f32).
template <typename R> struct Ar;
template <> struct Ar<double> {
    static __device__ __forceinline__ double mul(double x, double y) { return __dmul_rn(x, y); }
    static __device__ __forceinline__ double add(double x, double y) { return __dadd_rn(x, y); }
    static __device__ __forceinline__ double sub(double x, double y) { return __dsub_rn(x, y); }
    static __device__ __forceinline__ double twice_plus(double p, double c) { return __fma_rn(2.0, p, c); }   // == (p + p) + c exactly
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float mul(float x, float y) { return __fmul_rn(x, y); }
    static __device__ __forceinline__ float add(float x, float y) { return __fadd_rn(x, y); }
    static __device__ __forceinline__ float sub(float x, float y) { return __fsub_rn(x, y); }
    static __device__ __forceinline__ float twice_plus(float p, float c) { return __fmaf_rn(2.0f, p, c); }
};

// one unfused iteration z <- z*z + c given the squares a, b of the current z (R = the kernel's working precision)
#define LM_STEP6()                                   \
    do {                                             \
        const R p__ = Ar<R>::mul(zr, zi);            \
        const R t__ = Ar<R>::sub(a, b);              \
        zr = Ar<R>::add(t__, cr);                    \
        zi = Ar<R>::twice_plus(p__, ci);             \
        a = Ar<R>::mul(zr, zr);                      \
        b = Ar<R>::mul(zi, zi);                      \
    } while (0)

// field value at the end of an orbit.  iters = iterations performed (1-based escape index
// when escaped, max_iter otherwise); (zr,zi) = z after `iters` iterations.
template <int FM>
__device__ __forceinline__ double field_value(double zr, double zi, int iters, bool escaped,
                                              int* overflow_flag) {
    if (FM == LM_FIELD_GREEN) {
        // lucas_equipotential_test_v3.py:142-149: Re(log z) * exp2(-k), clamp
        if (!escaped) return 0.0;
        // log|z| as log(zr^2 + zi^2) / 2: at the first escape |z|^2 <= (R^2 + |c|)^2, so nothing overflows and the
        // ~35 instructions of hypot() are saved in this divergent path (only the retiring lanes run it); the value
        // differs from log(hypot()) by < 3e-16 relative (|z|^2 > R^2 keeps the logarithm away from 0), well inside the
        // 5e-15 parity bar of this field.  Escape radii near 1 (logarithm near 0) and huge ones (|z|^2 > 1e300) take hypot.
        const double m2 = fma(zr, zr, zi * zi);
        const double lg = (m2 >= 2.0 && m2 < 1e300) ? 0.5 * log(m2) : log(hypot(zr, zi));
        double g = __dmul_rn(lg, scalbn(1.0, -iters));
        if (!(g >= 0.0) || isinf(g)) g = 0.0;
        return g;
    } else if (FM == LM_FIELD_POW2_ALWAYS) {
        // Potentials.py:43-46: evaluated whether or not the orbit escaped, k 0-based
        const double r = hypot(zr, zi);
        if (!(r > 0.0)) return 0.0;
        const int k = iters - 1;
        if (k > 1023) { atomicExch(overflow_flag, 1); return 0.0; }
        return __dmul_rn(log(r), scalbn(1.0, -k));
    } else if (FM == LM_FIELD_INV_K) {
        // Laplacian_C-M.py:41, Iterative_Variogram_Laplacian.py:126
        if (!escaped) return 0.0;
        return __ddiv_rn(log(hypot(zr, zi)), static_cast<double>(iters));
    } else if (FM == LM_FIELD_POW2_FIRST) {
        // variograms_construct_mandelbrot.py:163
        if (!escaped) return 0.0;
        if (iters > 1023) { atomicExch(overflow_flag, 1); return 0.0; }
        return __dmul_rn(log(hypot(zr, zi)), scalbn(1.0, -iters));
    }
    return 0.0;
}

// POINTS = false : tiles are row segments of the (ny, nx) grid, c = xs[col] + i ys[row]
// POINTS = true  : one "row" of nx points, c = xs[k] + i ys[k]; outputs g/it/phi written directly
// FM             : LM_FIELD_* (LM_FIELD_NONE: dwell only)
// HYPOT          : escape test is hypot(zr,zi) > R (the loop test is a slightly lowered
//                  a+b threshold, confirmed with hypot in the handler)
// R              : working precision (double; float only for the dwell-only grid)
template <bool POINTS, int FM, bool HYPOT, typename R>
__global__ void __launch_bounds__(CTA_THREADS, LM_K1_MIN_CTAS) lm_escape_kernel(const EscapeArgs A) {
    constexpr bool FIELD = (FM != LM_FIELD_NONE);
    const R thr2 = static_cast<R>(A.thr2), cfar2 = static_cast<R>(A.cfar2);
    constexpr int CB = HYPOT ? 1 : 4;          // iterations per careful block

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long s_base[WARPS][RING];
    __shared__ int s_width[WARPS][RING];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // item count: the launch arguments, or (points, pass 2) a device-side counter
    long long NX = A.nx;
    unsigned long long NTILES = A.ntiles, CHUNKS = A.chunks_per_row;
    if (POINTS && A.count_dev) {
        NX = static_cast<long long>(*A.count_dev);
        NTILES = static_cast<unsigned long long>((NX + TILE - 1) / TILE);
        CHUNKS = NTILES;
    }

    int* s_dwell = reinterpret_cast<int*>(smem_raw) + warp * (RING * TILE);
    double* s_field = reinterpret_cast<double*>(smem_raw + WARPS * RING * TILE * sizeof(int)) +
                      warp * (RING * TILE);

    // ---- lane state
    R zr = 0, zi = 0, a = 0, b = 0, cr = 0, ci = 0;
    int n = 0;                       // iterations performed on the current pixel
    bool idle = true;                // lane holds no pixel
    bool far = false;                // |c| too large for the absorbing-escape argument
    bool redo = false;               // rolled back to the start of a blind block: escapes within the next FB iterations
    int my_off = 0;                  // offset of my pixel inside its tile
    unsigned my_seq = 0;             // sequence number (per warp) of my pixel's tile
    long long my_g = 0;              // flat output index of my pixel
    unsigned long long work = 0;     // iterations of finished pixels (this lane)

    // ---- warp state (uniform across lanes)
    unsigned seq = 0;                // tiles acquired so far
    int cursor = 0, width = 0;       // next unassigned pixel of the current tile, its size
    long long base = 0, col0 = 0;    // flat index / first column of the current tile
    R row_ci = 0;
    bool exhausted = false;
    int cool = 0;                    // careful iterations since the last refill / rollback
    unsigned long long pref = 0;     // prefetched tile id (valid in lane 0)
    if (lane == 0) pref = atomicAdd(A.tile_counter, 1ULL);

    auto flush_slot = [&](int slot) {
        __syncwarp();
        const long long fb = s_base[warp][slot];
        const int w = s_width[warp][slot];
        const int* sd = s_dwell + slot * TILE;
        if (A.dwell) {
            if (w == TILE && A.vec_i32) {
                const int4 v = reinterpret_cast<const int4*>(sd)[lane];
                reinterpret_cast<int4*>(A.dwell + fb)[lane] = v;
            } else {
                for (int k = lane; k < w; k += 32) A.dwell[fb + k] = sd[k];
            }
        }
        if (A.dwell_f64) {
            if (w == TILE && A.vec_f64) {
                const int2 lo = reinterpret_cast<const int2*>(sd)[lane];
                const int2 hi = reinterpret_cast<const int2*>(sd)[32 + lane];
                double2* o = reinterpret_cast<double2*>(A.dwell_f64 + fb);
                o[lane] = make_double2(static_cast<double>(lo.x), static_cast<double>(lo.y));
                o[32 + lane] = make_double2(static_cast<double>(hi.x), static_cast<double>(hi.y));
            } else {
                for (int k = lane; k < w; k += 32) A.dwell_f64[fb + k] = static_cast<double>(sd[k]);
            }
        }
        if (FIELD && A.field) {
            const double* sf = s_field + slot * TILE;
            if (w == TILE && A.vec_field) {
                double2* o = reinterpret_cast<double2*>(A.field + fb);
                o[lane] = reinterpret_cast<const double2*>(sf)[lane];
                o[32 + lane] = reinterpret_cast<const double2*>(sf)[32 + lane];
            } else {
                for (int k = lane; k < w; k += 32) A.field[fb + k] = sf[k];
            }
        }
        __syncwarp();
    };

    // hand pixels to the lanes that need one; acquire tiles as required
    auto refill = [&](bool need, bool cool_down) {
        bool assigned_any = false;
        while (true) {
            const unsigned mask = __ballot_sync(FULL, need);
            if (mask == 0u) break;
            if (cursor == width) {
                if (exhausted) break;
                const unsigned long long t = __shfl_sync(FULL, pref, 0);
                if (t >= NTILES) { exhausted = true; break; }
                if (lane == 0) pref = atomicAdd(A.tile_counter, 1ULL);
                unsigned long long row, chunk;
                if ((NTILES >> 32) == 0ULL) {
                    const unsigned r32 = static_cast<unsigned>(t) / static_cast<unsigned>(CHUNKS);
                    row = r32;
                    chunk = static_cast<unsigned>(t) - r32 * static_cast<unsigned>(CHUNKS);
                } else {
                    row = t / CHUNKS;
                    chunk = t - row * CHUNKS;
                }
                col0 = static_cast<long long>(chunk) * TILE;
                const long long left = NX - col0;
                width = left < TILE ? static_cast<int>(left) : TILE;
                base = static_cast<long long>(row) * NX + col0;
                if (!POINTS) {
                    const int slot = seq % RING;
                    if (seq >= RING) flush_slot(slot);
                    if (lane == 0) { s_base[warp][slot] = base; s_width[warp][slot] = width; }
                    row_ci = static_cast<R>(__ldg(A.ys + row));
                }
                cursor = 0;
                ++seq;
            }
            const int rank = __popc(mask & ((1u << lane) - 1u));
            const int avail = width - cursor;
            const int cnt = __popc(mask);
            const int take = cnt < avail ? cnt : avail;
            if (need && rank < take) {
                my_off = cursor + rank;
                my_seq = seq - 1;
                my_g = base + my_off;
                if (POINTS) {
                    if (A.index) my_g = __ldg(A.index + my_g);
                    cr = static_cast<R>(__ldg(A.xs + my_g));
                    ci = static_cast<R>(__ldg(A.ys + my_g));
                } else {
                    cr = static_cast<R>(__ldg(A.xs + col0 + my_off));
                    ci = row_ci;
                }
                zr = 0; zi = 0; a = 0; b = 0; n = 0;
                far = !(cr * cr + ci * ci <= cfar2);
                redo = false;
                idle = false;
                need = false;
            }
            cursor += take;
            assigned_any = true;
        }
        if (need) {          // nothing left for this lane: spin on the origin, emit nothing
            idle = true;
            far = false;
            redo = false;
            cr = 0; ci = 0; zr = 0; zi = 0; a = 0; b = 0; n = 0;
        }
        if (assigned_any && cool_down) cool = 0;
    };

    refill(true, true);

    while (true) {
        if (exhausted && __all_sync(FULL, idle)) break;

        int safe = __reduce_min_sync(FULL, A.max_iter - n);   // >= 1: iterations until the first lane hits max_iter
        bool blind_ok = !HYPOT && !__any_sync(FULL, far || redo);
        bool done = false;             // this lane escaped inside the current run
        int n_fin = 0;                 // iterations performed when it escaped
        R ze_r = 0, ze_i = 0;          // z at the escape (FIELD modes)

        while (true) {
            if (blind_ok && cool >= COOL_MIN && (LM_K1_REDO_ESCAPED_ONLY || safe >= FB)) {
                // ---- blind block: FB iterations, 6 FP64 instructions each, one test at the end.
                // The block runs whatever `safe` says: a lane that passes max_iter inside it simply overshoots
                // (its extra iterates are never looked at).  If its end-of-block test holds, no iterate of the
                // block -- in particular none up to max_iter -- tripped the reference's test (backward induction,
                // see the header), so the pixel retires with dwell = max_iter; if it fails, the lane repeats only
                // the iterations that count.  Round 1 fell back to 4-iteration careful blocks for the WHOLE warp
                // whenever any lane was within FB of max_iter, which is most of the time once refills have
                // staggered the lanes (config 2: 0.56 -> see DESIGN.md).
                const R szr = zr, szi = zi, sa = a, sb = b;
#pragma unroll
                for (int k = 0; k < FB; ++k) LM_STEP6();
                const R m = Ar<R>::add(a, b);
                const bool esc = !(m <= thr2);
#if LM_K1_REDO_ESCAPED_ONLY && LM_K1_ESC_MODE == 2
                if (__any_sync(FULL, esc)) {
                    // Some lane escaped inside this block.  The lanes that did not keep the FB iterations they just
                    // made (their end-of-block test proves no earlier escape); the escaped lanes go back to the saved
                    // state, and the whole warp goes on with careful blocks -- which find their first-escape index
                    // within the next FB iterations (or retire them at max_iter) -- while everybody else keeps
                    // iterating: no lane waits for another one's repeat.
                    if (esc) { zr = szr; zi = szi; a = sa; b = sb; redo = true; }
                    else n += FB;
                    safe -= FB;                                 // still a lower bound of every lane's remaining iterations
                    blind_ok = false;                           // until the rolled-back lanes have retired
                    if (safe <= 0) break;                       // somebody passed max_iter: retire it first
                    continue;
                }
                n += FB;
                safe -= FB;
                if (safe <= 0) break;
                continue;
#elif LM_K1_REDO_ESCAPED_ONLY
                if (__any_sync(FULL, esc)) {
                    // Some lane escaped inside this block.  The lanes that did not keep the FB iterations they
                    // just made (their end-of-block test proves no earlier escape); only the escaped lanes go
                    // back to the saved state and repeat the block with the exact per-iteration test to find
                    // their first-escape index (among the iterations before max_iter).
                    if (esc) {
                        zr = szr; zi = szi; a = sa; b = sb;
                        const int left = A.max_iter - n;
                        const int lim = left < FB ? left : FB;
                        int k = 0;
                        bool found = false;
#pragma unroll 1
                        for (; k < lim; ++k) {
                            LM_STEP6();
                            if (Ar<R>::add(a, b) > thr2) { found = true; break; }
                        }
                        if (found) {
                            done = true; n_fin = n + k + 1;
                            if (FIELD || HYPOT || POINTS) { ze_r = zr; ze_i = zi; }
                        }
                    }
                    n += FB;
                    safe -= FB;
                    break;                                      // to the handler: retire and refill
                }
                n += FB;
                safe -= FB;
                if (safe <= 0) break;
                continue;
#else
                if (__any_sync(FULL, esc)) {
                    zr = szr; zi = szi; a = sa; b = sb;     // some lane escaped in here: redo carefully
                    cool = 0;
                } else {
                    n += FB;
                    safe -= FB;
                    if (safe == 0) break;
                    continue;
                }
#endif
            }
            // ---- careful block: CB iterations with the exact test (first escape kept per lane)
            int cnt;
            if (safe >= CB) {
                unsigned esc_bits = 0u;
#pragma unroll
                for (int k = 0; k < CB; ++k) {
                    LM_STEP6();
                    const R m = Ar<R>::add(a, b);
                    if (m > thr2) {
                        if ((FIELD || HYPOT || POINTS) && esc_bits == 0u) { ze_r = zr; ze_i = zi; }
                        esc_bits |= 1u << k;
                    }
                }
                if (esc_bits) { done = true; n_fin = n + __ffs(esc_bits); }
                cnt = CB;
            } else {
                // fewer than CB iterations left before some lane reaches max_iter: single step
                LM_STEP6();
                const R m = Ar<R>::add(a, b);
                if (m > thr2) {
                    done = true; n_fin = n + 1;
                    if (FIELD || HYPOT || POINTS) { ze_r = zr; ze_i = zi; }
                }
                cnt = 1;
            }
            n += cnt;
            safe -= cnt;
            cool += cnt;
            if (__any_sync(FULL, done) || safe <= 0) break;
        }

        // ---- handler: retire finished pixels, refill their lanes
        if (HYPOT && done) {
            // candidate only: the reference tests abs(z) > R  (CB == 1, so z is still ze)
            if (!(hypot(static_cast<double>(ze_r), static_cast<double>(ze_i)) > A.bailout)) done = false;
        }
        bool need = false;
        if (idle) {
            if (done || n >= A.max_iter) n = 0;
        } else if (POINTS && !done && n >= A.max_iter && A.survivors) {
            // pass 1 of the two-pass point schedule: still bounded after the short budget -> queue for pass 2
            A.survivors[atomicAdd(A.survivor_count, 1ULL)] = my_g;
            need = true;
        } else if (done || n >= A.max_iter) {
            const int iters = done ? n_fin : A.max_iter;
            const int dw = done ? n_fin - 1 : A.max_iter;
            work += static_cast<unsigned long long>(iters);
            double f = 0.0;
            if (FIELD) f = field_value<FM>(static_cast<double>(done ? ze_r : zr), static_cast<double>(done ? ze_i : zi), iters, done, A.overflow_flag);
            if (POINTS) {
                // lucas_equipotential_test_v3.py:140-151
                if (A.field) A.field[my_g] = f;
                if (A.it64) A.it64[my_g] = static_cast<long long>(iters);
                if (A.phi_re || A.phi_im) {
                    double pr = nan(""), pi = nan("");
                    if (done) {
                        // phi = exp(log(z) * 2^-k), principal branch
                        const double s = scalbn(1.0, -iters);
                        const double lr = __dmul_rn(log(hypot(static_cast<double>(ze_r), static_cast<double>(ze_i))), s);
                        const double li = __dmul_rn(atan2(static_cast<double>(ze_i), static_cast<double>(ze_r)), s);
                        const double e = exp(lr);
                        double sn, cs;
                        sincos(li, &sn, &cs);
                        pr = __dmul_rn(e, cs);
                        pi = __dmul_rn(e, sn);
                    }
                    if (A.phi_re) A.phi_re[my_g] = pr;
                    if (A.phi_im) A.phi_im[my_g] = pi;
                }
            } else {
                const unsigned age = seq - 1u - my_seq;
                if (age < static_cast<unsigned>(RING)) {
                    const int slot = my_seq % RING;
                    s_dwell[slot * TILE + my_off] = dw;
                    if (FIELD) s_field[slot * TILE + my_off] = f;
                } else {               // my tile left the ring: patch my own words
                    if (A.dwell) A.dwell[my_g] = dw;
                    if (A.dwell_f64) A.dwell_f64[my_g] = static_cast<double>(dw);
                    if (FIELD && A.field) A.field[my_g] = f;
                }
            }
            need = true;
        }
#if LM_K1_ADAPTIVE_COOL
        refill(need, __any_sync(FULL, done && !idle));
#else
        refill(need, true);
#endif
    }

    if (!POINTS) {
        const unsigned first = seq > static_cast<unsigned>(RING) ? seq - RING : 0u;
        for (unsigned s = first; s < seq; ++s) flush_slot(static_cast<int>(s % RING));
    }
    if (A.work_counter) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) work += __shfl_xor_sync(FULL, work, o);
        if (lane == 0 && work) atomicAdd(A.work_counter, work);
    }
}

size_t smem_bytes(bool points, bool field) {
    if (points) return 16;
    size_t b = static_cast<size_t>(WARPS) * RING * TILE * sizeof(int);
    if (field) b += static_cast<size_t>(WARPS) * RING * TILE * sizeof(double);
    return b;
}

template <bool POINTS, int FM, bool HYPOT, typename R = double>
int32_t launch_one(const EscapeArgs& A, cudaStream_t stream) {
    auto kern = lm_escape_kernel<POINTS, FM, HYPOT, R>;
    const size_t smem = smem_bytes(POINTS, FM != LM_FIELD_NONE);
    LM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    LM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CTA_THREADS, smem));
    if (per_sm < 1) return lm::fail(LM_E_CUDA, "escape kernel does not fit on an SM");
    // no more CTAs than there are tiles for their warps
    unsigned long long want = (A.ntiles + WARPS - 1) / WARPS;
    unsigned long long grid = static_cast<unsigned long long>(lm::sm_count()) * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<static_cast<unsigned>(grid), CTA_THREADS, smem, stream>>>(A);
    LM_CUDA_TRY(cudaGetLastError());
    return LM_OK;
}

int32_t launch_escape(bool points, int field_mode, EscapeArgs& A, cudaStream_t stream, bool single = false) {
    const bool hypot_test = !(field_mode == LM_FIELD_NONE || field_mode == LM_FIELD_GREEN);
    const double R = A.bailout;
    double thr2 = R * R;
    if (hypot_test) thr2 = thr2 * (1.0 - 1e-9);   // candidate threshold; hypot decides
    A.thr2 = thr2;
    {
        // blind path precondition (see header): |c| <= R^2 - R - 0.05
        const double cmax = R * R - R - 0.05;
        A.cfar2 = (!hypot_test && cmax > 0.0 && cmax < 1e100) ? cmax * cmax : -1.0;
    }
    if (single) {
        if (points || field_mode != LM_FIELD_NONE) return lm::fail(LM_E_INVALID, "the single-precision kernel produces the dwell grid only");
        return launch_one<false, LM_FIELD_NONE, false, float>(A, stream);
    }
    if (points) {
        if (field_mode == LM_FIELD_GREEN) return launch_one<true, LM_FIELD_GREEN, false>(A, stream);
        return lm::fail(LM_E_INVALID, "points kernel supports LM_FIELD_GREEN only");
    }
    switch (field_mode) {
        case LM_FIELD_NONE:        return launch_one<false, LM_FIELD_NONE, false>(A, stream);
        case LM_FIELD_GREEN:       return launch_one<false, LM_FIELD_GREEN, false>(A, stream);
        case LM_FIELD_POW2_ALWAYS: return launch_one<false, LM_FIELD_POW2_ALWAYS, true>(A, stream);
        case LM_FIELD_INV_K:       return launch_one<false, LM_FIELD_INV_K, true>(A, stream);
        case LM_FIELD_POW2_FIRST:  return launch_one<false, LM_FIELD_POW2_FIRST, true>(A, stream);
        default: return lm::fail(LM_E_INVALID, "unknown field_mode %d", field_mode);
    }
}

// counters: [0] tile counter, [1] work counter, [2] overflow flag (as 8-byte slots).
// A ring of counter blocks lets several launches be in flight on different streams.  Every block carries an event
// recorded behind the last launch that uses it; a block is handed out again only after that event has completed
// (a caller of the *_dev entry points with more than COUNTER_BLOCKS launches in flight simply waits for the oldest
// one instead of sharing a live tile queue with it).
constexpr int COUNTER_BLOCKS = 64;
int g_counter_next = 0;
int g_counter_dev = -1;
cudaEvent_t g_counter_ev[COUNTER_BLOCKS] = {};
bool g_counter_busy[COUNTER_BLOCKS] = {};
unsigned long long* g_last_counters = nullptr;   // block handed out by the latest get_counters()
void counters_reset() {
    for (int k = 0; k < COUNTER_BLOCKS; ++k) {
        if (g_counter_ev[k]) cudaEventDestroy(g_counter_ev[k]);
        g_counter_ev[k] = nullptr; g_counter_busy[k] = false;
    }
    g_counter_dev = -1;
}
int32_t get_counters(unsigned long long** out, int* slot, cudaStream_t stream) {
    void* p = nullptr;
    int32_t rc = lm::ws_get(lm::WS_COUNTERS, 64 * COUNTER_BLOCKS, &p);
    if (rc != LM_OK) return rc;
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != g_counter_dev) {                   // the workspace moved to another device: events are per device
        counters_reset();
        g_counter_dev = dev;
        lm::register_release_hook(counters_reset);
    }
    const int k = g_counter_next++ % COUNTER_BLOCKS;
    if (!g_counter_ev[k]) LM_CUDA_TRY(cudaEventCreateWithFlags(&g_counter_ev[k], cudaEventDisableTiming));
    if (g_counter_busy[k]) { LM_CUDA_TRY(cudaEventSynchronize(g_counter_ev[k])); g_counter_busy[k] = false; }
    unsigned char* blk = static_cast<unsigned char*>(p) + 64 * k;
    LM_CUDA_TRY(cudaMemsetAsync(blk, 0, 64, stream));
    *out = reinterpret_cast<unsigned long long*>(blk);
    *slot = k;
    g_last_counters = *out;
    return LM_OK;
}
// call behind the last launch that reads / writes the block
int32_t counters_in_flight(int slot, cudaStream_t stream) {
    LM_CUDA_TRY(cudaEventRecord(g_counter_ev[slot], stream));
    g_counter_busy[slot] = true;
    return LM_OK;
}

int32_t check_grid_args(const char* who, const void* xs, int64_t nx, const void* ys, int64_t ny,
                        int32_t max_iter, double bailout) {
    LM_REQUIRE(xs && ys, "%s: xs/ys is NULL", who);
    LM_REQUIRE(nx >= 0 && ny >= 0, "%s: negative grid size", who);
    LM_REQUIRE(max_iter >= 1, "%s: max_iter must be >= 1 (got %d)", who, max_iter);
    LM_REQUIRE(bailout > 0.0 && bailout < 1e150, "%s: bailout out of range", who);
    return LM_OK;
}

// enqueue one K1 launch over rows [0, ny) of a grid whose outputs start at the given pointers
int32_t enqueue_grid(const double* xs, int64_t nx, const double* ys, int64_t ny,
                            int32_t max_iter, double bailout, int32_t field_mode,
                            int32_t* dwell_i32, double* dwell_f64, double* field,
                            unsigned long long* work_dev, int* overflow_dev, cudaStream_t s, bool single = false) {
    unsigned long long* counters = nullptr;
    int slot = 0;
    int32_t rc = get_counters(&counters, &slot, s);
    if (rc != LM_OK) return rc;
    EscapeArgs A{};
    A.xs = xs; A.ys = ys; A.nx = nx; A.ny = ny;
    A.chunks_per_row = static_cast<unsigned long long>((nx + TILE - 1) / TILE);
    A.ntiles = A.chunks_per_row * static_cast<unsigned long long>(ny);
    A.max_iter = max_iter;
    A.bailout = bailout;
    A.dwell = dwell_i32; A.dwell_f64 = dwell_f64;
    A.field = (field_mode == LM_FIELD_NONE) ? nullptr : field;
    A.tile_counter = counters;
    A.work_counter = work_dev;
    A.overflow_flag = overflow_dev ? overflow_dev : reinterpret_cast<int*>(counters + 2);
    A.vec_i32 = (nx % 4 == 0) && (reinterpret_cast<uintptr_t>(dwell_i32) % 16 == 0);
    A.vec_f64 = (nx % 2 == 0) && (reinterpret_cast<uintptr_t>(dwell_f64) % 16 == 0);
    A.vec_field = (nx % 2 == 0) && (reinterpret_cast<uintptr_t>(field) % 16 == 0);
    if ((rc = launch_escape(false, field_mode, A, s, single)) != LM_OK) return rc;
    return counters_in_flight(slot, s);
}

}  // namespace

namespace lm {

void GridHostJob::release_events() {
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    for (cudaEvent_t e : ev_chunk) if (e) cudaEventDestroy(e);
    ev_begin = ev_end = nullptr;
    ev_chunk.clear();
}

#define LM_JOB_TRY(expr)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            job->release_events();                                                             \
            cudaDeviceSynchronize();                                                           \
            return lm::fail(LM_E_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
        }                                                                                      \
    } while (0)

// Host-buffer K1.  Large outputs are produced in row chunks while a copy stream returns finished chunks, so the PCIe
// transfer hides behind the FP64 work.  The chunks alternate between TWO compute streams: a chunk's persistent CTAs
// leave the SMs one by one as the tile queue drains (the last ones hold pixels that run to max_iter), and the next
// chunk's CTAs move in behind them instead of waiting for the whole launch to end -- 16 serial launches cost 15 ms of
// ramps and tails on the 32768^2 grid (626 ms against 611 ms for one launch).
int32_t grid_host_begin(const double* xs, int64_t nx, const double* ys, int64_t ny, int32_t max_iter, double bailout,
                        int32_t field_mode, int32_t* dwell_i32, double* dwell_f64, double* field,
                        bool need_dev_dwell, int64_t extra_rows, GridHostJob* job, const double* row_cost) {
    int32_t rc;
    const size_t npx = static_cast<size_t>(nx) * static_cast<size_t>(ny);
    job->npx = npx;
    static cudaStream_t s_compute = nullptr, s_compute2 = nullptr, s_copy = nullptr;
    static int s_dev = -1;
    lm::register_release_hook([] {
        if (s_compute) {
            cudaStreamDestroy(s_compute); cudaStreamDestroy(s_compute2); cudaStreamDestroy(s_copy);
            s_compute = s_compute2 = s_copy = nullptr; s_dev = -1;
        }
    });
    int dev = 0;
    LM_CUDA_TRY(cudaGetDevice(&dev));
    if (s_dev != dev) {      // (re)create the pipeline streams on the current device
        if (s_compute) { cudaStreamDestroy(s_compute); cudaStreamDestroy(s_compute2); cudaStreamDestroy(s_copy); }
        LM_CUDA_TRY(cudaStreamCreateWithFlags(&s_compute, cudaStreamNonBlocking));
        LM_CUDA_TRY(cudaStreamCreateWithFlags(&s_compute2, cudaStreamNonBlocking));
        LM_CUDA_TRY(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
        s_dev = dev;
    }
    job->s_compute = s_compute;
    job->s_copy = s_copy;
    LM_CUDA_TRY(cudaDeviceSynchronize());   // earlier default-stream work may still use the workspaces

    const bool want_i32 = dwell_i32 != nullptr || need_dev_dwell;
    void *dxs, *dys, *dd = nullptr, *df64 = nullptr, *dfield = nullptr, *dwork;
    if ((rc = ws_get(WS_XS, nx * sizeof(double), &dxs)) != LM_OK) return rc;
    if ((rc = ws_get(WS_YS, ny * sizeof(double), &dys)) != LM_OK) return rc;
    if ((rc = ws_get(WS_K1_WORK, 64, &dwork)) != LM_OK) return rc;
    if (want_i32 && (rc = ws_get(WS_OUT_I32, (npx + static_cast<size_t>(extra_rows) * nx) * sizeof(int32_t), &dd)) != LM_OK) return rc;
    if (dwell_f64 && (rc = ws_get(WS_OUT_F64, npx * sizeof(double), &df64)) != LM_OK) return rc;
    if (field_mode != LM_FIELD_NONE && (rc = ws_get(WS_FIELD, npx * sizeof(double), &dfield)) != LM_OK) return rc;
    job->dwork = dwork;
    job->dwell_dev = static_cast<int32_t*>(dd);
    job->field_dev = static_cast<double*>(dfield);
    LM_CUDA_TRY(cudaMemcpyAsync(dxs, xs, nx * sizeof(double), cudaMemcpyHostToDevice, s_compute));
    LM_CUDA_TRY(cudaMemcpyAsync(dys, ys, ny * sizeof(double), cudaMemcpyHostToDevice, s_compute));
    LM_CUDA_TRY(cudaMemsetAsync(dwork, 0, 64, s_compute));
    unsigned long long* work_dev = static_cast<unsigned long long*>(dwork);
    int* overflow_dev = reinterpret_cast<int*>(work_dev + 1);

    const size_t bytes_per_row = static_cast<size_t>(nx) * ((dwell_i32 ? 4 : 0) + (dwell_f64 ? 8 : 0) + (dfield ? 8 : 0));
    const size_t total_bytes = bytes_per_row * static_cast<size_t>(ny);
    // ~64 MB of output per chunk (1.2 ms of PCIe, more when 8 ranks share the host): the copy of the LAST chunk is what
    // remains after the compute ends; launches are free now that consecutive chunks overlap on two streams
    int64_t nchunks = static_cast<int64_t>((total_bytes + (size_t(64) << 20) - 1) / (size_t(64) << 20));
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 64) nchunks = 64;
    if (nchunks > ny) nchunks = ny;
    const int64_t rows_per_chunk = (ny + nchunks - 1) / nchunks;
    job->ev_chunk.assign(static_cast<size_t>(nchunks), nullptr);

    LM_JOB_TRY(cudaEventCreate(&job->ev_begin));
    LM_JOB_TRY(cudaEventCreate(&job->ev_end));
    LM_JOB_TRY(cudaEventRecord(job->ev_begin, s_compute));                 // behind the coordinate uploads and the counter reset
    LM_JOB_TRY(cudaStreamWaitEvent(s_compute2, job->ev_begin, 0));
    // Order of the chunks.  What remains after the compute ends is the copy of the chunks that finished last, so the
    // cheap chunks (their copy takes longer than their compute) should go first and the expensive ones, whose copies
    // hide behind their own compute, last.  With a per-row cost estimate from the caller (the row profile the shards
    // were cut with) the chunks are sorted by ascending estimated cost; without one they are taken from both ends of
    // the row range towards the middle (0, n-1, 1, n-2, ...), which is that order for a window around the set.
    std::vector<int64_t> order;
    if (row_cost) {
        std::vector<std::pair<double, int64_t>> est;
        for (int64_t c = 0; c < nchunks; ++c) {
            const int64_t r0 = c * rows_per_chunk, r1 = (r0 + rows_per_chunk <= ny) ? r0 + rows_per_chunk : ny;
            double w = 0.0;
            for (int64_t r = r0; r < r1; ++r) w += row_cost[r];
            if (r1 > r0) est.emplace_back(w, c);
        }
        std::stable_sort(est.begin(), est.end(), [](const std::pair<double, int64_t>& a, const std::pair<double, int64_t>& b) { return a.first < b.first; });
        for (const auto& e : est) order.push_back(e.second);
    } else {
        for (int64_t lo_c = 0, hi_c = nchunks - 1; lo_c <= hi_c; ++lo_c, --hi_c) {
            order.push_back(lo_c);
            if (hi_c != lo_c) order.push_back(hi_c);
        }
    }
    int64_t last_on_2 = -1;
    for (size_t t = 0; t < order.size(); ++t) {
        const int64_t c = order[t];
        const int64_t r0 = c * rows_per_chunk;
        const int64_t rows = (r0 + rows_per_chunk <= ny) ? rows_per_chunk : ny - r0;
        if (rows <= 0) continue;
        const size_t off = static_cast<size_t>(r0) * nx;
        cudaStream_t sc = (t & 1) ? s_compute2 : s_compute;
        rc = enqueue_grid(static_cast<double*>(dxs), nx, static_cast<double*>(dys) + r0, rows, max_iter, bailout, field_mode,
                          dd ? static_cast<int32_t*>(dd) + off : nullptr, df64 ? static_cast<double*>(df64) + off : nullptr,
                          dfield ? static_cast<double*>(dfield) + off : nullptr, work_dev, overflow_dev, sc);
        if (rc != LM_OK) { job->release_events(); cudaDeviceSynchronize(); return rc; }
        ++job->launches;
        LM_JOB_TRY(cudaEventCreateWithFlags(&job->ev_chunk[c], cudaEventDisableTiming));
        LM_JOB_TRY(cudaEventRecord(job->ev_chunk[c], sc));
        if (t & 1) last_on_2 = c;
    }
    if (last_on_2 >= 0) LM_JOB_TRY(cudaStreamWaitEvent(s_compute, job->ev_chunk[last_on_2], 0));   // join: K2 and ev_end follow on s_compute
    LM_JOB_TRY(cudaEventRecord(job->ev_end, s_compute));
    for (size_t t = 0; t < order.size(); ++t) {
        const int64_t c = order[t];
        if (!job->ev_chunk[c]) continue;
        const int64_t r0 = c * rows_per_chunk;
        const int64_t rows = (r0 + rows_per_chunk <= ny) ? rows_per_chunk : ny - r0;
        const size_t off = static_cast<size_t>(r0) * nx, cnt = static_cast<size_t>(rows) * nx;
        LM_JOB_TRY(cudaStreamWaitEvent(s_copy, job->ev_chunk[c], 0));
        if (dwell_i32) LM_JOB_TRY(cudaMemcpyAsync(dwell_i32 + off, static_cast<int32_t*>(dd) + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, s_copy));
        if (dwell_f64) LM_JOB_TRY(cudaMemcpyAsync(dwell_f64 + off, static_cast<double*>(df64) + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, s_copy));
        if (dfield) LM_JOB_TRY(cudaMemcpyAsync(field + off, static_cast<double*>(dfield) + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, s_copy));
    }
    return LM_OK;
}

int32_t grid_host_finish(GridHostJob* job, lm_stats* stats) {
    unsigned long long host_counters[2] = {0, 0};
    LM_JOB_TRY(cudaMemcpyAsync(host_counters, job->dwork, sizeof(host_counters), cudaMemcpyDeviceToHost, job->s_compute));
    LM_JOB_TRY(cudaStreamSynchronize(job->s_compute));
    LM_JOB_TRY(cudaStreamSynchronize(job->s_copy));
    float ms = 0.f;
    LM_JOB_TRY(cudaEventElapsedTime(&ms, job->ev_begin, job->ev_end));
    job->release_events();
    if (stats) {
        stats->work_units = host_counters[0];
        stats->items = job->npx;
        stats->kernel_ms = ms;
        stats->launches = job->launches;
    }
    if (static_cast<int>(host_counters[1] & 0xffffffffu))
        return lm::fail(LM_E_OVERFLOW,
                        "lm_escape_grid_f64: 2**k with k > 1023 (the reference raises OverflowError here)");
    return LM_OK;
}
#undef LM_JOB_TRY

}  // namespace lm

extern "C" {

int32_t lm_escape_grid_f64_dev(const double* xs, int64_t nx, const double* ys, int64_t ny,
                               int32_t max_iter, double bailout, int32_t field_mode,
                               int32_t* dwell_i32, double* dwell_f64, double* field,
                               uint64_t* work_units_dev, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    rc = check_grid_args("lm_escape_grid_f64_dev", xs, nx, ys, ny, max_iter, bailout);
    if (rc != LM_OK) return rc;
    LM_REQUIRE(field_mode >= LM_FIELD_NONE && field_mode <= LM_FIELD_POW2_FIRST,
               "lm_escape_grid_f64_dev: unknown field_mode %d", field_mode);
    LM_REQUIRE(field_mode == LM_FIELD_NONE || field != nullptr,
               "lm_escape_grid_f64_dev: field_mode %d needs a field buffer", field_mode);
    cudaStream_t s = lm::as_stream(stream);
    if (work_units_dev) LM_CUDA_TRY(cudaMemsetAsync(work_units_dev, 0, sizeof(uint64_t), s));
    if (nx == 0 || ny == 0) return LM_OK;
    return enqueue_grid(xs, nx, ys, ny, max_iter, bailout, field_mode, dwell_i32, dwell_f64, field,
                        reinterpret_cast<unsigned long long*>(work_units_dev), nullptr, s);
}

int32_t lm_escape_grid_f32_dev(const double* xs, int64_t nx, const double* ys, int64_t ny,
                               int32_t max_iter, double bailout, int32_t* dwell_i32,
                               uint64_t* work_units_dev, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    rc = check_grid_args("lm_escape_grid_f32_dev", xs, nx, ys, ny, max_iter, bailout);
    if (rc != LM_OK) return rc;
    LM_REQUIRE(bailout < 1e15, "lm_escape_grid_f32_dev: bailout out of the single-precision range");
    cudaStream_t s = lm::as_stream(stream);
    if (work_units_dev) LM_CUDA_TRY(cudaMemsetAsync(work_units_dev, 0, sizeof(uint64_t), s));
    if (nx == 0 || ny == 0) return LM_OK;
    return enqueue_grid(xs, nx, ys, ny, max_iter, bailout, LM_FIELD_NONE, dwell_i32, nullptr, nullptr,
                        reinterpret_cast<unsigned long long*>(work_units_dev), nullptr, s, true);
}

int32_t lm_escape_grid_f32(const double* xs, int64_t nx, const double* ys, int64_t ny,
                           int32_t max_iter, double bailout, int32_t* dwell_i32, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    rc = check_grid_args("lm_escape_grid_f32", xs, nx, ys, ny, max_iter, bailout);
    if (rc != LM_OK) return rc;
    LM_REQUIRE(dwell_i32 != nullptr, "lm_escape_grid_f32: dwell_i32 is NULL");
    if (stats) *stats = lm_stats{};
    if (nx == 0 || ny == 0) return LM_OK;
    cudaStream_t s = nullptr;
    const size_t npx = static_cast<size_t>(nx) * ny;
    void *dxs, *dys, *dd, *dw;
    if ((rc = lm::ws_get(lm::WS_XS, nx * sizeof(double), &dxs)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_YS, ny * sizeof(double), &dys)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_I32, npx * sizeof(int), &dd)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_K1_WORK, 64, &dw)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dxs, xs, nx * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dys, ys, ny * sizeof(double), cudaMemcpyHostToDevice, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    rc = lm_escape_grid_f32_dev(static_cast<double*>(dxs), nx, static_cast<double*>(dys), ny, max_iter, bailout,
                                static_cast<int32_t*>(dd), static_cast<uint64_t*>(dw), s);
    if (rc != LM_OK) return rc;
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    uint64_t work = 0;
    LM_CUDA_TRY(cudaMemcpyAsync(dwell_i32, dd, npx * sizeof(int), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaMemcpyAsync(&work, dw, sizeof(work), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) { stats->items = npx; stats->work_units = work; stats->kernel_ms = ms; stats->launches = 1; }
    return LM_OK;
}

int32_t lm_escape_grid_f64(const double* xs, int64_t nx, const double* ys, int64_t ny,
                           int32_t max_iter, double bailout, int32_t field_mode,
                           int32_t* dwell_i32, double* dwell_f64, double* field,
                           lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    rc = check_grid_args("lm_escape_grid_f64", xs, nx, ys, ny, max_iter, bailout);
    if (rc != LM_OK) return rc;
    LM_REQUIRE(field_mode >= LM_FIELD_NONE && field_mode <= LM_FIELD_POW2_FIRST,
               "lm_escape_grid_f64: unknown field_mode %d", field_mode);
    LM_REQUIRE(field_mode == LM_FIELD_NONE || field != nullptr,
               "lm_escape_grid_f64: field_mode %d needs a field buffer", field_mode);
    if (stats) *stats = lm_stats{};
    if (nx == 0 || ny == 0) return LM_OK;
    lm::GridHostJob job;
    rc = lm::grid_host_begin(xs, nx, ys, ny, max_iter, bailout, field_mode, dwell_i32, dwell_f64, field, false, 0, &job, nullptr);
    if (rc != LM_OK) return rc;
    return lm::grid_host_finish(&job, stats);
}

int32_t lm_shard_escape(const double* xs, int64_t nx, const double* ys, int64_t ny, int32_t max_iter,
                        int32_t* dwell_i32, double* potential, int64_t halo_rows, const double* row_cost,
                        int32_t** dwell_dev_out, double** potential_dev_out, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    rc = check_grid_args("lm_shard_escape", xs, nx, ys, ny, max_iter, 2.0);
    if (rc != LM_OK) return rc;
    LM_REQUIRE(halo_rows >= 0 && dwell_dev_out, "lm_shard_escape: bad halo_rows / NULL dwell_dev_out");
    if (stats) *stats = lm_stats{};
    *dwell_dev_out = nullptr;
    if (potential_dev_out) *potential_dev_out = nullptr;
    if (nx == 0 || ny == 0) return LM_OK;
    lm::GridHostJob job;
    rc = lm::grid_host_begin(xs, nx, ys, ny, max_iter, 2.0, potential ? LM_FIELD_GREEN : LM_FIELD_NONE, dwell_i32, nullptr,
                             potential, true, halo_rows, &job, row_cost);
    if (rc != LM_OK) return rc;
    *dwell_dev_out = job.dwell_dev;
    if (potential_dev_out) *potential_dev_out = job.field_dev;
    return lm::grid_host_finish(&job, stats);
}

}  // extern "C"

namespace {

// Point lists mix short-lived and never-escaping orbits at random (a Lucas-Loci cloud: ~90 % of the
// points leave within a few iterations, the rest run to max_iter), which would keep every warp in
// the careful path: each refill resets the blind-mode cool-down.  Two passes restore coherence:
// pass 1 runs every point for PASS1_ITERS iterations and queues the ones still bounded; pass 2 runs
// only those (from z = 0 again, so the recurrence and the counts are unchanged) with the blind path.
constexpr int PASS1_ITERS = 128;
constexpr long long TWO_PASS_MIN_POINTS = 1 << 14;

int32_t enqueue_points(const double* c_re, const double* c_im, int64_t n, int32_t max_iter, double escape_radius,
                       double* g, long long* it, double* phi_re, double* phi_im,
                       unsigned long long* work_dev, int* launches, cudaStream_t s) {
    int32_t rc;
    EscapeArgs A{};
    A.xs = c_re; A.ys = c_im;
    A.nx = n; A.ny = 1;
    A.chunks_per_row = static_cast<unsigned long long>((n + TILE - 1) / TILE);
    A.ntiles = A.chunks_per_row;
    A.bailout = escape_radius;
    A.field = g;
    A.it64 = it;
    A.phi_re = phi_re;
    A.phi_im = phi_im;
    A.work_counter = work_dev;
    unsigned long long* counters = nullptr;
    int slot = 0, slot2 = 0;
    if ((rc = get_counters(&counters, &slot, s)) != LM_OK) return rc;
    A.tile_counter = counters;
    A.overflow_flag = reinterpret_cast<int*>(counters + 2);
    bool two_pass = max_iter > 4 * PASS1_ITERS && n >= TWO_PASS_MIN_POINTS;
    if (const char* e = getenv("LM_K1D_TWO_PASS")) two_pass = (e[0] == '1');     // tuning override
    if (!two_pass) {
        A.max_iter = max_iter;
        if (launches) *launches += 1;
        if ((rc = launch_escape(true, LM_FIELD_GREEN, A, s)) != LM_OK) return rc;
        return counters_in_flight(slot, s);
    }
    void* dsurv = nullptr;
    if ((rc = lm::ws_get(lm::WS_K1_SURVIVORS, static_cast<size_t>(n) * sizeof(long long), &dsurv)) != LM_OK) return rc;
    A.max_iter = PASS1_ITERS;
    A.survivors = static_cast<long long*>(dsurv);
    A.survivor_count = counters + 3;
    if ((rc = launch_escape(true, LM_FIELD_GREEN, A, s)) != LM_OK) return rc;
    unsigned long long* counters2 = nullptr;
    if ((rc = get_counters(&counters2, &slot2, s)) != LM_OK) return rc;
    A.tile_counter = counters2;
    A.overflow_flag = reinterpret_cast<int*>(counters2 + 2);
    A.max_iter = max_iter;
    A.survivors = nullptr; A.survivor_count = nullptr;
    A.index = static_cast<const long long*>(dsurv);
    A.count_dev = counters + 3;
    if (launches) *launches += 2;
    if ((rc = launch_escape(true, LM_FIELD_GREEN, A, s)) != LM_OK) return rc;      // grid sized for n; the kernel reads the real count
    if ((rc = counters_in_flight(slot, s)) != LM_OK) return rc;                      // pass 2 reads its item count from block 1
    return counters_in_flight(slot2, s);
}

}  // namespace

extern "C" {

int32_t lm_escape_points_f64_dev(const double* c_re_dev, const double* c_im_dev, int64_t n,
                                 int32_t max_iter, double escape_radius,
                                 double* g_dev, int64_t* it_dev, double* phi_re_dev, double* phi_im_dev,
                                 uint64_t* work_units_dev, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0, "lm_escape_points_f64_dev: negative n");
    LM_REQUIRE(n == 0 || (c_re_dev && c_im_dev), "lm_escape_points_f64_dev: c_re/c_im is NULL");
    LM_REQUIRE(max_iter >= 1, "lm_escape_points_f64_dev: max_iter must be >= 1");
    LM_REQUIRE(escape_radius > 0.0 && escape_radius < 1e150, "lm_escape_points_f64_dev: bad escape_radius");
    cudaStream_t s = lm::as_stream(stream);
    if (work_units_dev) LM_CUDA_TRY(cudaMemsetAsync(work_units_dev, 0, sizeof(uint64_t), s));
    if (n == 0) return LM_OK;
    return enqueue_points(c_re_dev, c_im_dev, n, max_iter, escape_radius, g_dev, reinterpret_cast<long long*>(it_dev),
                          phi_re_dev, phi_im_dev, reinterpret_cast<unsigned long long*>(work_units_dev), nullptr, s);
}

int32_t lm_escape_points_f64(const double* c_re, const double* c_im, int64_t n,
                             int32_t max_iter, double escape_radius,
                             double* g, int64_t* it, double* phi_re, double* phi_im,
                             lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(n >= 0, "lm_escape_points_f64: negative n");
    LM_REQUIRE(n == 0 || (c_re && c_im), "lm_escape_points_f64: c_re/c_im is NULL");
    LM_REQUIRE(max_iter >= 1, "lm_escape_points_f64: max_iter must be >= 1");
    LM_REQUIRE(escape_radius > 0.0 && escape_radius < 1e150, "lm_escape_points_f64: bad escape_radius");
    if (stats) *stats = lm_stats{};
    if (n == 0) return LM_OK;
    cudaStream_t s = nullptr;
    const size_t nb = static_cast<size_t>(n) * sizeof(double);
    void *dre, *dim, *dg, *dit, *dpr, *dpi, *dwork;
    if ((rc = lm::ws_get(lm::WS_IN_A, nb, &dre)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, nb, &dim)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, nb, &dg)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, nb, &dit)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, nb, &dpr)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_D, nb, &dpi)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dwork)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dre, c_re, nb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(dim, c_im, nb, cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(dwork, 0, 64, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    int launches = 0;
    const bool want_phi = phi_re || phi_im;
    rc = enqueue_points(static_cast<double*>(dre), static_cast<double*>(dim), n, max_iter, escape_radius,
                        static_cast<double*>(dg), static_cast<long long*>(dit),
                        want_phi ? static_cast<double*>(dpr) : nullptr, want_phi ? static_cast<double*>(dpi) : nullptr,
                        static_cast<unsigned long long*>(dwork), &launches, s);
    if (rc != LM_OK) return rc;
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    if (g) LM_CUDA_TRY(cudaMemcpyAsync(g, dg, nb, cudaMemcpyDeviceToHost, s));
    if (it) LM_CUDA_TRY(cudaMemcpyAsync(it, dit, nb, cudaMemcpyDeviceToHost, s));
    if (phi_re) LM_CUDA_TRY(cudaMemcpyAsync(phi_re, dpr, nb, cudaMemcpyDeviceToHost, s));
    if (phi_im) LM_CUDA_TRY(cudaMemcpyAsync(phi_im, dpi, nb, cudaMemcpyDeviceToHost, s));
    uint64_t work = 0;
    LM_CUDA_TRY(cudaMemcpyAsync(&work, dwork, sizeof(work), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->work_units = work;
        stats->items = static_cast<uint64_t>(n);
        stats->kernel_ms = ms;
        stats->launches = launches;
    }
    return LM_OK;
}

}  // extern "C"

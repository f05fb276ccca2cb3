// lm_roots.cu -- K3: batched roots of generalized-Lucas characteristic polynomials (FP64 pipe).
//
// Replaces np.linalg.eigvals(companion(top row)) followed by 1/lambda in
//   compute_inverse_eigenvalues[_family]   lucas_equipotential_test_v3.py:58-118
//   construct_points                       tci_construct_mandelbrot.py:11-19,
//                                          tci_construct_mandelbrot_v002_fixed.py:27-33,
//                                          construct_stage1_clean.py:34-48,
//                                          variograms_construct_mandelbrot.py:48-56
//   compute_inverse_eigenvalues            lucas_to_cardioid_v18_periodic_theta_crbins_artifacts.py:83-94
// The eigenvalues of the companion matrix with first row (a_1..a_d) and unit sub-diagonal are
// the roots of  p(x) = x^d - a_1 x^(d-1) - ... - a_d.  LAPACK's QR iteration is replaced by the
// Aberth-Ehrlich simultaneous iteration, which is embarrassingly parallel over roots:
//   z_i <- z_i - N_i / (1 - N_i * sum_{j != i} 1/(z_i - z_j)),   N_i = p(z_i)/p'(z_i).
//
// Mapping: a group of G lanes owns one polynomial (G = 4 for degree <= 32, 8 up to 128, 32
// above); coefficients and the current root estimates are staged in shared memory; lane l
// updates roots l, l+G, ... (Gauss-Seidel between rounds); group-wide decisions use masked warp
// shuffles/ballots.  For degree <= 32 a warp processes its 8 polynomials as one generation
// (roots_pool_kernel below).  The batch is counting-sorted by degree ON THE DEVICE first (histogram,
// scan, scatter), so the groups that share a warp hold polynomials of the same degree and run
// the same trip counts; the solver launches read their index ranges from device memory, so
// the whole pipeline is asynchronous on one stream (lm_roots_batched_dev).
// The O(d^2) inner loops avoid IEEE division: 1/|z_i - z_j|^2 comes from MUFU.RCP64H plus one
// Newton step (2^-44 relative; the Aberth sum only steers the iteration, its fixed points are
// the zeros of p whatever the sum's accuracy), the Newton ratio p/p' uses two steps.
// Initial guesses follow Bini's Newton-polygon rule (radii from the upper convex hull of
// (k, log|c_k|)), which puts the Lucas family straight onto the unit circle.  For |z| > 1 the
// reversed polynomial is evaluated at 1/z so degrees in the thousands neither overflow nor
// underflow.  A root is frozen when |p(z)| falls below the rounding-error bound of its own
// Horner evaluation.  Parity with LAPACK is tolerance based (sorted roots, 1e-10 relative).
#include "lm_common.cuh"

#include <math.h>

namespace {

constexpr int ROOTS_THREADS = 128;
constexpr int MAX_SWEEPS = 160;
#ifndef LM_K3_SUM_RCP_STEPS
#define LM_K3_SUM_RCP_STEPS 0            // Newton steps on MUFU.RCP64H inside the Aberth sum (0: the raw ~2^-22 estimate)
#endif
constexpr double TWO_PI = 6.283185307179586476925286766559;
constexpr double EPS = 2.220446049250313e-16;

constexpr int CLASS_SMALL_MAX = 32;      // degree <= 32  -> 4 lanes per polynomial
constexpr int CLASS_MID_MAX = 128;       // degree <= 128 -> 8 lanes, above: a full warp
constexpr int HIST_BINS = 4096;          // degrees >= HIST_BINS-1 share the last bin
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;           // polynomials per thread in the sort kernels

// device-side bookkeeping of one batch (lives in the workspace)
struct RootsPlan {
    long long bounds[4];                 // index ranges of the three classes: [bounds[c], bounds[c+1])
    int fail_flag;                       // some polynomial did not converge
    int bad_degree;                      // some deg[k] outside [1, maxdeg]
    unsigned long long cursor;           // next unclaimed polynomial of class 0 (pool kernel, work stealing)
    unsigned long long hist[HIST_BINS];  // per-degree counts, then running cursors
};

struct RootsArgs {
    const double* toprows;     // [npoly * maxdeg]
    const int* deg;            // [npoly]
    const long long* index;    // degree-sorted polynomial ids
    RootsPlan* plan;
    int cls;                   // which class this launch handles
    int maxdeg;
    int invert;
    double tol;
    double* out_re; double* out_im;   // [npoly * maxdeg]
    int* n_kept; int* iters;          // may be NULL
    int smem_deg;              // degree capacity of the shared-memory slices
};

struct cplx { double r, i; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }

// 1/x from MUFU.RCP64H and NR steps (x normal, positive here); STEPS = 1: 2^-44, 2: ~1 ulp
template <int STEPS>
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
    for (int k = 0; k < STEPS; ++k) r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ cplx cinv_fast(cplx a) {
    const double s = rcp_fast<2>(a.r * a.r + a.i * a.i);
    return {a.r * s, -a.i * s};
}
__device__ __forceinline__ cplx cinv_fast1(cplx a) {       // 2^-44: enough for a correction term
    const double s = rcp_fast<1>(a.r * a.r + a.i * a.i);
    return {a.r * s, -a.i * s};
}
__device__ __forceinline__ cplx cinv(cplx a) {
    const double s = 1.0 / (a.r * a.r + a.i * a.i);
    return {a.r * s, -a.i * s};
}

__host__ __device__ inline int degree_bin(int d) { return d < HIST_BINS - 1 ? d : HIST_BINS - 1; }

// ---------------------------------------------------------------------------------------
// counting sort of the polynomial ids by degree
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS) roots_hist_kernel(const int* __restrict__ deg, long long npoly, int maxdeg,
                                                                  RootsPlan* __restrict__ plan) {
    __shared__ unsigned sh[HIST_BINS];
    for (int k = threadIdx.x; k < HIST_BINS; k += SORT_THREADS) sh[k] = 0u;
    __syncthreads();
    const long long base = static_cast<long long>(blockIdx.x) * (SORT_THREADS * SORT_ITEMS);
    bool bad = false;
    for (int t = 0; t < SORT_ITEMS; ++t) {
        const long long k = base + t * SORT_THREADS + threadIdx.x;
        if (k < npoly) {
            const int d = deg[k];
            if (d < 1 || d > maxdeg) bad = true;
            else atomicAdd(&sh[degree_bin(d)], 1u);
        }
    }
    if (bad) plan->bad_degree = 1;
    __syncthreads();
    for (int k = threadIdx.x; k < HIST_BINS; k += SORT_THREADS)
        if (sh[k]) atomicAdd(&plan->hist[k], static_cast<unsigned long long>(sh[k]));
}

// exclusive scan of the histogram (in place: hist becomes the cursors) + class boundaries
__global__ void __launch_bounds__(1024) roots_plan_kernel(RootsPlan* __restrict__ plan) {
    __shared__ unsigned long long part[1024];
    constexpr int PER = HIST_BINS / 1024;
    unsigned long long v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = plan->hist[threadIdx.x * PER + k]; sum += v[k]; }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long y = (threadIdx.x >= o) ? part[threadIdx.x - o] : 0ULL;
        __syncthreads();
        part[threadIdx.x] += y;
        __syncthreads();
    }
    unsigned long long run = part[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int bin = threadIdx.x * PER + k;
        plan->hist[bin] = run;
        if (bin == 0) plan->bounds[0] = 0;
        if (bin == CLASS_SMALL_MAX + 1) plan->bounds[1] = static_cast<long long>(run);
        if (bin == CLASS_MID_MAX + 1) plan->bounds[2] = static_cast<long long>(run);
        run += v[k];
    }
    if (threadIdx.x == 1023) plan->bounds[3] = static_cast<long long>(run);
}

__global__ void __launch_bounds__(SORT_THREADS) roots_scatter_kernel(const int* __restrict__ deg, long long npoly, int maxdeg,
                                                                     RootsPlan* __restrict__ plan, long long* __restrict__ index) {
    __shared__ unsigned cnt[HIST_BINS];
    __shared__ unsigned long long start[HIST_BINS];
    for (int k = threadIdx.x; k < HIST_BINS; k += SORT_THREADS) cnt[k] = 0u;
    __syncthreads();
    const long long base = static_cast<long long>(blockIdx.x) * (SORT_THREADS * SORT_ITEMS);
    int bins[SORT_ITEMS];
    unsigned rank[SORT_ITEMS];
#pragma unroll
    for (int t = 0; t < SORT_ITEMS; ++t) {
        const long long k = base + t * SORT_THREADS + threadIdx.x;
        bins[t] = -1;
        if (k < npoly) {
            const int d = deg[k];
            if (d >= 1 && d <= maxdeg) { bins[t] = degree_bin(d); rank[t] = atomicAdd(&cnt[bins[t]], 1u); }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < HIST_BINS; k += SORT_THREADS)
        if (cnt[k]) start[k] = atomicAdd(&plan->hist[k], static_cast<unsigned long long>(cnt[k]));
    __syncthreads();
#pragma unroll
    for (int t = 0; t < SORT_ITEMS; ++t)
        if (bins[t] >= 0) index[start[bins[t]] + rank[t]] = base + t * SORT_THREADS + threadIdx.x;
}

// ---------------------------------------------------------------------------------------
// the solver
// ---------------------------------------------------------------------------------------
// per-group shared-memory slice: coef[D+1] | z[D] (re, im interleaved) | float logc[D+1] | int hull[D+1] | bytes frozen[D]
// (log|c_k| only places the initial guesses: single precision is plenty and keeps 8 CTAs of 32 polynomials per SM)
__host__ __device__ inline size_t group_smem_bytes(int D) {
    size_t b = sizeof(double) * (static_cast<size_t>(D + 2) + 2 * static_cast<size_t>(D));
    b += sizeof(float) * static_cast<size_t>(D + 1);
    b += sizeof(int) * static_cast<size_t>(D + 1);
    b += static_cast<size_t>(D);
    b = (b + 15) & ~static_cast<size_t>(15);
    if (((b >> 4) & 1) == 0) b += 16;        // stride = 16 * odd: the groups of a warp never share a bank row
    return b;
}

#ifndef LM_K3_MIN_CTAS
#define LM_K3_MIN_CTAS 8
#endif

// group-of-G primitives on raw warp intrinsics with the group's lane mask
struct Tile {
    unsigned mask; int shift, rank;
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) & mask) >> shift; }
    __device__ __forceinline__ bool all(bool p) const { return (__ballot_sync(mask, p) & mask) == mask; }
    __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(mask, v, shift + src); }
    __device__ __forceinline__ int shfl_xor(int v, int o) const { return __shfl_xor_sync(mask, v, o); }
};
template <int G>
__device__ __forceinline__ Tile make_tile() {
    const int lane = threadIdx.x & 31;
    Tile t;
    t.shift = lane & ~(G - 1);
    t.rank = lane & (G - 1);
    t.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << t.shift);
    return t;
}

// one polynomial's shared-memory slice
struct Slot {
    double* coef;            // coef[k] multiplies x^(d-k); coef[0] = 1
    double2* zz;             // current estimates
    float* logc;             // log|coef| (initial guesses only)
    int* hull;               // Newton polygon, later the per-sweep list of roots still moving
    unsigned char* frozen;
};
__device__ __forceinline__ Slot make_slot(unsigned char* base, int D) {
    Slot s;
    s.coef = reinterpret_cast<double*>(base);
    s.zz = reinterpret_cast<double2*>(s.coef + ((D + 2) & ~1));
    s.logc = reinterpret_cast<float*>(s.zz + D);
    s.hull = reinterpret_cast<int*>(s.logc + (D + 1));
    s.frozen = reinterpret_cast<unsigned char*>(s.hull + (D + 1));
    return s;
}

// Load polynomial `pid` into the slot (deflating trailing zero coefficients: roots at 0) and place the initial
// guesses by Bini's rule.  Returns the deflated degree; nzero = number of zero roots.  Called by the G lanes of a group.
template <int G>
__device__ __forceinline__ int load_and_init(const Tile& tile, const Slot& S, const RootsArgs& A, long long pid, int& nzero) {
    const int l = tile.rank;
    const int d_full = A.deg[pid];
    const double* top = A.toprows + pid * A.maxdeg;
    int d;
    {
        int last_nz = 0;
        for (int k = l; k < d_full; k += G)
            if (top[k] != 0.0) last_nz = k + 1;
        for (int o = G / 2; o > 0; o >>= 1) last_nz = max(last_nz, tile.shfl_xor(last_nz, o));
        d = last_nz;
    }
    nzero = d_full - d;
    for (int k = l; k <= d; k += G) {
        const double c = (k == 0) ? 1.0 : -top[k - 1];
        S.coef[k] = c;
        // log|c| only places the starting points (three digits are plenty): the single-precision hardware logarithm
        // whenever |c| is a normal float -- the binary64 log() of this line alone was 9 % of the kernel's
        // instructions (profiles/r02_k3_roots_pool_ncu_by_line.txt)
        const double ac = fabs(c);
        S.logc[k] = (c != 0.0) ? ((ac > 1e-37 && ac < 1e37) ? __logf(static_cast<float>(ac)) : static_cast<float>(log(ac))) : -INFINITY;
    }
    tile.sync();
    // Work with ascending powers: a_i = coef[d-i].  Upper convex hull of (i, log|a_i|), i = 0..d
    // (a_0 = coef[d] != 0 after deflation, a_d = 1).
    int nh = 0;
    if (l == 0 && d > 0) {
        // (single precision throughout: the polygon only places the starting points, and the binary64 version of
        // this one-lane loop spent most of its instructions on float -> double and int -> double conversions)
        for (int i = 0; i <= d; ++i) {
            const float yi = S.logc[d - i];
            if (yi == -INFINITY) continue;
            while (nh >= 2) {
                const int i1 = S.hull[nh - 2], i2 = S.hull[nh - 1];
                const float y1 = S.logc[d - i1], y2 = S.logc[d - i2];
                // keep i2 only if it lies strictly above the chord i1 -> i
                if ((y2 - y1) * static_cast<float>(i - i1) <= (yi - y1) * static_cast<float>(i2 - i1)) --nh; else break;
            }
            S.hull[nh++] = i;
        }
    }
    nh = tile.shfl(nh, 0);
    tile.sync();
    for (int k = l; k < d; k += G) {
        // root k belongs to the hull edge [hull[e], hull[e+1]) that contains k
        int e = 0;
        while (e + 2 < nh && S.hull[e + 1] <= k) ++e;
        const int i1 = S.hull[e], i2 = S.hull[e + 1];
        const int m = i2 - i1;
        // starting points only need a few digits: single-precision hardware exp / sincos (the double versions
        // cost as much as a whole Aberth sweep of a small polynomial)
        const float lr = (S.logc[d - i1] - S.logc[d - i2]) / static_cast<float>(m);
        const double radius = (fabsf(lr) < 80.0f) ? static_cast<double>(__expf(lr)) : exp(static_cast<double>(lr));
        const float ang = static_cast<float>(TWO_PI) * (static_cast<float>(k - i1) / static_cast<float>(m) +
                                                        static_cast<float>(e) / static_cast<float>(d)) + 0.7f;
        float sn, cs;
        __sincosf(ang, &sn, &cs);
        S.zz[k] = make_double2(radius * static_cast<double>(cs), radius * static_cast<double>(sn));
        S.frozen[k] = 0;
    }
    tile.sync();
    return d;
}

// One Aberth-Ehrlich update of root i from the slot's current estimates (read only).  freeze: |p(z)| is below
// the rounding bound of its own evaluation (z is kept); stagnant: the correction no longer changes z.
__device__ __forceinline__ void aberth_update(const Slot& S, int d, int i, cplx& znew, bool& freeze, bool& stagnant) {
    const double2 zi2 = S.zz[i];
    const cplx z = {zi2.x, zi2.y};
    const double az2 = z.r * z.r + z.i * z.i;
    // One Horner loop for both regimes, so lanes inside and outside the unit circle do not diverge:
    // |z| <= 1 evaluates p at w = z from coef[0] up;  |z| > 1 evaluates the reversed polynomial
    // q(w) = sum_k coef[k] w^k at w = 1/z from coef[d] down (p(z) = z^d q(1/z)).
    const bool inside = az2 <= 1.0;
    const double iz2 = rcp_fast<2>(inside ? 1.0 : az2);
    const cplx w = inside ? z : cplx{z.r * iz2, -z.i * iz2};
    // |w| only scales the rounding-error bound of the evaluation: |w|^2 * rsqrt(|w|^2) from the hardware estimate
    // (MUFU.RSQ64H, ~2^-22), rounded UP by 2^-20, is a valid and 1e-6 tight upper bound in three instructions (the
    // binary64 sqrt costs ~20; |w|^2 = 0 gives 0 * inf = NaN, which the max() turns into the harmless bound 0)
    const double aw2 = inside ? az2 : iz2;                       // |w|^2 (|1/z|^2 = 1/|z|^2)
    double rs;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(aw2));
    const double aw = fmax(aw2 * (1.0 + 9.5367431640625e-7) * rs, 0.0);
    const int k_first = inside ? 0 : d, k_step = inside ? 1 : -1;
    cplx b = {S.coef[k_first], 0.0}, bp = {0.0, 0.0};
    double s = fabs(b.r);
#pragma unroll 8
    for (int k = 1, idx = k_first + k_step; k <= d; ++k, idx += k_step) {
        const double ck = S.coef[idx];
        bp = {fma(bp.r, w.r, fma(-bp.i, w.i, b.r)), fma(bp.r, w.i, fma(bp.i, w.r, b.i))};
        b = {fma(b.r, w.r, fma(-b.i, w.i, ck)), fma(b.r, w.i, b.i * w.r)};
        s = fma(s, aw, fabs(ck));
    }
    const double ab2 = b.r * b.r + b.i * b.i;
    const double bound = EPS * s * (d + 1);
    freeze = ab2 <= bound * bound;
    stagnant = false;
    znew = z;
    if (freeze) return;
    // inside:  p/p' = b / bp;   outside:  p/p' = z / (d - w q'(w)/q(w)) = z b / (d b - w bp)
    cplx num, den;
    if (inside) { num = b; den = bp; }
    else {
        const cplx wbp = cmul(w, bp);
        num = cmul(z, b);
        den = {fma(static_cast<double>(d), b.r, -wbp.r), fma(static_cast<double>(d), b.i, -wbp.i)};
    }
    const double dd = den.r * den.r + den.i * den.i;
    const cplx newton = (dd > 1e-290 && dd < 1e290) ? cmul(num, cinv_fast(den))
                      : (dd > 0.0 ? cmul(num, cinv(den)) : cplx{1e-3 * (sqrt(az2) + 1e-3), 1e-3 * (sqrt(az2) + 1e-3)});
    double Sr = 0.0, Si = 0.0;
#pragma unroll 8
    for (int j = 0; j < d; ++j) {
        const double2 zj = S.zz[j];
        const double dr = z.r - zj.x, di = z.i - zj.y;
        // |z_i - z_j|^2 + 1e-300: the j == i term becomes 0 * 1e300 = 0 without a compare / select per term (3 of the
        // 13 instructions of this loop), and the offset is far below the square of any distance that matters
        const double q = fma(dr, dr, fma(di, di, 1e-300));
        const double inv = rcp_fast<LM_K3_SUM_RCP_STEPS>(q);
        Sr = fma(dr, inv, Sr);
        Si = fma(-di, inv, Si);
    }
    const cplx ns = cmul(newton, cplx{Sr, Si});
    const cplx den2 = {1.0 - ns.r, -ns.i};
    const double d2 = den2.r * den2.r + den2.i * den2.i;
    const cplx corr = (d2 > 1e-290 && d2 < 1e290) ? cmul(newton, cinv_fast1(den2))
                    : (d2 > 0.0 ? cmul(newton, cinv(den2)) : newton);
    znew = {z.r - corr.r, z.i - corr.i};
    const double c2 = corr.r * corr.r + corr.i * corr.i;
    if (!(c2 < 1e300)) znew = {z.r * 0.5 + 1e-3, z.i * 0.5 - 1e-3};      // a non-finite (or absurd) correction: restart nearby
    // stagnation: the correction is below the resolution of z -> it is frozen with the new value
    if (c2 <= (4.0 * EPS * EPS) * az2) stagnant = true;
}

// One sweep over the roots of a slot that are still moving (rounds of G roots, Gauss-Seidel between rounds).
// Returns true when the sweep updated nothing (every root is frozen).
template <int G>
__device__ __forceinline__ bool sweep_slot(const Tile& tile, const Slot& S, int d) {
    const int l = tile.rank;
    // compact the roots that are still moving into the (now free) hull array, so late sweeps with a few
    // stragglers take one round instead of ceil(d/G)
    int nact = 0;
    for (int r0 = 0; r0 < d; r0 += G) {
        const int i = r0 + l;
        const bool live = (i < d) && !S.frozen[i];
        const unsigned bal = tile.ballot(live);
        if (live) S.hull[nact + __popc(bal & ((1u << l) - 1u))] = i;
        nact += __popc(bal);
    }
    tile.sync();
    bool mine_done = true;
    for (int r0 = 0; r0 < nact; r0 += G) {
        const bool active = r0 + l < nact;
        const int i = active ? S.hull[r0 + l] : 0;
        bool freeze = false, stagnant = false;
        cplx znew = {0.0, 0.0};
        if (active) {
            aberth_update(S, d, i, znew, freeze, stagnant);
            if (!freeze) mine_done = false;
        }
        tile.sync();               // everybody has read the old estimates of this round
        if (active) {
            if (freeze) S.frozen[i] = 1;
            else { S.zz[i] = make_double2(znew.r, znew.i); if (stagnant) S.frozen[i] = 1; }
        }
        tile.sync();
    }
    return tile.all(mine_done);
}

// [zero roots] + computed roots, optionally inverted / filtered / compacted, NaN padding, counters
template <int G>
__device__ __forceinline__ void write_output(const Tile& tile, const Slot& S, const RootsArgs& A, long long pid, int d, int nzero,
                                             int sweeps, bool converged) {
    const int l = tile.rank;
    double* ore = A.out_re + pid * A.maxdeg;
    double* oim = A.out_im + pid * A.maxdeg;
    int kept = 0;
    const int total = d + nzero;
    for (int r0 = 0; r0 < total; r0 += G) {
        const int i = r0 + l;
        bool keep = false;
        cplx v = {0.0, 0.0};
        if (i < total) {
            cplx z = {0.0, 0.0};
            if (i < d) { const double2 t = S.zz[i]; z = {t.x, t.y}; }
            if (A.invert) {
                const double az = sqrt(z.r * z.r + z.i * z.i);
                keep = az > A.tol;
                if (keep) v = cinv(z);
            } else {
                keep = true;
                v = z;
            }
        }
        const unsigned bal = tile.ballot(keep);
        if (keep) {
            const int pos = kept + __popc(bal & ((1u << l) - 1u));
            ore[pos] = v.r;
            oim[pos] = v.i;
        }
        kept += __popc(bal);
    }
    for (int k = kept + l; k < A.maxdeg; k += G) { ore[k] = nan(""); oim[k] = nan(""); }
    if (l == 0) {
        if (A.n_kept) A.n_kept[pid] = kept;
        if (A.iters) A.iters[pid] = converged ? sweeps : -sweeps;
        if (!converged) atomicExch(&A.plan->fail_flag, 1);
    }
    tile.sync();
}

// ---- group kernel: G lanes own one polynomial from start to end (degrees above CLASS_SMALL_MAX) ----
template <int G>
__global__ void __launch_bounds__(ROOTS_THREADS, LM_K3_MIN_CTAS) roots_kernel(const RootsArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tile tile = make_tile<G>();
    const int groups_per_cta = blockDim.x / G;
    const int gid = threadIdx.x / G;
    const int l = tile.rank;
    const int D = A.smem_deg;
    const Slot S = make_slot(smem + static_cast<size_t>(gid) * group_smem_bytes(D), D);
    const long long first = A.plan->bounds[A.cls], last = A.plan->bounds[A.cls + 1];

    for (long long item = first + static_cast<long long>(blockIdx.x) * groups_per_cta + gid; item < last;
         item += static_cast<long long>(gridDim.x) * groups_per_cta) {
        const long long pid = A.index[item];
        int nzero = 0;
        const int d = load_and_init<G>(tile, S, A, pid, nzero);
        int sweeps = 0;
        bool all_done = (d == 0);
        while (!all_done && sweeps < MAX_SWEEPS) {
            ++sweeps;
            all_done = sweep_slot<G>(tile, S, d);
        }
        write_output<G>(tile, S, A, pid, d, nzero, sweeps, all_done);
    }
}

// ---- generation kernel (degree <= CLASS_SMALL_MAX): a warp takes 8 polynomials at a time (one generation, claimed
// with an atomic cursor over the degree-sorted batch), one per 4-lane group.  The groups load and initialise their
// polynomials at the same time, walk through the sweeps side by side, and write their results together.  In the
// first version every group ran on its own schedule and ncu showed the sequential Newton-polygon scan of a new
// polynomial running on ONE lane while the other 31 waited (13 % of all instructions at 1.4 active threads); in a
// generation 8 scans run side by side.  Every slot keeps its own lanes and its own update schedule (rounds of 4
// roots, Gauss-Seidel between rounds), so the result of a polynomial does not depend on which polynomials share
// its warp: runs are bit-reproducible.  (A variant that pooled the 32 lanes over all (polynomial, root) pairs still
// moving was measured too: same speed, but the update schedule -- and with it the last bits of ill-conditioned
// roots -- then depends on the neighbours, so it was dropped.)
constexpr int POOL_SLOTS = 8;
constexpr int POOL_WARPS = ROOTS_THREADS / 32;

__global__ void __launch_bounds__(ROOTS_THREADS, LM_K3_MIN_CTAS) roots_pool_kernel(const RootsArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tile tile = make_tile<4>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, l = tile.rank;                   // my slot (for load / output), my rank in its group
    const int D = A.smem_deg;
    const size_t slot_bytes = group_smem_bytes(D);
    unsigned char* warp_base = smem + static_cast<size_t>(warp) * POOL_SLOTS * slot_bytes;
    const Slot S = make_slot(warp_base + static_cast<size_t>(g) * slot_bytes, D);
    const long long first = A.plan->bounds[A.cls], last = A.plan->bounds[A.cls + 1];

    while (true) {
        // ---- claim a generation of 8 polynomials
        unsigned long long gen = 0;
        if (lane == 0) gen = atomicAdd(&A.plan->cursor, static_cast<unsigned long long>(POOL_SLOTS));
        gen = __shfl_sync(0xffffffffu, gen, 0);
        const long long item = first + static_cast<long long>(gen) + g;
        if (first + static_cast<long long>(gen) >= last) break;
        const bool occupied = item < last;
        long long pid = -1;
        int d = 0, nzero = 0, sweeps = 0;
        if (occupied) {
            pid = A.index[item];
            d = load_and_init<4>(tile, S, A, pid, nzero);
        }
        __syncwarp();
        bool moving = occupied && d > 0;          // some root of my slot was still updated in its last sweep
        bool gave_up = false;

        // ---- sweeps over the whole generation; a slot that has converged waits for the generation to finish
        while (__ballot_sync(0xffffffffu, moving) != 0u) {
            if (moving) {
                if (sweeps >= MAX_SWEEPS) { moving = false; gave_up = true; }
                else { ++sweeps; if (sweep_slot<4>(tile, S, d)) moving = false; }
            }
        }

        // ---- the generation's results
        if (occupied) write_output<4>(tile, S, A, pid, d, nzero, sweeps, !gave_up);
        __syncwarp();
    }
}

template <int G>
int32_t launch_roots(RootsArgs A, long long npoly, cudaStream_t s) {
    auto kern = roots_kernel<G>;
    int groups = ROOTS_THREADS / G;
    const size_t per_group = group_smem_bytes(A.smem_deg);
    const size_t budget = 200 * 1024;
    if (per_group * groups > budget) groups = static_cast<int>(budget / per_group);
    if (groups < 1)
        return lm::fail(LM_E_INVALID, "lm_roots_batched: degree %d needs %zu bytes of shared memory per polynomial",
                        A.smem_deg, per_group);
    const size_t smem = per_group * groups;
    LM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    long long blocks = (npoly + groups - 1) / groups;          // the class size is only known on the device
    const long long cap = static_cast<long long>(lm::sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    kern<<<static_cast<unsigned>(blocks), groups * G, smem, s>>>(A);
    LM_CUDA_TRY(cudaGetLastError());
    return LM_OK;
}

// the whole K3 pipeline on one stream, device buffers in and out; plan_out = device bookkeeping
int32_t roots_enqueue(const double* toprows, const int* deg, long long npoly, int maxdeg, int invert, double tol,
                      double* out_re, double* out_im, int* n_kept, int* iters, RootsPlan** plan_out, int* launches,
                      cudaStream_t s) {
    int32_t rc;
    void *dplan, *dindex;
    if ((rc = lm::ws_get(lm::WS_ROOTS_PLAN, sizeof(RootsPlan), &dplan)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_ROOTS_INDEX, static_cast<size_t>(npoly) * sizeof(long long), &dindex)) != LM_OK) return rc;
    RootsPlan* plan = static_cast<RootsPlan*>(dplan);
    LM_CUDA_TRY(cudaMemsetAsync(plan, 0, sizeof(RootsPlan), s));
    const unsigned sort_blocks = static_cast<unsigned>((npoly + SORT_THREADS * SORT_ITEMS - 1) / (SORT_THREADS * SORT_ITEMS));
    roots_hist_kernel<<<sort_blocks, SORT_THREADS, 0, s>>>(deg, npoly, maxdeg, plan);
    roots_plan_kernel<<<1, 1024, 0, s>>>(plan);
    roots_scatter_kernel<<<sort_blocks, SORT_THREADS, 0, s>>>(deg, npoly, maxdeg, plan, static_cast<long long*>(dindex));
    LM_CUDA_TRY(cudaGetLastError());
    RootsArgs A{};
    A.toprows = toprows; A.deg = deg; A.index = static_cast<const long long*>(dindex); A.plan = plan;
    A.maxdeg = maxdeg; A.invert = invert; A.tol = tol;
    A.out_re = out_re; A.out_im = out_im; A.n_kept = n_kept; A.iters = iters;
    int n = 3;
    A.cls = 0; A.smem_deg = maxdeg < CLASS_SMALL_MAX ? maxdeg : CLASS_SMALL_MAX;
    {
        const size_t smem = group_smem_bytes(A.smem_deg) * POOL_SLOTS * POOL_WARPS;
        // function attribute and occupancy are per device (and per shared-memory size)
        static int per_sm_dev[64] = {};
        static size_t per_sm_smem[64] = {};
        int dev = 0;
        LM_CUDA_TRY(cudaGetDevice(&dev));
        const int di = (dev >= 0 && dev < 64) ? dev : 63;
        if (per_sm_dev[di] == 0 || per_sm_smem[di] != smem || di == 63) {
            int v = 0;
            LM_CUDA_TRY(cudaFuncSetAttribute(roots_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            LM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, roots_pool_kernel, ROOTS_THREADS, smem));
            per_sm_dev[di] = v < 1 ? 1 : v;
            per_sm_smem[di] = smem;
        }
        const int per_sm = per_sm_dev[di];
        long long blocks = (npoly + POOL_SLOTS * POOL_WARPS - 1) / (POOL_SLOTS * POOL_WARPS);
        const long long cap = static_cast<long long>(lm::sm_count()) * per_sm;      // persistent CTAs, polynomials by work stealing
        if (blocks > cap) blocks = cap;
        roots_pool_kernel<<<static_cast<unsigned>(blocks), ROOTS_THREADS, smem, s>>>(A);
        LM_CUDA_TRY(cudaGetLastError());
    }
    ++n;
    if (maxdeg > CLASS_SMALL_MAX) {
        A.cls = 1; A.smem_deg = maxdeg < CLASS_MID_MAX ? maxdeg : CLASS_MID_MAX;
        if ((rc = launch_roots<8>(A, npoly, s)) != LM_OK) return rc;
        ++n;
    }
    if (maxdeg > CLASS_MID_MAX) {
        A.cls = 2; A.smem_deg = maxdeg;
        if ((rc = launch_roots<32>(A, npoly, s)) != LM_OK) return rc;
        ++n;
    }
    if (plan_out) *plan_out = plan;
    if (launches) *launches = n;
    return LM_OK;
}

// ---------------------------------------------------------------------------------------
// cloud compaction: the valid slots of every polynomial, concatenated in polynomial order
// ---------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;             // elements per thread -> 2048 per block

__global__ void __launch_bounds__(SCAN_THREADS) kept_block_sums_kernel(const int* __restrict__ n_kept, long long npoly,
                                                                       unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long wsum[SCAN_THREADS / 32];
    const long long base = static_cast<long long>(blockIdx.x) * (SCAN_THREADS * SCAN_ITEMS) + threadIdx.x * SCAN_ITEMS;
    unsigned long long s = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) if (base + t < npoly) s += static_cast<unsigned long long>(n_kept[base + t]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) t += wsum[k];
        block_sums[blockIdx.x] = t;
    }
}

// exclusive scan of the block sums in place (single CTA), total -> *total_out.  append != 0: the scan starts
// at the value *total_out already holds (the cloud of the previous chunks) and adds to it.
__global__ void __launch_bounds__(1024) kept_scan_sums_kernel(unsigned long long* __restrict__ block_sums, long long nblocks,
                                                              long long* __restrict__ total_out, int append) {
    __shared__ unsigned long long part[1024];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = append ? static_cast<unsigned long long>(*total_out) : 0ULL;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long idx = base + threadIdx.x;
        const unsigned long long v = idx < nblocks ? block_sums[idx] : 0ULL;
        part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const unsigned long long y = (threadIdx.x >= o) ? part[threadIdx.x - o] : 0ULL;
            __syncthreads();
            part[threadIdx.x] += y;
            __syncthreads();
        }
        const unsigned long long carry = carry_s;
        if (idx < nblocks) block_sums[idx] = carry + part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = static_cast<long long>(carry_s);
}

__global__ void __launch_bounds__(SCAN_THREADS) cloud_gather_kernel(const double* __restrict__ re, const double* __restrict__ im,
                                                                    const int* __restrict__ n_kept, long long npoly, int maxdeg,
                                                                    const unsigned long long* __restrict__ block_offsets,
                                                                    double* __restrict__ px, double* __restrict__ py,
                                                                    long long cap) {
    // phase 1: exclusive offsets of the block's SCAN_THREADS * SCAN_ITEMS polynomials (thread t owns SCAN_ITEMS
    // consecutive ones) into shared memory; phase 2: a warp per polynomial, lane r copies kept root r, so both
    // the padded rows and the packed cloud are touched in contiguous runs
    __shared__ unsigned long long wsum[SCAN_THREADS / 32];
    __shared__ unsigned long long s_pos[SCAN_THREADS * SCAN_ITEMS];
    __shared__ int s_cnt[SCAN_THREADS * SCAN_ITEMS];
    const long long block_base = static_cast<long long>(blockIdx.x) * (SCAN_THREADS * SCAN_ITEMS);
    const long long base = block_base + threadIdx.x * SCAN_ITEMS;
    int cnt[SCAN_ITEMS];
    unsigned long long mine = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) { cnt[t] = (base + t < npoly) ? n_kept[base + t] : 0; mine += cnt[t]; }
    unsigned long long incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long before = block_offsets[blockIdx.x];
    for (int k = 0; k < warp; ++k) before += wsum[k];
    unsigned long long pos = before + incl - mine;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) {
        s_pos[threadIdx.x * SCAN_ITEMS + t] = pos;
        s_cnt[threadIdx.x * SCAN_ITEMS + t] = cnt[t];
        pos += cnt[t];
    }
    __syncthreads();
    for (int q = warp; q < SCAN_THREADS * SCAN_ITEMS; q += SCAN_THREADS / 32) {
        const long long pid = block_base + q;
        if (pid >= npoly) break;
        const int c = s_cnt[q];
        const unsigned long long p0 = s_pos[q];
        for (int r = lane; r < c; r += 32) {
            if (static_cast<long long>(p0 + r) < cap) { px[p0 + r] = re[pid * maxdeg + r]; py[p0 + r] = im[pid * maxdeg + r]; }
        }
    }
}

}  // namespace

extern "C" {

int32_t lm_roots_batched_dev(const double* toprows_dev, const int32_t* deg_dev, int64_t npoly, int32_t maxdeg,
                             int32_t invert, double tol, double* out_re_dev, double* out_im_dev,
                             int32_t* n_kept_dev, int32_t* iters_dev, int32_t* status_dev, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npoly >= 0 && maxdeg >= 1, "lm_roots_batched_dev: bad sizes");
    LM_REQUIRE(npoly == 0 || (toprows_dev && deg_dev && out_re_dev && out_im_dev), "lm_roots_batched_dev: NULL buffer");
    cudaStream_t s = lm::as_stream(stream);
    if (npoly == 0) {
        if (status_dev) LM_CUDA_TRY(cudaMemsetAsync(status_dev, 0, 2 * sizeof(int32_t), s));
        return LM_OK;
    }
    RootsPlan* plan = nullptr;
    rc = roots_enqueue(toprows_dev, deg_dev, npoly, maxdeg, invert, tol, out_re_dev, out_im_dev, n_kept_dev, iters_dev,
                       &plan, nullptr, s);
    if (rc != LM_OK) return rc;
    if (status_dev)
        LM_CUDA_TRY(cudaMemcpyAsync(status_dev, &plan->fail_flag, 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    return LM_OK;
}

namespace {
int32_t cloud_compact_impl(const double* re_dev, const double* im_dev, const int32_t* n_kept_dev, int64_t npoly,
                           int32_t maxdeg, double* px_dev, double* py_dev, int64_t cap_points,
                           int64_t* n_points_dev, int append, void* stream);
}

int32_t lm_cloud_compact_dev(const double* re_dev, const double* im_dev, const int32_t* n_kept_dev, int64_t npoly,
                             int32_t maxdeg, double* px_dev, double* py_dev, int64_t cap_points,
                             int64_t* n_points_dev, void* stream) {
    return cloud_compact_impl(re_dev, im_dev, n_kept_dev, npoly, maxdeg, px_dev, py_dev, cap_points, n_points_dev, 0, stream);
}

int32_t lm_cloud_append_dev(const double* re_dev, const double* im_dev, const int32_t* n_kept_dev, int64_t npoly,
                            int32_t maxdeg, double* px_dev, double* py_dev, int64_t cap_points,
                            int64_t* n_points_inout_dev, void* stream) {
    return cloud_compact_impl(re_dev, im_dev, n_kept_dev, npoly, maxdeg, px_dev, py_dev, cap_points, n_points_inout_dev, 1, stream);
}

}  // extern "C"

namespace {
int32_t cloud_compact_impl(const double* re_dev, const double* im_dev, const int32_t* n_kept_dev, int64_t npoly,
                           int32_t maxdeg, double* px_dev, double* py_dev, int64_t cap_points,
                           int64_t* n_points_dev, int append, void* stream) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npoly >= 0 && maxdeg >= 1 && cap_points >= 0, "lm_cloud_compact_dev: bad sizes");
    LM_REQUIRE(n_points_dev != nullptr, "lm_cloud_compact_dev: n_points_dev is NULL");
    LM_REQUIRE(npoly == 0 || (re_dev && im_dev && n_kept_dev && (cap_points == 0 || (px_dev && py_dev))),
               "lm_cloud_compact_dev: NULL buffer");
    cudaStream_t s = lm::as_stream(stream);
    if (npoly == 0) {
        if (!append) LM_CUDA_TRY(cudaMemsetAsync(n_points_dev, 0, sizeof(int64_t), s));
        return LM_OK;
    }
    const long long per_block = SCAN_THREADS * SCAN_ITEMS;
    const long long nblocks = (npoly + per_block - 1) / per_block;
    void* dsums;
    if ((rc = lm::ws_get(lm::WS_CLOUD_SCAN, static_cast<size_t>(nblocks) * sizeof(unsigned long long), &dsums)) != LM_OK) return rc;
    unsigned long long* sums = static_cast<unsigned long long*>(dsums);
    kept_block_sums_kernel<<<static_cast<unsigned>(nblocks), SCAN_THREADS, 0, s>>>(n_kept_dev, npoly, sums);
    kept_scan_sums_kernel<<<1, 1024, 0, s>>>(sums, nblocks, reinterpret_cast<long long*>(n_points_dev), append);
    cloud_gather_kernel<<<static_cast<unsigned>(nblocks), SCAN_THREADS, 0, s>>>(re_dev, im_dev, n_kept_dev, npoly, maxdeg, sums,
                                                                                px_dev, py_dev, cap_points);
    LM_CUDA_TRY(cudaGetLastError());
    return LM_OK;
}
}  // namespace

extern "C" {

int32_t lm_roots_batched(const double* toprows, const int32_t* deg, int64_t npoly, int32_t maxdeg,
                         int32_t invert, double tol, double* out_re, double* out_im,
                         int32_t* n_kept, int32_t* iters, lm_stats* stats) {
    int32_t rc = lm::require_device();
    if (rc != LM_OK) return rc;
    LM_REQUIRE(npoly >= 0 && maxdeg >= 1, "lm_roots_batched: bad sizes");
    LM_REQUIRE(npoly == 0 || (toprows && deg && out_re && out_im), "lm_roots_batched: NULL buffer");
    if (stats) *stats = lm_stats{};
    if (npoly == 0) return LM_OK;
    cudaStream_t s = nullptr;
    const size_t ncoef = static_cast<size_t>(npoly) * maxdeg;
    void *dtop, *ddeg, *dre, *dim, *dkept, *diters;
    if ((rc = lm::ws_get(lm::WS_IN_A, ncoef * sizeof(double), &dtop)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, static_cast<size_t>(npoly) * sizeof(int), &ddeg)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, ncoef * sizeof(double), &dre)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, ncoef * sizeof(double), &dim)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, static_cast<size_t>(npoly) * sizeof(int), &dkept)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_D, static_cast<size_t>(npoly) * sizeof(int), &diters)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dtop, toprows, ncoef * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(ddeg, deg, static_cast<size_t>(npoly) * sizeof(int), cudaMemcpyHostToDevice, s));
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    RootsPlan* plan = nullptr;
    int launches = 0;
    rc = roots_enqueue(static_cast<double*>(dtop), static_cast<int*>(ddeg), npoly, maxdeg, invert, tol,
                       static_cast<double*>(dre), static_cast<double*>(dim), static_cast<int*>(dkept),
                       static_cast<int*>(diters), &plan, &launches, s);
    if (rc != LM_OK) return rc;
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    int flags[2] = {0, 0};
    LM_CUDA_TRY(cudaMemcpyAsync(flags, &plan->fail_flag, sizeof(flags), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (flags[1]) {
        for (int64_t k = 0; k < npoly; ++k)
            if (deg[k] < 1 || deg[k] > maxdeg)
                return lm::fail(LM_E_INVALID, "lm_roots_batched: deg[%lld] = %d outside [1, %d]", static_cast<long long>(k),
                                deg[k], maxdeg);
        return lm::fail(LM_E_INVALID, "lm_roots_batched: a degree outside [1, %d]", maxdeg);
    }
    LM_CUDA_TRY(cudaMemcpyAsync(out_re, dre, ncoef * sizeof(double), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaMemcpyAsync(out_im, dim, ncoef * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (n_kept) LM_CUDA_TRY(cudaMemcpyAsync(n_kept, dkept, static_cast<size_t>(npoly) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (iters) LM_CUDA_TRY(cudaMemcpyAsync(iters, diters, static_cast<size_t>(npoly) * sizeof(int), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        uint64_t nroots = 0;
        for (int64_t k = 0; k < npoly; ++k) nroots += static_cast<uint64_t>(deg[k]);
        stats->items = static_cast<uint64_t>(npoly);
        stats->work_units = nroots;
        stats->kernel_ms = ms;
        stats->launches = launches;
    }
    if (flags[0])
        return lm::fail(LM_E_NOCONV, "lm_roots_batched: Aberth iteration did not converge within %d sweeps for some polynomial "
                        "(iters < 0 marks them; outputs hold the last estimates)", MAX_SWEEPS);
    return LM_OK;
}

}  // extern "C"

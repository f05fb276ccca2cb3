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
// Mapping: a group of G lanes (G = 8, 16 or 32, chosen per polynomial from its degree by
// host-side binning) owns one polynomial; coefficients and the current root estimates are
// staged in shared memory; lane l updates roots l, l+G, ... (Gauss-Seidel between rounds);
// group-wide decisions use tile shuffles/ballots.  Initial guesses follow Bini's Newton-
// polygon rule (radii from the upper convex hull of (k, log|c_k|)), which puts the Lucas
// family straight onto the unit circle.  For |z| > 1 the reversed polynomial is evaluated
// at 1/z so degrees in the thousands neither overflow nor underflow.  A root is frozen when
// |p(z)| falls below the rounding-error bound of its own Horner evaluation.
// Parity with LAPACK is tolerance based (sorted roots, 1e-10 relative; see tests).
#include "lm_common.cuh"

#include <cooperative_groups.h>
#include <math.h>
#include <vector>

namespace cg = cooperative_groups;

namespace {

constexpr int ROOTS_THREADS = 128;
constexpr int MAX_SWEEPS = 160;
constexpr double TWO_PI = 6.283185307179586476925286766559;
constexpr double EPS = 2.220446049250313e-16;

struct RootsArgs {
    const double* toprows;     // [npoly * maxdeg]
    const int* deg;            // [npoly]
    const long long* index;    // polynomial ids handled by this launch [count]
    long long count;
    int maxdeg;
    int invert;
    double tol;
    double* out_re; double* out_im;   // [npoly * maxdeg]
    int* n_kept; int* iters;          // may be NULL
    int* fail_flag;
    int smem_deg;              // degree capacity of the shared-memory slices
};

struct cplx { double r, i; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ cplx cinv(cplx a) {
    const double s = 1.0 / (a.r * a.r + a.i * a.i);
    return {a.r * s, -a.i * s};
}

// per-group shared-memory slice layout (doubles): coef[D+1] | zr[D] | zi[D] | logc[D+1], then ints hull[D+1], then bytes frozen[D]
__host__ __device__ inline size_t group_smem_bytes(int D) {
    size_t b = sizeof(double) * (static_cast<size_t>(D + 1) + D + D + (D + 1));
    b += sizeof(int) * static_cast<size_t>(D + 1);
    b += static_cast<size_t>(D);
    return (b + 15) & ~static_cast<size_t>(15);
}

template <int G>
__global__ void __launch_bounds__(ROOTS_THREADS) roots_kernel(const RootsArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<G> tile = cg::tiled_partition<G>(block);
    const int groups_per_cta = blockDim.x / G;
    const int gid = threadIdx.x / G;
    const int l = tile.thread_rank();
    const int D = A.smem_deg;
    unsigned char* base = smem + static_cast<size_t>(gid) * group_smem_bytes(D);
    double* coef = reinterpret_cast<double*>(base);            // coef[k] multiplies x^(d-k); coef[0] = 1
    double* zr = coef + (D + 1);
    double* zi = zr + D;
    double* logc = zi + D;
    int* hull = reinterpret_cast<int*>(logc + (D + 1));
    unsigned char* frozen = reinterpret_cast<unsigned char*>(hull + (D + 1));

    for (long long item = static_cast<long long>(blockIdx.x) * groups_per_cta + gid; item < A.count;
         item += static_cast<long long>(gridDim.x) * groups_per_cta) {
        const long long pid = A.index ? A.index[item] : item;
        const int d_full = A.deg[pid];
        const double* top = A.toprows + pid * A.maxdeg;
        // trailing zero coefficients are roots at 0: deflate
        int d = d_full;
        {
            int last_nz = 0;
            for (int k = l; k < d_full; k += G)
                if (top[k] != 0.0) last_nz = k + 1;
            for (int o = G / 2; o > 0; o >>= 1) last_nz = max(last_nz, tile.shfl_xor(last_nz, o));
            d = last_nz;
        }
        const int nzero = d_full - d;
        for (int k = l; k <= d; k += G) {
            const double c = (k == 0) ? 1.0 : -top[k - 1];
            coef[k] = c;
            logc[k] = (c != 0.0) ? log(fabs(c)) : -INFINITY;
        }
        tile.sync();

        // ---- initial guesses: Bini's rule.  Work with ascending powers: a_i = coef[d-i].
        // Upper convex hull of (i, log|a_i|), i = 0..d (a_0 = coef[d] != 0 after deflation, a_d = 1).
        int nh = 0;
        if (l == 0 && d > 0) {
            for (int i = 0; i <= d; ++i) {
                const double yi = logc[d - i];
                if (yi == -INFINITY) continue;
                while (nh >= 2) {
                    const int i1 = hull[nh - 2], i2 = hull[nh - 1];
                    const double y1 = logc[d - i1], y2 = logc[d - i2];
                    // keep i2 only if it lies strictly above the chord i1 -> i
                    if ((y2 - y1) * (i - i1) <= (yi - y1) * (i2 - i1)) --nh; else break;
                }
                hull[nh++] = i;
            }
        }
        nh = tile.shfl(nh, 0);
        tile.sync();
        for (int k = l; k < d; k += G) {
            // root k belongs to the hull edge [hull[e], hull[e+1]) that contains k
            int e = 0;
            while (e + 2 < nh && hull[e + 1] <= k) ++e;
            const int i1 = hull[e], i2 = hull[e + 1];
            const int m = i2 - i1;
            const double radius = exp((logc[d - i1] - logc[d - i2]) / m);
            const double ang = TWO_PI * (k - i1) / m + TWO_PI * e / d + 0.7;
            double sn, cs;
            sincos(ang, &sn, &cs);
            zr[k] = radius * cs;
            zi[k] = radius * sn;
            frozen[k] = 0;
        }
        tile.sync();

        // ---- Aberth sweeps
        int sweeps = 0;
        bool all_done = (d == 0);
        while (!all_done && sweeps < MAX_SWEEPS) {
            ++sweeps;
            bool mine_done = true;
            for (int r0 = 0; r0 < d; r0 += G) {
                const int i = r0 + l;
                cplx znew = {0.0, 0.0};
                bool active = (i < d) && !frozen[i];
                bool freeze = false, stagnant = false;
                if (active) {
                    const cplx z = {zr[i], zi[i]};
                    const double az = sqrt(z.r * z.r + z.i * z.i);
                    cplx newton;           // p/p'
                    if (az <= 1.0) {
                        cplx b = {coef[0], 0.0}, bp = {0.0, 0.0};
                        double s = fabs(coef[0]);
                        for (int k = 1; k <= d; ++k) {
                            bp = cmul(bp, z); bp.r += b.r; bp.i += b.i;
                            b = cmul(b, z); b.r += coef[k];
                            s = s * az + fabs(coef[k]);
                        }
                        const double ab = sqrt(b.r * b.r + b.i * b.i);
                        freeze = ab <= EPS * s * (d + 1) || ab == 0.0;
                        const double dp = bp.r * bp.r + bp.i * bp.i;
                        newton = (dp > 0.0) ? cmul(b, cinv(bp)) : cplx{1e-3 * (az + 1e-3), 1e-3 * (az + 1e-3)};
                    } else {
                        // p(z) = z^d q(w), w = 1/z, q(w) = sum_k coef[k] w^k; Horner from coef[d] down
                        const cplx w = cinv(z);
                        const double aw = 1.0 / az;
                        cplx b = {coef[d], 0.0}, bp = {0.0, 0.0};
                        double s = fabs(coef[d]);
                        for (int k = d - 1; k >= 0; --k) {
                            bp = cmul(bp, w); bp.r += b.r; bp.i += b.i;
                            b = cmul(b, w); b.r += coef[k];
                            s = s * aw + fabs(coef[k]);
                        }
                        const double ab = sqrt(b.r * b.r + b.i * b.i);
                        freeze = ab <= EPS * s * (d + 1) || ab == 0.0;
                        // p/p' = z / (d - w q'(w)/q(w))
                        cplx den = {static_cast<double>(d), 0.0};
                        if (ab > 0.0) {
                            const cplx t = cmul(cmul(w, bp), cinv(b));
                            den.r -= t.r; den.i -= t.i;
                        }
                        const double dd = den.r * den.r + den.i * den.i;
                        newton = (dd > 0.0) ? cmul(z, cinv(den)) : cplx{1e-3 * az, 1e-3 * az};
                    }
                    if (!freeze) {
                        cplx S = {0.0, 0.0};
                        for (int j = 0; j < d; ++j) {
                            if (j == i) continue;
                            const double dr = z.r - zr[j], di = z.i - zi[j];
                            const double q = dr * dr + di * di;
                            if (q > 0.0) {
                                const double inv = 1.0 / q;
                                S.r += dr * inv; S.i -= di * inv;
                            }
                        }
                        const cplx ns = cmul(newton, S);
                        cplx den = {1.0 - ns.r, -ns.i};
                        const double dd = den.r * den.r + den.i * den.i;
                        const cplx corr = (dd > 0.0) ? cmul(newton, cinv(den)) : newton;
                        znew = {z.r - corr.r, z.i - corr.i};
                        if (!(isfinite(znew.r) && isfinite(znew.i))) znew = {z.r * 0.5 + 1e-3, z.i * 0.5 - 1e-3};
                        // stagnation: the correction is below the resolution of z -> next sweep freezes it
                        if (corr.r * corr.r + corr.i * corr.i <= (4.0 * EPS * EPS) * (az * az)) stagnant = true;
                        mine_done = false;
                    }
                }
                tile.sync();               // everybody has read the old estimates of this round
                if (active) {
                    if (freeze) frozen[i] = 1;
                    else { zr[i] = znew.r; zi[i] = znew.i; if (stagnant) frozen[i] = 1; }
                }
                tile.sync();
            }
            all_done = tile.all(mine_done);
        }
        if (!all_done && A.fail_flag) { if (l == 0) atomicExch(A.fail_flag, 1); }

        // ---- output: [zero roots] + computed roots, optionally inverted / filtered / compacted
        double* ore = A.out_re + pid * A.maxdeg;
        double* oim = A.out_im + pid * A.maxdeg;
        int kept = 0;
        const int total = d + nzero;
        for (int r0 = 0; r0 < total; r0 += G) {
            const int i = r0 + l;
            bool keep = false;
            cplx v = {0.0, 0.0};
            if (i < total) {
                const cplx z = (i < d) ? cplx{zr[i], zi[i]} : cplx{0.0, 0.0};
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
            if (A.iters) A.iters[pid] = all_done ? sweeps : -sweeps;
        }
        tile.sync();
    }
}

template <int G>
int32_t launch_roots(RootsArgs A, cudaStream_t s) {
    if (A.count == 0) return LM_OK;
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
    long long blocks = (A.count + groups - 1) / groups;
    const long long cap = static_cast<long long>(lm::sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    kern<<<static_cast<unsigned>(blocks), groups * G, smem, s>>>(A);
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
    // bin the polynomials by degree: 8, 16 or 32 lanes per polynomial
    std::vector<long long> bins[3];
    uint64_t nroots = 0;
    int deg_max_seen = 1;
    for (int64_t k = 0; k < npoly; ++k) {
        const int d = deg[k];
        LM_REQUIRE(d >= 1 && d <= maxdeg, "lm_roots_batched: deg[%lld] = %d outside [1, %d]", static_cast<long long>(k), d, maxdeg);
        nroots += static_cast<uint64_t>(d);
        if (d > deg_max_seen) deg_max_seen = d;
        bins[d <= 8 ? 0 : (d <= 16 ? 1 : 2)].push_back(k);
    }
    cudaStream_t s = nullptr;
    const size_t ncoef = static_cast<size_t>(npoly) * maxdeg;
    void *dtop, *ddeg, *dre, *dim, *dkept, *diters, *dindex, *dflag;
    if ((rc = lm::ws_get(lm::WS_IN_A, ncoef * sizeof(double), &dtop)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_B, static_cast<size_t>(npoly) * sizeof(int), &ddeg)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_A, ncoef * sizeof(double), &dre)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_B, ncoef * sizeof(double), &dim)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_C, static_cast<size_t>(npoly) * sizeof(int), &dkept)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_OUT_D, static_cast<size_t>(npoly) * sizeof(int), &diters)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_IN_C, static_cast<size_t>(npoly) * sizeof(long long), &dindex)) != LM_OK) return rc;
    if ((rc = lm::ws_get(lm::WS_SCRATCH, 64, &dflag)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(dtop, toprows, ncoef * sizeof(double), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemcpyAsync(ddeg, deg, static_cast<size_t>(npoly) * sizeof(int), cudaMemcpyHostToDevice, s));
    LM_CUDA_TRY(cudaMemsetAsync(dflag, 0, 64, s));
    size_t off = 0;
    long long* dindex_ll = static_cast<long long*>(dindex);
    for (int b = 0; b < 3; ++b) {
        if (bins[b].empty()) continue;
        LM_CUDA_TRY(cudaMemcpyAsync(dindex_ll + off, bins[b].data(), bins[b].size() * sizeof(long long),
                                    cudaMemcpyHostToDevice, s));
        off += bins[b].size();
    }
    RootsArgs A{};
    A.toprows = static_cast<const double*>(dtop);
    A.deg = static_cast<const int*>(ddeg);
    A.maxdeg = maxdeg;
    A.invert = invert;
    A.tol = tol;
    A.out_re = static_cast<double*>(dre);
    A.out_im = static_cast<double*>(dim);
    A.n_kept = static_cast<int*>(dkept);
    A.iters = static_cast<int*>(diters);
    A.fail_flag = static_cast<int*>(dflag);
    lm::Timer tm;
    if ((rc = tm.begin(s)) != LM_OK) return rc;
    off = 0;
    int launches = 0;
    for (int b = 0; b < 3; ++b) {
        if (bins[b].empty()) continue;
        A.index = dindex_ll + off;
        A.count = static_cast<long long>(bins[b].size());
        A.smem_deg = (b == 0) ? 8 : (b == 1) ? 16 : deg_max_seen;
        if (b == 0) rc = launch_roots<8>(A, s);
        else if (b == 1) rc = launch_roots<16>(A, s);
        else rc = launch_roots<32>(A, s);
        if (rc != LM_OK) return rc;
        off += bins[b].size();
        ++launches;
    }
    float ms = 0.f;
    if ((rc = tm.end(s, &ms)) != LM_OK) return rc;
    LM_CUDA_TRY(cudaMemcpyAsync(out_re, dre, ncoef * sizeof(double), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaMemcpyAsync(out_im, dim, ncoef * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (n_kept) LM_CUDA_TRY(cudaMemcpyAsync(n_kept, dkept, static_cast<size_t>(npoly) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (iters) LM_CUDA_TRY(cudaMemcpyAsync(iters, diters, static_cast<size_t>(npoly) * sizeof(int), cudaMemcpyDeviceToHost, s));
    int failed = 0;
    LM_CUDA_TRY(cudaMemcpyAsync(&failed, dflag, sizeof(int), cudaMemcpyDeviceToHost, s));
    LM_CUDA_TRY(cudaStreamSynchronize(s));
    if (stats) {
        stats->items = static_cast<uint64_t>(npoly);
        stats->work_units = nroots;
        stats->kernel_ms = ms;
        stats->launches = launches;
    }
    if (failed)
        return lm::fail(LM_E_NOCONV, "lm_roots_batched: Aberth iteration did not converge within %d sweeps for some polynomial "
                        "(iters < 0 marks them; outputs hold the last estimates)", MAX_SWEEPS);
    return LM_OK;
}

}  // extern "C"

// emrifd.cu -- sm_100a kernels + C-ABI for the FD EMRI mode-sum / likelihood hot path.
//
// Kernels (DESIGN.md has the layout, rooflines and algorithmic bytes of each):
//   spline_build_kernel   A3  batched not-a-knot cubic spline (shared tridiagonal factorisation in smem,
//                             one RHS per thread, coalesced quad stores)
//   segment_kernel        A4  per-mode monotone-branch segmentation + bit-exact bin ranges
//   group_index / group_combine  one stationary point per (m, n) group: combined amplitude quads
//   piece_build_kernel    per (group branch, spline segment): combined cubic of f_mn, phase constants, amplitude quads and a
//                             verified degree-5 inverse interpolant x(f) -- the "piece table" the mode sum reads through the TMA
//   empty_tile_kernel     A5-A7 pass 1: tiles no harmonic touches are zero-filled (a store stream at the HBM write rate),
//                             the others are queued
//   mode_sum_kernel       A5-A7 (+A11 fused) pass 2, one warp-specialised CTA per SM fed by the tile queue (PERSISTENT = false:
//                             plain grid for small launches): a producer warp scans the walker's cached records and streams
//                             (header + TMA-copied piece) sub-entries through an mbarrier ring; 19 consumer warps own two
//                             adjacent (+f,-f) bin pairs per thread with all accumulators in registers: root = interpolant + one
//                             Newton step, SPA factor, phase in cycles, no atomics, no CTA barrier; h+/hx split + scale +
//                             rotation fused into the store, optional fused |d - h|^2 / <d|h> / <h|h> reduction
//                             (like_finalize_kernel adds the per-warp partials in a fixed order)
//   inner_product / loglike kernels  A10/A11 on materialised arrays
//   ylm / synth_amplitude / mode_select / compact_* kernels  the producers right before the path
//
// The per-harmonic formula follows the reference's statement of it
// (Tutorial_FD_construction_single_mode.ipynb:548-623, cell 26); see include/emrifd.h for the
// reference interface each entry point replaces.  No tensor cores: nothing here is a contraction.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <limits.h>
#include <new>

#include "../../include/emrifd.h"
#include "k13_tables.h"

#define MAXBR EMRIFD_MAX_BRANCHES
// mode-sum CTA: SUM_CW consumer warps (bin owners: they evaluate and read out) + one producer warp (record scan, sub-entry fill).
// One CTA per SM: 20 warps = 5 per SM sub-partition at 96 registers (5 x 32 x 96 = 15 360 of the sub-partition's 16 384
// registers): the evaluation loop keeps its eight accumulators and both bins' temporaries in registers without spilling.
// Measured on the bench batch (kernel ms): 19 + 1 warps @ 96 regs 3.55 | 23 + 1 @ 80 (spills) 3.74 | 2 CTAs x (11 + 1) @ 80 3.92 |
// 27 + 1 @ 72 4.12 | 31 + 1 @ 64 4.30.
#ifndef SUM_CW
#define SUM_CW 19
#endif
#define SUM_CT (SUM_CW * 32)      /* consumer threads */
#define SUM_THREADS (SUM_CT + 32) /* + the producer warp */
#ifndef SUM_BPT
#define SUM_BPT 2 /* consecutive bins per consumer thread: one pair, evaluated together in straight-line code */
#endif
#ifndef SUM_MAXNREG
#define SUM_MAXNREG 96
#endif
#ifndef SUM_RING
#define SUM_RING 4 /* passes (of up to SUM_SUBCAP sub-entries) in flight between the producer warp and the consumer warps */
#endif
#ifndef SUM_RCAP
#define SUM_RCAP 512 /* group records (4 per (m, n) group) the producer warp caches in shared memory per walker */
#endif
#define SUM_CHUNK 256 /* group records per chunk of the work-list (chunk_range_kernel's hull granularity) */
#define SUM_BOUNDS __maxnreg__(SUM_MAXNREG)
#define SUM_TILE (SUM_CT * SUM_BPT) /* bins per tile: the tile size every API-visible quantity refers to (emrifd_tile_bins) */
#define SEG_THREADS 128
#define SMEM_PER_KNOT 3 /* doubles staged per knot by the mode sum: t, f_phi, f_r */

struct emrifd_handle {
    int device;
    cudaStream_t stream;
    char err[512];
    int *d_status;              // device error word
    emrifd_walker_t *d_walkers; // device copy of the walker descriptors
    void *d_queue;              // [2 x u64 control words][B * ntiles] non-empty tile queue of the mode sum
    int64_t queue_cap;
    int num_sms;
    int64_t walkers_cap;
    emrifd_walker_t *h_stage[4]; // pinned staging ring
    cudaEvent_t stage_ev[4];
    int64_t stage_cap;
    int stage_next;
    double *d_partial; // likelihood partial sums
    int64_t partial_cap;
    long long *d_chunk; // per-chunk bin hulls
    int64_t chunk_cap;
    double *d_tiledd;   // per-tile sum |d~|^2 of the injected data (tiles of SUM_TILE bins)
    int64_t tiledd_cap;
    const double *d_data; // whitened data [2][n]
    const double *d_wfac; // noise factor  [2][n]
    int64_t n_data;
    // host-buffer path workspace
    char *d_ws; int64_t ws_cap;
    char *h_ws; int64_t h_ws_cap;
    int64_t launches;
    int max_dyn_smem;
    // per-walker status words, (m, n) group index and combined amplitude quads (workspace of the mode sum)
    int *d_wstatus; int64_t wstatus_cap;
    int *d_leader; int64_t leader_cap;
    int *d_gmem; int64_t gmem_cap;
    int *d_goff; int64_t goff_cap;
    int *d_gcount; int64_t gcount_cap;
    double *d_gq; int64_t gq_cap;
    void *d_pieces; int64_t pieces_cap; // piece table of the batch being summed
    char *d_selovf; int64_t selovf_cap; // mode selection: samples left to the full-capacity launch
    int64_t tot_modes, tot_teuk; // packed sizes of the batch validated last (sum K, sum L*K)
    int k13_few;
    // kernel timing
    int timing;
    cudaEvent_t ev_a[64], ev_b[64], ev_m[64], ev_s[64]; // before the tile classification, after the join of the launch pair, around mode_sum_kernel
    cudaStream_t aux;                         // empty_tile_kernel's stream (forked after the classification, joined before the finalize)
    cudaEvent_t ev_fork, ev_join;
    int overlap_mode;                         // EMRIFD_OVERLAP=0: zero-fill after the sum on the same stream (A/B runs); default: underneath it
    int ev_n;
    double sum_ms, sum_ms_main;
    int64_t sum_launches;
};

static int set_err(emrifd_handle *h, int code, const char *msg) {
    if (h) snprintf(h->err, sizeof(h->err), "%s", msg);
    return code;
}
#define CUDA_TRY(h, call)                                                                       \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            if (h) snprintf((h)->err, sizeof((h)->err), "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return EMRIFD_ERR_CUDA;                                                             \
        }                                                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------
// exactly-rounded helpers: spline build + segmentation must round every operation
// (no FMA contraction) so that they reproduce the oracle's index sets bit for bit.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double rmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double radd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double rsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double rdiv(double a, double b) { return __ddiv_rn(a, b); }

// ==========================================================================================
// A3: batched not-a-knot spline build
// ==========================================================================================
struct SplineParams {
    const emrifd_walker_t *w; // device, batched mode
    const double *t, *teuk, *trk0, *trk1, *trk2, *trk3;
    double *coeff;
    // generic single-spline mode
    const double *ygen;
    long long rs, ks;
    int Lgen, Rgen;
};

// One CTA = 32 rows of one walker.  Warp 0: one row per lane; warp 1 lane 0 factorises the (shared) tridiagonal
// matrix while warp 0 stages its rows into shared memory, so the serial pivot chain overlaps the loads.
// TILED: y and the forward-sweep intermediates live in shared memory [L][32] (fits for L <~ 400); otherwise
// y is re-read from global memory and the intermediates are parked in the c1 slot of the output.
// Every operation is individually rounded in the oracle's order (oracle/emrifd_oracle.c orc_spline_build).
#define SPL_ROWS 32
#define SPL_CTA 64
template <bool GENERIC, bool TILED>
__global__ void __launch_bounds__(SPL_CTA) spline_build_kernel(SplineParams p, int *status) {
    extern __shared__ double sm[];
    int L, K, R;
    const double *t;
    double *coeff;
    long long teuk_off = 0, knot_off = 0;
    if (GENERIC) {
        L = p.Lgen; R = p.Rgen; K = 0; t = p.t; coeff = p.coeff;
    } else {
        const emrifd_walker_t wd = p.w[blockIdx.y];
        L = wd.L; K = wd.K; R = 2 * K + 4;
        t = p.t + wd.knot_off; coeff = p.coeff + wd.coeff_off;
        teuk_off = wd.teuk_off; knot_off = wd.knot_off;
    }
    const int r0 = blockIdx.x * SPL_ROWS;
    if (r0 >= R) return;
    double *sh = sm, *srh = sm + L, *sw = sm + 2 * L, *sinv = sm + 3 * L, *scup = sm + 4 * L;
    double *sy = sm + 5 * L, *sb = sy + (TILED ? L * SPL_ROWS : 0);
    __shared__ int bad;
    const int tid = threadIdx.x, lr = tid & 31, r = r0 + lr;
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int j = tid; j < L - 1; j += SPL_CTA) {
        const double hj = rsub(t[j + 1], t[j]);
        sh[j] = hj;
        srh[j] = rdiv(1.0, hj);
        if (!(hj > 0.0)) bad = 1;
    }
    __syncthreads();
    if (bad) {
        if (tid == 0) atomicMin(status, EMRIFD_ERR_KNOT_ORDER);
        return;
    }
    const double dd0 = rsub(t[2], t[0]), ddn = rsub(t[L - 1], t[L - 3]);
    // row source
    const double *yb = nullptr;
    long long ks = 1;
    if (r < R) {
        if (GENERIC) { yb = p.ygen + (long long)r * p.rs; ks = p.ks; }
        else if (r < K)     { yb = p.teuk + 2 * (teuk_off + r); ks = 2 * K; }
        else if (r < 2 * K) { yb = p.teuk + 2 * (teuk_off + (r - K)) + 1; ks = 2 * K; }
        else {
            const int q = r - 2 * K;
            const double *b = q == 0 ? p.trk0 : q == 1 ? p.trk1 : q == 2 ? p.trk2 : p.trk3;
            yb = b + knot_off; ks = 1;
        }
    }
    if (tid == 32) {
        // LU without pivoting: inv_0 = 1/d_0; w_i = a_i*inv_{i-1}; d'_i = d_i - w_i*c_{i-1}; inv_i = 1/d'_i
        double dprev = sh[1];
        scup[0] = dd0; sinv[0] = rdiv(1.0, dprev); sw[0] = 0.0;
        double invp = sinv[0], cupp = dd0;
        for (int i = 1; i < L; i++) {
            double a, d, c;
            if (i < L - 1) { a = sh[i]; d = rmul(2.0, radd(sh[i - 1], sh[i])); c = sh[i - 1]; }
            else           { a = ddn;   d = sh[L - 3];                         c = 0.0; }
            const double wi = rmul(a, invp);
            dprev = rsub(d, rmul(wi, cupp));
            invp = rdiv(1.0, dprev);
            scup[i] = c; sw[i] = wi; sinv[i] = invp;
            cupp = c;
        }
    } else if (TILED && tid < 32 && r < R) {
        for (int j = 0; j < L; j++) sy[j * SPL_ROWS + lr] = yb[(long long)j * ks];
    }
    __syncthreads();
    if (tid >= 32 || r >= R) return;
#define Y(j) (TILED ? sy[(j) * SPL_ROWS + lr] : yb[(long long)(j) * ks])
#define CO(j, c) coeff[((long long)(j) * R + r) * 4 + (c)]
#define BPW(i, v) do { if (TILED) sb[(i) * SPL_ROWS + lr] = (v); else CO(i, 1) = (v); } while (0)
#define BPR(i) (TILED ? sb[(i) * SPL_ROWS + lr] : CO(i, 1))
    // forward sweep
    double sprev, dm, dp;
    {
        const double y0 = Y(0), y1 = Y(1), y2 = Y(2);
        dm = rmul(rsub(y1, y0), srh[0]);
        dp = rmul(rsub(y2, y1), srh[1]);
        const double num = radd(rmul(rmul(radd(sh[0], rmul(2.0, dd0)), sh[1]), dm), rmul(rmul(sh[0], sh[0]), dp));
        sprev = rdiv(num, dd0);
        BPW(0, sprev);
    }
    {
        double yc = Y(1);
        for (int i = 1; i < L - 1; i++) {
            const double yp = Y(i + 1);
            dp = rmul(rsub(yp, yc), srh[i]);
            const double b = rmul(3.0, radd(rmul(sh[i], dm), rmul(sh[i - 1], dp)));
            sprev = rsub(b, rmul(sw[i], sprev));
            BPW(i, sprev);
            if (i < L - 2) dm = dp;
            yc = yp;
        }
    }
    {
        const double hl2 = sh[L - 2], hl3 = sh[L - 3];
        const double num = radd(rmul(rmul(hl2, hl2), dm), rmul(rmul(radd(rmul(2.0, ddn), hl2), hl3), dp));
        const double b = rdiv(num, ddn);
        sprev = rsub(b, rmul(sw[L - 1], sprev));
    }
    // back substitution + coefficients
    double snext = rmul(sprev, sinv[L - 1]);
    double ynext = Y(L - 1);
    *reinterpret_cast<double4 *>(&CO(L - 1, 0)) = make_double4(ynext, snext, 0.0, 0.0);
    for (int i = L - 2; i >= 0; i--) {
        const double bi = BPR(i);
        const double si = rmul(rsub(bi, rmul(scup[i], snext)), sinv[i]);
        const double yi = Y(i);
        const double rhi = srh[i];
        const double dl = rmul(rsub(ynext, yi), rhi);
        const double tau = rmul(rsub(radd(si, snext), rmul(2.0, dl)), rhi);
        const double c2 = rsub(rmul(rsub(dl, si), rhi), tau);
        const double c3 = rmul(tau, rhi);
        *reinterpret_cast<double4 *>(&CO(i, 0)) = make_double4(yi, si, c2, c3);
        snext = si; ynext = yi;
    }
#undef Y
#undef CO
#undef BPW
#undef BPR
}

__global__ void spline_eval_kernel(const double *__restrict__ t, const double *__restrict__ coeff, int L, int R,
                                   const double *__restrict__ tnew, long long n, double *__restrict__ out) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double tq = tnew[q];
    int lo = 0, hi = L - 2;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (t[mid] <= tq) lo = mid; else hi = mid - 1; }
    double x = rsub(tq, t[lo]);
    for (int r = blockIdx.y; r < R; r += gridDim.y) {
        const double4 c = *reinterpret_cast<const double4 *>(coeff + ((long long)lo * R + r) * 4);
        // same rounding sequence as the oracle (plain Horner, no contraction)
        out[(long long)r * n + q] = radd(c.x, rmul(x, radd(c.y, rmul(x, radd(c.z, rmul(x, c.w))))));
    }
}

// ==========================================================================================
// frequency grid helpers
// ==========================================================================================
struct Grid {
    long long N, zero;
    double val;
    const double *fpos;
};
__device__ __forceinline__ double grid_f(const Grid &g, long long i) {
    long long k = i - g.zero;
    if (g.fpos) return k >= 0 ? g.fpos[k] : -g.fpos[-k];
    return rmul((double)k, g.val);
}
// smallest i in [0,N] with f_i >= F (strict=0) or f_i > F (strict=1)
__device__ long long grid_lower(const Grid &g, double F, int strict) {
    if (!(F == F)) return g.N; // NaN track: no bin (keeps the walks below finite)
    double df = g.fpos ? rdiv(g.fpos[g.zero], (double)g.zero) : g.val;
    double e = radd(rdiv(F, df), (double)g.zero);
    long long i;
    if (!(e > 0.0)) i = 0; else if (e >= (double)g.N) i = g.N; else i = (long long)e;
    while (i > 0) { double f = grid_f(g, i - 1); if (strict ? (f > F) : (f >= F)) i--; else break; }
    while (i < g.N) { double f = grid_f(g, i); if (strict ? (f > F) : (f >= F)) break; i++; }
    return i;
}

// ==========================================================================================
// A4: segmentation -- one thread per (walker, mode)
// ==========================================================================================
struct SegParams {
    const emrifd_walker_t *w;
    const double *t, *coeff;
    const int *m, *n;
    emrifd_branch_t *br;
    long long *n_eval; // [B][2] or NULL
    int *wstatus;      // [B] per-walker status word (zeroed by the host before the launch)
    Grid g;
};

__device__ __forceinline__ void push_sub(emrifd_branch_t *br, int &nb, int &overflow, int k, int j,
                                         double xa, double Fa, double xb, double Fb) {
    int dir = (Fb > Fa) - (Fb < Fa);
    if (dir == 0 || !(xb > xa)) return;
    if (nb > 0 && br[nb - 1].dir == dir) { br[nb - 1].jb = j; br[nb - 1].xb = xb; br[nb - 1].Fb = Fb; return; }
    if (nb >= MAXBR) { overflow = 1; return; }
    emrifd_branch_t b;
    b.mode = k; b.dir = dir; b.ja = j; b.jb = j; b.closed_end = 0; b.pad = 0;
    b.start = 0; b.end = -1; b.xa = xa; b.xb = xb; b.Fa = Fa; b.Fb = Fb;
    br[nb++] = b;
}

// Two phases per CTA (a group of `mpc` modes of one walker):
//   A  one thread per (mode, segment): roots of fdot inside the segment and f there (sqrt / divisions in parallel);
//   B  one thread per mode: merge the sub-intervals into monotone branches and turn their frequency ranges into bin
//      ranges.  Both phases use exactly the oracle's operations (individually rounded), so the work-list is bit-exact.
__global__ void __launch_bounds__(SEG_THREADS) segment_kernel(SegParams p, int *status, int mpc) {
    extern __shared__ double sm[]; // t[L] | (f_phi quad, f_r quad)[L][8] | xr[mpc][L][2] | Fx[mpc][L][2] | nr[mpc][L]
    const emrifd_walker_t wd = p.w[blockIdx.y];
    const int L = wd.L, K = wd.K, R = 2 * K + 4;
    const int k0 = blockIdx.x * mpc;
    if (k0 >= K) return;
    const int nm = (K - k0) < mpc ? (K - k0) : mpc;
    const double *t = p.t + wd.knot_off;
    const double *coeff = p.coeff + wd.coeff_off;
    double *st = sm, *sq = sm + L, *sxr = sm + 9 * L, *sFx = sxr + 2 * mpc * L;
    int *snr = reinterpret_cast<int *>(sFx + 2 * mpc * L);
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    for (int i = threadIdx.x; i < L; i += SEG_THREADS) st[i] = t[i];
    for (int i = threadIdx.x; i < 2 * L; i += SEG_THREADS) {
        const int jj = i >> 1, q = i & 1;
        const double4 c = *reinterpret_cast<const double4 *>(coeff + ((long long)jj * R + 2 * K + q) * 4);
        double *d = sq + jj * 8 + q * 4;
        d[0] = c.x; d[1] = c.y; d[2] = c.z; d[3] = c.w;
    }
    __syncthreads();
    // knots that are not strictly increasing (or NaN): the spline kernel has refused this walker and its coefficients are
    // undefined.  The walker gets an empty work-list and its status word; the other walkers of the batch are unaffected.
    for (int i = threadIdx.x; i < L - 1; i += SEG_THREADS) if (!(st[i + 1] > st[i])) s_bad = 1;
    __syncthreads();
    if (s_bad) {
        if ((int)threadIdx.x < nm) {
            emrifd_branch_t b;
            b.mode = k0 + threadIdx.x; b.dir = 0; b.ja = 0; b.jb = 0; b.closed_end = 0; b.pad = 0;
            b.start = 0; b.end = -1; b.xa = 0; b.xb = 0; b.Fa = 0; b.Fb = 0;
            for (int q = 0; q < MAXBR; q++) p.br[(wd.mode_off + k0 + threadIdx.x) * MAXBR + q] = b;
        }
        if (threadIdx.x == 0) { atomicMin(status, EMRIFD_ERR_KNOT_ORDER); atomicMin(&p.wstatus[blockIdx.y], EMRIFD_ERR_KNOT_ORDER); }
        return;
    }
    // ---- phase A ----
    for (int item = threadIdx.x; item < nm * (L - 1); item += SEG_THREADS) {
        const int kl = item / (L - 1), j = item - kl * (L - 1);
        const double dm = (double)p.m[wd.mode_off + k0 + kl], dn = (double)p.n[wd.mode_off + k0 + kl];
        const double *c = sq + j * 8;
        const double hj = rsub(st[j + 1], st[j]);
        const double c0 = radd(rmul(dm, c[0]), rmul(dn, c[4]));
        const double c1 = radd(rmul(dm, c[1]), rmul(dn, c[5]));
        const double c2 = radd(rmul(dm, c[2]), rmul(dn, c[6]));
        const double c3 = radd(rmul(dm, c[3]), rmul(dn, c[7]));
        double xr[2] = {0.0, 0.0};
        int nr = 0;
        const double qa = rmul(3.0, c3), qb = rmul(2.0, c2), qc = c1;
        if (qa == 0.0) {
            if (qb != 0.0) { double r0 = rdiv(-qc, qb); if (r0 > 0.0 && r0 < hj) xr[nr++] = r0; }
        } else {
            double disc = rsub(rmul(qb, qb), rmul(rmul(4.0, qa), qc));
            if (disc >= 0.0) {
                double sqd = __dsqrt_rn(disc);
                double qq = (qb >= 0.0) ? rmul(-0.5, radd(qb, sqd)) : rmul(-0.5, rsub(qb, sqd));
                double r0 = rdiv(qq, qa);
                double r1 = (qq != 0.0) ? rdiv(qc, qq) : r0;
                if (r0 > r1) { double tmp = r0; r0 = r1; r1 = tmp; }
                if (r0 > 0.0 && r0 < hj) xr[nr++] = r0;
                if (r1 > 0.0 && r1 < hj && r1 != r0) xr[nr++] = r1;
            }
        }
        const int o = (kl * L + j) * 2;
        snr[kl * L + j] = nr;
        for (int q = 0; q < 2; q++) {
            const double x = xr[q];
            sxr[o + q] = x;
            sFx[o + q] = radd(c0, rmul(x, radd(c1, rmul(x, radd(c2, rmul(x, c3))))));
        }
    }
    __syncthreads();
    // ---- phase B ----
    if ((int)threadIdx.x >= nm) return;
    const int kl = threadIdx.x, k = k0 + kl;
    const int mi = p.m[wd.mode_off + k], ni = p.n[wd.mode_off + k];
    const double dm = (double)mi, dn = (double)ni;
    emrifd_branch_t *out = p.br + (wd.mode_off + k) * MAXBR;
    emrifd_branch_t br[MAXBR];
    for (int q = 0; q < MAXBR; q++) {
        br[q].mode = k; br[q].dir = 0; br[q].ja = 0; br[q].jb = 0; br[q].closed_end = 0; br[q].pad = 0;
        br[q].start = 0; br[q].end = -1; br[q].xa = 0; br[q].xb = 0; br[q].Fa = 0; br[q].Fb = 0;
    }
    int nb = 0, overflow = 0;
    double Fk = radd(rmul(dm, sq[0]), rmul(dn, sq[4]));
    for (int j = 0; j < L - 1; j++) {
        const double *c = sq + j * 8;
        const double hj = rsub(st[j + 1], st[j]);
        const double Fnext = radd(rmul(dm, c[8]), rmul(dn, c[12]));
        const int nr = snr[kl * L + j];
        double xa = 0.0, Fa = Fk;
        for (int q = 0; q < nr; q++) {
            const double x = sxr[(kl * L + j) * 2 + q], Fx = sFx[(kl * L + j) * 2 + q];
            push_sub(br, nb, overflow, k, j, xa, Fa, x, Fx);
            xa = x; Fa = Fx;
        }
        push_sub(br, nb, overflow, k, j, xa, Fa, hj, Fnext);
        Fk = Fnext;
    }
    if (nb > 0) br[nb - 1].closed_end = 1;
    long long evals = 0;
    for (int q = 0; q < nb; q++) {
        emrifd_branch_t &b = br[q];
        double Flo = b.dir > 0 ? b.Fa : b.Fb, Fhi = b.dir > 0 ? b.Fb : b.Fa;
        int lo_strict = (b.dir > 0) ? 0 : !b.closed_end;
        int hi_strict = (b.dir > 0) ? !b.closed_end : 0;
        b.start = grid_lower(p.g, Flo, lo_strict);
        b.end = grid_lower(p.g, Fhi, hi_strict ? 0 : 1) - 1;
        if (b.end >= b.start) evals += b.end - b.start + 1;
    }
    for (int q = 0; q < MAXBR; q++) out[q] = br[q];
    if (overflow) { atomicMin(status, EMRIFD_ERR_BRANCHES); atomicMin(&p.wstatus[blockIdx.y], EMRIFD_ERR_BRANCHES); }
    if (p.n_eval && evals) {
        atomicAdd((unsigned long long *)&p.n_eval[2 * blockIdx.y], (unsigned long long)evals);
        atomicAdd((unsigned long long *)&p.n_eval[2 * blockIdx.y + 1], (unsigned long long)(evals * (mi > 0 ? 2 : 1)));
    }
}

// ==========================================================================================
// SPA factor  R(X) = K_{1/3}(-iX) e^{-iX} sqrt(2X/pi) e^{-i pi/4}
// ==========================================================================================
__device__ __noinline__ void k13_mid(double X, double &re, double &im) {
    // 1 <= X < 32: octave polynomial in s = 4/mant - 3
    int ex;
    double mant = frexp(X, &ex); // X = mant * 2^ex, mant in [0.5,1)
    int oct = ex - 1;            // X in [2^oct, 2^(oct+1))
    oct = oct < 0 ? 0 : (oct > K13_NOCT - 1 ? K13_NOCT - 1 : oct);
    double s = 2.0 / mant - 3.0; // 4/(2 mant) - 3
    double pr = k13_poly_re[oct][K13_DEG], pi = k13_poly_im[oct][K13_DEG];
#pragma unroll
    for (int k = K13_DEG - 1; k >= 0; k--) { pr = fma(pr, s, k13_poly_re[oct][k]); pi = fma(pi, s, k13_poly_im[oct][k]); }
    re = pr; im = pi;
}
__device__ __noinline__ void k13_small_S(double X, double &re, double &im) {
    // S(X) = R(X)/X^{1/6}, X < 1 (turnover regime)
    const double q = -0.25 * X * X;
    double A = k13_ser_a[K13_NSER - 1], B = k13_ser_b[K13_NSER - 1];
#pragma unroll
    for (int k = K13_NSER - 2; k >= 0; k--) { A = fma(A, q, k13_ser_a[k]); B = fma(B, q, k13_ser_b[k]); }
    const double c13 = 1.2599210498948732; // 2^{1/3}
    double x13 = cbrt(X), x23 = x13 * x13;
    double cb = c13 * B, ca = x23 * A / c13;
    const double s3h = 0.8660254037844386;
    double ure = s3h * (cb - ca), uim = 0.5 * (cb + ca);
    double sn, cs;
    sincos(-(X + 0.7853981633974483), &sn, &cs);
    const double pref = 1.4472025091165353; // sqrt(2 pi / 3)
    re = pref * (ure * cs - uim * sn);
    im = pref * (ure * sn + uim * cs);
}
// G(fdot, fddot) = i fdot/|fddot| (2/sqrt3) K_{1/3}(-iX) e^{-iX},  X = 2 pi fdot^3 / (3 fddot^2)
//               = e^{+-i 3pi/4} R(|X|)/sqrt|fdot|   (conjugated for fdot < 0): evaluated inside eval_sub (common path) and
// spa_fix (X < 1024); k13_mid / k13_small_S are its rarer ranges.
// ==========================================================================================
// A5-A7 (+A11): bin-owner mode sum
// ==========================================================================================
struct SumParams {
    const emrifd_walker_t *w;
    const double *t, *coeff;
    const int *m, *n;
    const double2 *ylm;
    const emrifd_branch_t *br;
    Grid g;
    long long j_lo, j_cnt;
    int include_minus_m, mask_positive;
    double2 *hp, *hc;
    const double2 *dw; // [2][n_data]
    const double *wf;  // [2][n_data]
    long long n_data;
    double *partial;   // [B][ntiles][3]
    const long long *chunk_rng; // [B][cpw][2] hull of positive bins per record chunk
    int cpw;
    const double *tile_dd;      // [ceil(n_data/SUM_TILE)] sum |d~|^2 per tile of the injected data, or NULL
    long long tile_first, tile_stride; // this launch owns tiles tile_first + i * tile_stride of [j_lo, j_lo + j_cnt) (0, 1 = all)
    int no_empty;               // 1: treat every tile as non-empty (likelihood on a slice that is not tile-aligned)
    int ntiles;                 // tiles per walker (= ceil(j_cnt / SUM_TILE))
    unsigned long long *queue;  // [B * ntiles] non-empty tiles as (walker << 32 | tile), filled by empty_tile_kernel
    unsigned int *qctl;         // [0] number of queued tiles, [1] next item handed to a persistent mode_sum CTA
    unsigned char *tile_flag;   // [B * ntiles] 1: the tile has work for mode_sum_kernel, 0: empty (empty_tile_kernel's)
    // (m, n) groups (group_kernel): one stationary point per (group, bin)
    const int *leader;          // [sum K] walker block at mode_off: first member (mode index) of each group
    const int *gcount;          // [B] groups per walker
    const double *gq;           // walker block at 16 * teuk_off: [L][G][16] combined amplitude quads
    const int *wstatus;         // [B] per-walker status (0 = ok): failed walkers give zeros / a NaN likelihood, the rest of the batch is unaffected
    int k13_few;                // 1: FastEMRIWaveforms-compatible K_1/3 evaluation (emrifd_set_k13_mode)
    const struct Piece *pieces; // piece table (piece_build_kernel): row (mode_off * MAXBR + group record), lstride entries per row
    int lstride;
};


// ---- fast reciprocal for Newton steps: 24-bit seed is enough (the iteration is self-correcting) ----
// polynomial coefficients live in the constant bank so that DFMA takes them as c[bank][offset] operands
// (as immediates each costs two UMOVs per use: 48 of 365 issued instructions per evaluation in v3)
__constant__ double c_sin[8] = {6.283185307179586, -41.34170224039976, 81.60524927607506, -76.70585975306139,
                                42.058693944897655, -15.09464257682299, 3.819952584848282, -0.7181223017785006};
__constant__ double c_cos[9] = {1.0, -19.739208802178716, 64.9393940226683, -85.45681720669373, 60.24464137187666,
                                -26.4262567833744, 7.903536371318469, -1.714390711088672, 0.28200596845579123};
// (MUFU.RCP64H, ~20 bits; the Newton iteration it feeds is self-correcting)
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
}
// 1/sqrt(a): MUFU.RSQ64H seed + two Newton steps (full double accuracy for normal a > 0)
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double ha = 0.5 * a;
#if OPT_HALLEY
    const double e = fma(-ha * y, y, 0.5);           // e = (1 - a y^2)/2;  1/sqrt(a) = y (1 + e + 1.5 e^2 + O(e^3))
    return fma(y * e, fma(1.5, e, 1.0), y);
#else
    double e = fma(-ha * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-ha * y, y, 0.5);
    return fma(y, e, y);
#endif
}
#ifndef OPT_RINT
#define OPT_RINT 0   /* round-to-nearest-integer by adding 1.5 * 2^52 (two DADDs) instead of FRND / F2I (conversion pipe, quarter rate) */
#endif
#ifndef OPT_SIGN
#define OPT_SIGN 1   /* multiplications by +-1 as a sign-bit XOR (integer pipe) */
#endif
#ifndef OPT_HALLEY
#define OPT_HALLEY 1 /* rsqrt: one cubic (Halley) step instead of two Newton steps */
#endif
#ifndef OPT_SPLIT
#define OPT_SPLIT 1  /* stage 1: bins 1 and 2 both start from bin 0 (two independent Newton chains), bin 3 from bin 2 */
#endif
#define RINT_MAGIC 6755399441055744.0 /* 1.5 * 2^52: (x + M) - M = rint(x) for |x| < 2^51, and the low word of x + M is (int)rint(x) */
__device__ __forceinline__ double rint_fast(double x) {
#if OPT_RINT
    return __dadd_rn(__dadd_rn(x, RINT_MAGIC), -RINT_MAGIC);
#else
    return rint(x);
#endif
}
// x * s for s = +-1 given as its sign mask (0 or 0x80000000 for the high word)
__device__ __forceinline__ double flip_sign(double x, unsigned int mask_hi) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask_hi, __double2loint(x));
}
// sin and cos of 2*pi*c for |c| <~ 2^20: quarter-turn reduction (exact) + degree-15/16 polynomials in r = c - q/4
__device__ __forceinline__ void sincos_cycles(double c, double &sn, double &cs) {
#if OPT_RINT
    const double tq = fma(4.0, c, RINT_MAGIC);
    const int qi = __double2loint(tq);
    const double q = __dadd_rn(tq, -RINT_MAGIC);
#else
    const double q = rint(4.0 * c);
    const int qi = (int)q;
#endif
    const double r = fma(-0.25, q, c);
    const double r2 = r * r;
    double ps = c_sin[7], pc = c_cos[8];
#pragma unroll
    for (int k = 6; k >= 0; k--) { ps = fma(ps, r2, c_sin[k]); pc = fma(pc, r2, c_cos[k + 1]); }
    ps *= r;
    pc = fma(pc, r2, c_cos[0]);
    const bool swap = qi & 1;
    const double a = swap ? pc : ps, b = swap ? ps : pc;
    sn = (qi & 2) ? -a : a;
    cs = ((qi + 1) & 2) ? -b : b;
}

__constant__ double c_rot64[128] = {
    1.0, 0.0, 0.9951847266721969, 0.0980171403295606, 0.9807852804032304, 0.19509032201612828, 0.9569403357322088, 0.2902846772544624,
    0.9238795325112867, 0.3826834323650898, 0.881921264348355, 0.47139673682599764, 0.8314696123025452, 0.5555702330196022, 0.773010453362737, 0.6343932841636455,
    0.7071067811865476, 0.7071067811865476, 0.6343932841636455, 0.773010453362737, 0.5555702330196022, 0.8314696123025452, 0.47139673682599764, 0.881921264348355,
    0.3826834323650898, 0.9238795325112867, 0.2902846772544624, 0.9569403357322088, 0.19509032201612828, 0.9807852804032304, 0.0980171403295606, 0.9951847266721969,
    2.0670321098263988e-43, 1.0, -0.0980171403295606, 0.9951847266721969, -0.19509032201612828, 0.9807852804032304, -0.2902846772544624, 0.9569403357322088,
    -0.3826834323650898, 0.9238795325112867, -0.47139673682599764, 0.881921264348355, -0.5555702330196022, 0.8314696123025452, -0.6343932841636455, 0.773010453362737,
    -0.7071067811865476, 0.7071067811865476, -0.773010453362737, 0.6343932841636455, -0.8314696123025452, 0.5555702330196022, -0.881921264348355, 0.47139673682599764,
    -0.9238795325112867, 0.3826834323650898, -0.9569403357322088, 0.2902846772544624, -0.9807852804032304, 0.19509032201612828, -0.9951847266721969, 0.0980171403295606,
    -1.0, 4.1340642196527976e-43, -0.9951847266721969, -0.0980171403295606, -0.9807852804032304, -0.19509032201612828, -0.9569403357322088, -0.2902846772544624,
    -0.9238795325112867, -0.3826834323650898, -0.881921264348355, -0.47139673682599764, -0.8314696123025452, -0.5555702330196022, -0.773010453362737, -0.6343932841636455,
    -0.7071067811865476, -0.7071067811865476, -0.6343932841636455, -0.773010453362737, -0.5555702330196022, -0.8314696123025452, -0.47139673682599764, -0.881921264348355,
    -0.3826834323650898, -0.9238795325112867, -0.2902846772544624, -0.9569403357322088, -0.19509032201612828, -0.9807852804032304, -0.0980171403295606, -0.9951847266721969,
    2.2338764406549882e-41, -1.0, 0.0980171403295606, -0.9951847266721969, 0.19509032201612828, -0.9807852804032304, 0.2902846772544624, -0.9569403357322088,
    0.3826834323650898, -0.9238795325112867, 0.47139673682599764, -0.881921264348355, 0.5555702330196022, -0.8314696123025452, 0.6343932841636455, -0.773010453362737,
    0.7071067811865476, -0.7071067811865476, 0.773010453362737, -0.6343932841636455, 0.8314696123025452, -0.5555702330196022, 0.881921264348355, -0.47139673682599764,
    0.9238795325112867, -0.3826834323650898, 0.9569403357322088, -0.2902846772544624, 0.9807852804032304, -0.19509032201612828, 0.9951847266721969, -0.0980171403295606};
__constant__ double c_sin64[4] = {6.283185307179586, -41.34170224039976, 81.60524927607506, -76.70585975306139};
__constant__ double c_cos64[4] = {-19.739208802178716, 64.9393940226683, -85.45681720669373, 60.24464137187666};
// sin and cos of 2*pi*c by a 1/64-turn table (shared memory, filled from c_rot64) and degree-7/8 Taylor polynomials on the
// remainder |rho| <= 1/128: e^{2 pi i c} = T[k] e^{2 pi i rho}, k = rint(64 c) mod 64, rho = c - k/64 (exact).  Four FP64
// operations and all of the quadrant selects fewer than the quarter-turn reduction; the W arguments advance together.
template <int W>
__device__ __forceinline__ void sincos_cycles_tab(const double (&c)[W], const double2 *__restrict__ rot, double (&sn)[W], double (&cs)[W]) {
    double rho[W], x2[W], ps[W], pc[W];
    double2 T[W];
#pragma unroll
    for (int i = 0; i < W; i++) {
        const double kq = rint(64.0 * c[i]);
        rho[i] = fma(-0.015625, kq, c[i]);
        T[i] = rot[(int)kq & 63];
        x2[i] = rho[i] * rho[i];
        ps[i] = c_sin64[3]; pc[i] = c_cos64[3];
    }
#pragma unroll
    for (int k = 2; k >= 0; k--) {
#pragma unroll
        for (int i = 0; i < W; i++) { ps[i] = fma(ps[i], x2[i], c_sin64[k]); pc[i] = fma(pc[i], x2[i], c_cos64[k]); }
    }
#pragma unroll
    for (int i = 0; i < W; i++) {
        ps[i] *= rho[i];                 // sin(2 pi rho)
        pc[i] = fma(pc[i], x2[i], 1.0);  // cos(2 pi rho)
        cs[i] = fma(T[i].x, pc[i], -T[i].y * ps[i]);
        sn[i] = fma(T[i].x, ps[i], T[i].y * pc[i]);
    }
}

// robust bracketed Newton (rare path: cold-start failures, turnover neighbourhood)
__device__ __noinline__ double solve_bracketed(double c1, double c2, double c3, double delta, double xl, double xh,
                                               double sdir, double hj) {
    double gl = xl * fma(xl, fma(xl, c3, c2), c1) - delta;
    double gh = xh * fma(xh, fma(xh, c3, c2), c1) - delta;
    double x = (gh == gl) ? 0.5 * (xl + xh) : xl - gl * (xh - xl) / (gh - gl);
    x = fmin(fmax(x, xl), xh);
    const double tol = 1e-9 * hj;
    for (int it = 0; it < 80; it++) {
        const double gx = x * fma(x, fma(x, c3, c2), c1) - delta;
        const double dg = fma(x, fma(3.0 * c3, x, 2.0 * c2), c1);
        if (gx * sdir > 0.0) xh = x; else xl = x;
        double xn = x - gx / dg;
        if (!(xn >= xl && xn <= xh)) xn = 0.5 * (xl + xh);
        const double dx = fabs(xn - x);
        x = xn;
        if (dx <= tol) break;
    }
    return x;
}

// out-of-line slow path of the root solve: first bin of a thread, segment change, or a fast step that did not converge
__device__ __noinline__ double solve_slow(double c1, double c2, double c3, double delta, double xlo, double xhi, double tol,
                                          double sdir) {
    const double gl = xlo * fma(xlo, fma(xlo, c3, c2), c1) - delta;
    const double gh = xhi * fma(xhi, fma(xhi, c3, c2), c1) - delta;
    double x = xlo - gl * (xhi - xlo) * fast_rcp(gh - gl);
    for (int it = 0; it < 6; it++) {
        const double gx = x * fma(x, fma(x, c3, c2), c1) - delta;
        const double dx = gx * fast_rcp(fma(x, fma(3.0 * c3, x, 2.0 * c2), c1));
        x -= dx;
        if (fabs(dx) <= tol) {
            if (x >= xlo && x <= xhi) return x;
            break;
        }
    }
    return solve_bracketed(c1, c2, c3, delta, xlo, xhi, sdir, tol * 1e6);
}

// ==========================================================================================
// (m, n) groups.  The stationary point t*, the SPA factor and the phase of a harmonic depend on (m, n) only
// (Tutorial_FD_construction_single_mode.ipynb:558-616, cell 26): modes (l, m, n) that share (m, n) differ in A_lmn(t) Y_lm
// alone, and the cubic-spline quads are linear in the knot values.  group_kernel finds each walker's distinct (m, n) pairs
// ("groups", numbered in order of first occurrence) and combines the amplitude quads of a group's members into two complex
// rows per (knot, group),
//     Cp = e^{+i 3pi/4} sum_l Y_lm A_lmn,        Cm = e^{-i 3pi/4} sum_l Y_l-m conj(A_lmn),
// so that mode_sum_kernel solves ONE stationary point per (group, bin) instead of one per (mode, bin): 30 modes -> 27 groups
// at eps = 1e-2, 3843 -> 671 with all modes.  e^{+-i 3pi/4} is the constant phase of the SPA factor G on rising branches;
// falling branches add a quarter turn to the phase instead (SubEntry::mu_hi).  The per-mode work-list of segment_kernel stays
// the exported A4 result; a group uses the records of its first member.
// ==========================================================================================
#define GRP_THREADS 256
#define GRP_TAB 8192 /* cells of the direct (m, n) -> first mode table; larger index ranges take the quadratic fallback */
struct GroupParams {
    const emrifd_walker_t *w;
    const double *coeff;
    const int *m, *n;
    const double2 *ylm;
    int *leader; // [sum K]: walker block at mode_off, its first G entries = mode index of each group's first member
    int *gcount; // [B] number of groups G
    int *gmem;   // [sum K]: walker block at mode_off = member modes, grouped (ascending mode index inside a group)
    int *goff;   // [sum K + B]: walker w's block at mode_off + w = G + 1 offsets into its gmem block
    double *gq;  // walker block at 16 * teuk_off doubles: [L][G][16] = quads of Re Cp, Im Cp, Re Cm, Im Cm
};

// Step 1, one CTA per walker: distinct (m, n) pairs numbered in order of first occurrence, and the member list of every group
// in ascending mode order (a fixed summation order for the combination).
__global__ void __launch_bounds__(GRP_THREADS) group_index_kernel(GroupParams p) {
    extern __shared__ int gsm[]; // tab [GRP_TAB] | grp [K] | first [K] | cur [K]
    __shared__ int s_mn[4], s_warp[GRP_THREADS / 32], s_base, s_g[GRP_THREADS];
    const emrifd_walker_t wd = p.w[blockIdx.x];
    const int K = wd.K;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int *marr = p.m + wd.mode_off, *narr = p.n + wd.mode_off;
    int *tab = gsm, *grp = gsm + GRP_TAB, *first = grp + K, *cur = first + K;
    int *lead = p.leader + wd.mode_off, *mem = p.gmem + wd.mode_off, *off = p.goff + wd.mode_off + blockIdx.x;
    if (tid == 0) { s_mn[0] = INT_MAX; s_mn[1] = INT_MIN; s_mn[2] = INT_MAX; s_mn[3] = INT_MIN; s_base = 0; }
    __syncthreads();
    {
        int mlo = INT_MAX, mhi = INT_MIN, nlo = INT_MAX, nhi = INT_MIN;
        for (int k = tid; k < K; k += GRP_THREADS) {
            const int mk = marr[k], nk = narr[k];
            mlo = min(mlo, mk); mhi = max(mhi, mk); nlo = min(nlo, nk); nhi = max(nhi, nk);
        }
        mlo = __reduce_min_sync(0xffffffffu, mlo); mhi = __reduce_max_sync(0xffffffffu, mhi);
        nlo = __reduce_min_sync(0xffffffffu, nlo); nhi = __reduce_max_sync(0xffffffffu, nhi);
        if (lane == 0) { atomicMin(&s_mn[0], mlo); atomicMax(&s_mn[1], mhi); atomicMin(&s_mn[2], nlo); atomicMax(&s_mn[3], nhi); }
    }
    __syncthreads();
    const int mmin = s_mn[0], nmin = s_mn[2];
    const long long mspan = (long long)s_mn[1] - mmin + 1, nspan = (long long)s_mn[3] - nmin + 1;
    // ---- first member of every mode's group -> first[k] ----
    if (mspan * nspan <= GRP_TAB) {
        const int cells = (int)(mspan * nspan), ns = (int)nspan;
        for (int i = tid; i < cells; i += GRP_THREADS) tab[i] = INT_MAX;
        __syncthreads();
        for (int k = tid; k < K; k += GRP_THREADS) atomicMin(&tab[(marr[k] - mmin) * ns + (narr[k] - nmin)], k);
        __syncthreads();
        for (int k = tid; k < K; k += GRP_THREADS) first[k] = tab[(marr[k] - mmin) * ns + (narr[k] - nmin)];
    } else {
        for (int k = tid; k < K; k += GRP_THREADS) {
            const int mk = marr[k], nk = narr[k];
            int j = 0;
            while (j < k && !(marr[j] == mk && narr[j] == nk)) j++;
            first[k] = j;
        }
    }
    __syncthreads();
    // ---- number the groups in order of first occurrence ----
    for (int k0 = 0; k0 < K; k0 += GRP_THREADS) {
        const int k = k0 + tid;
        const bool f = k < K && first[k] == k;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int o = s_base;
        for (int q = 0; q < wid; q++) o += s_warp[q];
        if (f) { const int g = o + __popc(bal & ((1u << lane) - 1u)); grp[k] = g; lead[g] = k; }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int q = 0; q < GRP_THREADS / 32; q++) t += s_warp[q]; s_base += t; }
        __syncthreads();
    }
    const int G = s_base;
    for (int k = tid; k < K; k += GRP_THREADS) if (first[k] != k) grp[k] = grp[first[k]];
    for (int g = tid; g < G; g += GRP_THREADS) cur[g] = 0;
    __syncthreads();
    // ---- group sizes -> offsets (cur is the count first, then the fill cursor) ----
    for (int k = tid; k < K; k += GRP_THREADS) atomicAdd(&cur[grp[k]], 1);
    __syncthreads();
    if (wid == 0) {
        int carry = 0;
        for (int g0 = 0; g0 < G; g0 += 32) {
            const int g = g0 + lane;
            const int c = g < G ? cur[g] : 0;
            int v = c;
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            if (g < G) { off[g] = carry + v - c; cur[g] = 0; }
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) { off[G] = carry; p.gcount[blockIdx.x] = G; }
    }
    __syncthreads();
    // ---- stable fill, a chunk of GRP_THREADS modes at a time: slot = start + members of earlier chunks + earlier threads of
    //      this chunk with the same group ----
    for (int k0 = 0; k0 < K; k0 += GRP_THREADS) {
        const int k = k0 + tid, g = k < K ? grp[k] : -1;
        s_g[tid] = g;
        __syncthreads();
        if (g >= 0) {
            int rank = 0;
            for (int t = 0; t < tid; t++) rank += (s_g[t] == g);
            mem[off[g] + cur[g] + rank] = k;
        }
        __syncthreads();
        if (g >= 0) atomicAdd(&cur[g], 1);
        __syncthreads();
    }
}

// Step 2, grid (knot slices, walkers): the combined quads of a slice of knots.
__global__ void __launch_bounds__(GRP_THREADS) group_combine_kernel(GroupParams p) {
    const emrifd_walker_t wd = p.w[blockIdx.y];
    const int K = wd.K, L = wd.L, R = 2 * K + 4, tid = threadIdx.x;
    const int G = p.gcount[blockIdx.y];
    const int *mem = p.gmem + wd.mode_off, *off = p.goff + wd.mode_off + blockIdx.y;
    const double2 *ylm = p.ylm + 2 * wd.mode_off;
    const double *coeff = p.coeff + wd.coeff_off;
    double *gq = p.gq + 16 * wd.teuk_off;
    const double r2 = 0.7071067811865476;
    for (int item = blockIdx.x * GRP_THREADS + tid; item < L * G; item += gridDim.x * GRP_THREADS) {
        const int j = item / G, g = item - j * G;
        double pr[4] = {0, 0, 0, 0}, pi[4] = {0, 0, 0, 0}, mr[4] = {0, 0, 0, 0}, mi[4] = {0, 0, 0, 0};
        for (int i = off[g]; i < off[g + 1]; i++) {
            const int k = mem[i];
            const double4 a4 = *reinterpret_cast<const double4 *>(coeff + ((long long)j * R + k) * 4);
            const double4 b4 = *reinterpret_cast<const double4 *>(coeff + ((long long)j * R + K + k) * 4);
            const double2 yp = ylm[k], ym = ylm[K + k];
            // Y_lm e^{+i 3pi/4} and Y_l-m e^{-i 3pi/4}
            const double ypr = -r2 * (yp.x + yp.y), ypi = r2 * (yp.x - yp.y);
            const double ymr = r2 * (ym.y - ym.x), ymi = -r2 * (ym.x + ym.y);
            const double a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                pr[c] = fma(ypr, a[c], fma(-ypi, b[c], pr[c])); // Y (a + i b)
                pi[c] = fma(ypr, b[c], fma(ypi, a[c], pi[c]));
                mr[c] = fma(ymr, a[c], fma(ymi, b[c], mr[c]));  // Y_- (a - i b)
                mi[c] = fma(ymi, a[c], fma(-ymr, b[c], mi[c]));
            }
        }
        double4 *o = reinterpret_cast<double4 *>(gq + (long long)item * 16);
        o[0] = make_double4(pr[0], pr[1], pr[2], pr[3]);
        o[1] = make_double4(pi[0], pi[1], pi[2], pi[3]);
        o[2] = make_double4(mr[0], mr[1], mr[2], mr[3]);
        o[3] = make_double4(mi[0], mi[1], mi[2], mi[3]);
    }
}

// ==========================================================================================
// mode-sum building blocks
// ==========================================================================================
// One monotone cubic piece of one group's f_mn(t) restricted to the current tile: a (branch, side, spline segment) triple with
// every per-segment constant combined once per tile by the fill warp, so the evaluation loop neither searches segments nor
// recombines (m, n) with the track quads.
#define SE_SIDE 1   /* bins are at -f: the direct term lands in the -f accumulators, the mirrored -m term in the +f ones */
#define SE_MIRROR 2 /* m > 0 and include_minus_m: add the mirrored term */
#define SE_FALL 4   /* falling branch: G is conjugated (and mu_hi carries the extra quarter turn) */
// One monotone cubic piece of one group's f_mn(t): a (branch, spline segment) pair with every per-segment constant combined
// once per batch by piece_build_kernel -- the evaluation loop neither searches segments nor recombines (m, n) with the track
// quads, and gets its root from the piece's inverse interpolant x(f) plus one Newton step.
#define PIECE_DEG 5
struct __align__(16) Piece {
    double c0, c1, c2, c3;       // f_mn(t_j + x) = c0 + c1 x + c2 x^2 + c3 x^3
    double d2, d3, fmid, finv;   // 2 c2, 3 c3; interpolant variable u = (f - fmid) finv in [-1, 1] over the piece
    double q0, q1, q2, q3;       // x(u) ~ q0 + u (q1 + u (... + u q5)): degree-5 interpolant through the Chebyshev points of the piece
    double q4, q5, tol, tj;      // tol: largest Newton step after which the root is good to 1e-7 s (< 0: no usable interpolant,
                                 // every root goes through the bracketed solver); knot time
    double xlo, xhi, mu_hi, mu_lo; // slackened root bracket; (m Phi_phi + n Phi_r)(t_j) / 2pi mod 1 as a double-double (-1/4 on falling branches)
    double p1, p2, p3, pad;      // -(1/2pi) x (m Phi_phi + n Phi_r) cubic remainder, in cycles
    double4 amp[4];              // quads of Re Cp, Im Cp, Re Cm, Im Cm of (segment, group)
};
// A piece restricted to the current tile (one entry of a pass): the 16-byte header is written by the producer warp, the body is
// copied from the piece table by the TMA (cp.async.bulk) straight into the ring slot.
struct __align__(16) SubEntry {
    int s, e;           // tile-local bin range (inclusive)
    unsigned int fmask; // 0 / 0x80000000: sign mask applied to the bin frequency
    int flags;
    Piece P;
};
// overlapping work-list record of the current fill round (written and read by the producer warp only)
struct __align__(16) FillEntry {
    double dm, dn;
    int ja, jb, dir, r; // r: group record (group * MAXBR + branch) -> row of the piece table
    int s[2], e[2];     // tile-local bin range per side (+f, -f); empty if s > e
    int jlo[2], jhi[2]; // spline segments those bins fall in
    int mirror, off;    // off: index of its first sub-entry in the round's list
    int nsub, pad;
};
// group record as the producer warp caches it per walker
struct __align__(16) RecC {
    long long start, end;
    int ja, jb, dir, mirror;
    double dm, dn;
    int lo, hi;         // hull of the positive bins it touches (through +f or -f)
    int pad0, pad1;
};
#ifndef SUM_SUBCAP
#define SUM_SUBCAP 32 /* sub-entries per pass */
#endif
#define SUM_ECAP 32   /* work-list records per fill round: one per lane of the producer warp */

__device__ __noinline__ double2 spa_fix(double fdot, double fddot, double s, double u, int few) {
    // rare path of the SPA factor: X = 1/u < 1024 (late inspiral, turnover neighbourhood); returns R/sqrt|fdot|.
    // Out of line and returning by value: the hot loop keeps no address-taken locals and no code of this path.
    // few != 0: FastEMRIWaveforms-compatible evaluation: 9-term asymptotic series for X > 7, 14-term
    // ascending series below (2.5e-7 off at the seam); default: <= 3e-15 everywhere.
    double re, im;
    if (few) {
        if (u < 1.0 / 7.0) {
            const double w = u * u;
            double pr = k13_asym_re[4], pi = k13_asym_im[3];
#pragma unroll
            for (int k = 3; k >= 0; k--) pr = fma(pr, w, k13_asym_re[k]);
#pragma unroll
            for (int k = 2; k >= 0; k--) pi = fma(pi, w, k13_asym_im[k]);
            re = pr * s; im = u * pi * s;
        } else {
            const double af = fabs(fdot);
            const double X = 2.0943951023931953 * af * af * af / (fddot * fddot);
            k13_small_S(X, re, im);
            const double sc = cbrt(1.4472025091165353 / fabs(fddot));
            re *= sc; im *= sc;
        }
        return make_double2(re, im);
    }
    if (u <= 0.03125) {
        const double w = u * u;
        double pr = k13_asym_re[6], pi = k13_asym_im[6];
#pragma unroll
        for (int k = 5; k >= 0; k--) { pr = fma(pr, w, k13_asym_re[k]); pi = fma(pi, w, k13_asym_im[k]); }
        re = pr * s; im = u * pi * s;
    } else if (u <= 1.0) {
        k13_mid(1.0 / u, re, im);
        re *= s; im *= s;
    } else {
        const double af = fabs(fdot);
        const double X = 2.0943951023931953 * af * af * af / (fddot * fddot);
        k13_small_S(X, re, im);
        const double sc = cbrt(1.4472025091165353 / fabs(fddot));
        re *= sc; im *= sc;
    }
    return make_double2(re, im);
}

// ---- evaluation of the W bins of one sub-entry in straight-line code: the W independent dependency chains (SPA factor,
//      phase, sincos, amplitude Horner) interleave in the instruction stream (ILP without more warps).  The accumulators
//      are registers: ad_* receive the direct term, ao_* the mirrored -m term; a bin with in[i] == false is computed (as a
//      copy of its neighbour) but not accumulated ----
template <int W>
__device__ __forceinline__ void eval_sub(const double (&x)[W], const double (&f)[W], const bool (&in)[W], const Piece &S,
                                         const int fl, const int few, const double2 *__restrict__ rot, double (&ad_r)[W],
                                         double (&ad_i)[W], double (&ao_r)[W], double (&ao_i)[W]) {
    double er[W], ei[W];
    {
        double re[W], im[W], s[W], uu[W], fd[W], fdd[W], sn[W], cs[W], cyc[W];
        const double c1 = S.c1, d2 = S.d2, d3 = S.d3;
        const double tj = S.tj, mu_hi = S.mu_hi, mu_lo = S.mu_lo, p1 = S.p1, p2 = S.p2, p3 = S.p3;
#pragma unroll
        for (int i = 0; i < W; i++) {
            const double xi = x[i], fi = f[i];
            fd[i] = fma(xi, fma(d3, xi, d2), c1);
            fdd[i] = fma(2.0 * d3, xi, d2);
            // SPA factor, common path (X >= 1024): s = 1/sqrt|fdot|, u = 1/X = 3 fddot^2 s^6/(2 pi)
            s[i] = fast_rsqrt(fabs(fd[i]));
            const double s2 = s[i] * s[i];
            const double tq = (0.6909882989426709 * fdd[i]) * s2;         // sqrt(3 / 2 pi) fddot s^2
            uu[i] = (tq * tq) * s2;
            const double w = uu[i] * uu[i];
            re[i] = fma(w, fma(w, k13_asym_re[2], k13_asym_re[1]), 1.0) * s[i];
            im[i] = uu[i] * fma(w, fma(w, k13_asym_im[2], k13_asym_im[1]), k13_asym_im[0]) * s[i];
            // phase in cycles: f t_j as an exact two-product reduced mod 1, the knot phase as a double-double,
            // only the small polynomial remainder in plain double
            double p0 = fi * tj;
            const double e0 = fma(fi, tj, -p0);
            p0 -= rint_fast(p0);
            const double poly = fma(fi, xi, xi * fma(xi, fma(xi, p3, p2), p1));
            cyc[i] = ((p0 - mu_hi) + (e0 - mu_lo)) + poly;
        }
        sincos_cycles_tab<W>(cyc, rot, sn, cs);
        bool slow = false; // one branch for all W bins: the rare path is taken by the whole group
#pragma unroll
        for (int i = 0; i < W; i++) slow |= !(uu[i] <= 0.0009765625);
        if (slow) {
#pragma unroll
            for (int i = 0; i < W; i++)
                if (!(uu[i] <= 0.0009765625)) { const double2 g = spa_fix(fd[i], fdd[i], s[i], uu[i], few); re[i] = g.x; im[i] = g.y; }
        }
        const unsigned int cmask = (fl & SE_FALL) ? 0x80000000u : 0u;
#pragma unroll
        for (int i = 0; i < W; i++) { // E = R~ / sqrt|fdot| e^{i phase}  (R conjugated on falling branches)
            const double gim = flip_sign(im[i], cmask);
            er[i] = re[i] * cs[i] - gim * sn[i];
            ei[i] = re[i] * sn[i] + gim * cs[i];
        }
    }
    const double4 *ap = S.amp;
    {
        const double4 qa = ap[0], qb = ap[1];
#pragma unroll
        for (int i = 0; i < W; i++) { // W(+-f) += Cp E
            const double xi = x[i];
            const double cr = fma(xi, fma(xi, fma(xi, qa.w, qa.z), qa.y), qa.x);
            const double ci = fma(xi, fma(xi, fma(xi, qb.w, qb.z), qb.y), qb.x);
            if (in[i]) {
                ad_r[i] = fma(cr, er[i], fma(-ci, ei[i], ad_r[i]));
                ad_i[i] = fma(cr, ei[i], fma(ci, er[i], ad_i[i]));
            }
        }
    }
    if (fl & SE_MIRROR) {
        const double4 qc = ap[2], qd = ap[3];
#pragma unroll
        for (int i = 0; i < W; i++) { // W(-+f) += Cm conj(E)
            const double xi = x[i];
            const double cr = fma(xi, fma(xi, fma(xi, qc.w, qc.z), qc.y), qc.x);
            const double ci = fma(xi, fma(xi, fma(xi, qd.w, qd.z), qd.y), qd.x);
            if (in[i]) {
                ao_r[i] = fma(cr, er[i], fma(ci, ei[i], ao_r[i]));
                ao_i[i] = fma(ci, er[i], fma(-cr, ei[i], ao_i[i]));
            }
        }
    }
}

// Hull of the positive-bin indices touched by each chunk of SUM_CHUNK group records (either through the +f or the -f side):
// lets the mode sum skip a whole chunk (no ballot, no barrier pair) when its tile is outside, and empty_tile_kernel classify tiles.
__global__ void __launch_bounds__(SUM_CHUNK) chunk_range_kernel(const emrifd_walker_t *w, const emrifd_branch_t *brs,
                                                                   const int *leader, const int *gcount, long long zero,
                                                                   long long *rng, int cpw) {
    __shared__ long long s_lo[SUM_CHUNK / 32], s_hi[SUM_CHUNK / 32];
    const emrifd_walker_t wd = w[blockIdx.y];
    const int nrec = gcount[blockIdx.y] * MAXBR, r = blockIdx.x * SUM_CHUNK + threadIdx.x;
    long long lo = 0x7fffffffffffffffLL, hi = -1;
    if (r < nrec) {
        const emrifd_branch_t *b = brs + (wd.mode_off + leader[wd.mode_off + r / MAXBR]) * MAXBR + (r % MAXBR);
        const long long s0 = b->start, e0 = b->end;
        if (e0 >= s0) {
            if (e0 >= zero) { lo = (s0 > zero ? s0 : zero) - zero; hi = e0 - zero; }
            if (s0 <= zero) {
                const long long a = zero - (e0 < zero ? e0 : zero), c = zero - s0;
                lo = a < lo ? a : lo; hi = c > hi ? c : hi;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const long long l2 = __shfl_down_sync(0xffffffffu, lo, o), h2 = __shfl_down_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < SUM_CHUNK / 32; q++) { lo = s_lo[q] < lo ? s_lo[q] : lo; hi = s_hi[q] > hi ? s_hi[q] : hi; }
        long long *o = rng + ((long long)blockIdx.y * cpw + blockIdx.x) * 2;
        o[0] = lo; o[1] = hi;
    }
}

// sum over both channels of |d~|^2 for every tile of SUM_TILE bins of the injected data: a tile no harmonic touches
// contributes exactly this to sum |d~ - h~|^2, so mode_sum_kernel does not have to read the data there
__global__ void __launch_bounds__(256) tile_dd_kernel(const double2 *__restrict__ dw, long long n, int tile, double *__restrict__ out) {
    __shared__ double s[8];
    const long long j0 = (long long)blockIdx.x * tile;
    double a = 0.0;
    for (int i = threadIdx.x; i < tile; i += 256) {
        const long long j = j0 + i;
        if (j < n) { const double2 d0 = dw[j], d1 = dw[n + j]; a += d0.x * d0.x + d0.y * d0.y + d1.x * d1.x + d1.y * d1.y; }
    }
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int q = 0; q < 8; q++) t += s[q]; out[blockIdx.x] = t; }
}

// A tile that a bin slice cuts short (a slice end that is neither tile-aligned nor the end of the data) must not take the
// precomputed whole-tile sum |d~|^2: its bins beyond the slice end belong to another rank.  Such a tile goes through the
// per-bin path, which stops at the slice end.
__device__ __forceinline__ bool tile_truncated(const SumParams &p, long long jt0, int tile_bins) {
    const long long jend = p.j_lo + p.j_cnt;
    return jt0 + tile_bins > jend && jend != p.n_data;
}

// Does anything of walker `wy`'s work-list touch tile `tx`?  (chunk hulls; a failed walker has no work.)  `per_bin`: the tile's
// likelihood term cannot come from the precomputed whole-tile sum |d~|^2 (no table for this slice, or the slice cuts the tile).
template <bool LIKE, int BPT>
__device__ __forceinline__ bool tile_has_work(const SumParams &p, int tx, int wy, long long &jt0, long long &jt1) {
    constexpr int TILE = SUM_CT * BPT;
    const long long *crng = p.chunk_rng + (long long)wy * p.cpw * 2;
    const int nrec = p.gcount[wy] * MAXBR;
    jt0 = p.j_lo + (p.tile_first + (long long)tx * p.tile_stride) * TILE;
    const long long jend = p.j_lo + p.j_cnt;
    jt1 = (jt0 + TILE < jend ? jt0 + TILE : jend) - 1;
    bool any = false;
    for (int ch = 0; ch * SUM_CHUNK < nrec; ch++) any |= !(crng[2 * ch] > jt1 || crng[2 * ch + 1] < jt0);
    if (p.wstatus[wy]) any = false;
    const bool per_bin = LIKE && (p.no_empty || tile_truncated(p, jt0, TILE));
    return any || per_bin;
}

// Pass 1a: one thread per (tile, walker) flags the tiles that have work for mode_sum_kernel; queue_build_kernel then compacts
// the flagged tiles into the work queue IN ORDER (walker-major, ascending tile: a walker's tiles stay together for the producer
// warp's per-walker cache, the order is the same from run to run, and the long low-frequency tiles of a many-mode waveform come
// first instead of landing in the tail: atomically appended queues cost configs[3] 7 %).  A few microseconds each.
#define CLASSIFY_THREADS 256
template <bool LIKE, int BPT>
__global__ void __launch_bounds__(CLASSIFY_THREADS) classify_tiles_kernel(SumParams p) {
    const int tx = blockIdx.x * CLASSIFY_THREADS + threadIdx.x, wy = blockIdx.y;
    if (tx >= p.ntiles) return;
    long long jt0, jt1;
    p.tile_flag[(long long)wy * p.ntiles + tx] = tile_has_work<LIKE, BPT>(p, tx, wy, jt0, jt1) ? 1 : 0;
}
#define QBUILD_THREADS 1024
__global__ void __launch_bounds__(QBUILD_THREADS) queue_build_kernel(const unsigned char *__restrict__ flag, long long ntot, int ntiles,
                                                                     unsigned long long *__restrict__ queue, unsigned int *__restrict__ qctl) {
    __shared__ unsigned int s_warp[QBUILD_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // a thread owns a run of flags that is a multiple of 8 long and reads them eight at a time (the flag array is 8-byte aligned
    // and allocated up to the next multiple of 8; the bytes beyond ntot are masked here)
    long long per = (ntot + QBUILD_THREADS - 1) / QBUILD_THREADS;
    per = (per + 7) & ~7LL;
    const long long lo = (long long)tid * per, hi = lo + per < ntot ? lo + per : ntot;
    const unsigned long long *f8 = reinterpret_cast<const unsigned long long *>(flag);
    unsigned int cnt = 0;
    for (long long t = lo; t < hi; t += 8) {
        unsigned long long w = f8[t >> 3];
        if (t + 8 > ntot) w &= (1ull << (8 * (ntot - t))) - 1ull; // bytes past the end
        cnt += __popcll(w & 0x0101010101010101ull);
    }
    unsigned int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned int v = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        s_warp[lane] = v; // inclusive totals of the warps
    }
    __syncthreads();
    unsigned int off = incl - cnt + (wid > 0 ? s_warp[wid - 1] : 0u);
    if (cnt) {
        unsigned int wy = (unsigned int)(lo / ntiles), tx = (unsigned int)(lo - (long long)wy * ntiles);
        for (long long t = lo; t < hi; t += 8) {
            unsigned long long w = f8[t >> 3];
            if (t + 8 > ntot) w &= (1ull << (8 * (ntot - t))) - 1ull;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if ((w >> (8 * c)) & 1ull) queue[off++] = ((unsigned long long)wy << 32) | tx;
                if (++tx == (unsigned int)ntiles) { tx = 0; wy++; }
            }
        }
    }
    if (tid == QBUILD_THREADS - 1) qctl[0] = s_warp[QBUILD_THREADS / 32 - 1];
}

// Pass 1b: tiles no harmonic touches (most of the band of a non-plunging eps = 1e-2 system): h = 0 is stored and the tile's
// likelihood term is the precomputed sum |d~|^2.  A kernel of its own because this work is a pure store stream, and on a stream
// of its own: one 128-thread, 32-register CTA per tile slots into the 4 K registers per SM that mode_sum_kernel's CTA leaves
// free, so the zeros go out underneath the FP64-bound sum (the bench batch's 0.27 ms store stream costs the sum 0.12 ms), and
// 16 CTAs per SM take the whole machine once the sum is done (the sparse regime: 7 TB/s).  A walker whose status word is set
// (bad knots, too many branches) has all its tiles treated as empty: zeros in h, NaN in the likelihood (like_finalize_kernel).
// Measured alternatives: persistent zero-fill CTAs with dynamic tile hand-out stream 30 % slower on their own (4.9 TB/s) and,
// by keeping the memory system busy during the whole sparse sum, make that latency-bound sum 2.3x slower.
#define EMPTY_THREADS 128
template <bool WRITE_H, bool LIKE, int BPT>
__global__ void __launch_bounds__(EMPTY_THREADS, 16) empty_tile_kernel(SumParams p) {
    constexpr int TILE = SUM_CT * BPT;
    const int tid = threadIdx.x;
    if (p.tile_flag[(long long)blockIdx.y * p.ntiles + blockIdx.x]) return; // mode_sum_kernel's tile
    const long long out_off = p.w[blockIdx.y].out_off;
    const long long tglob = p.tile_first + (long long)blockIdx.x * p.tile_stride; // tile index within the slice
    const long long jt0 = p.j_lo + tglob * TILE;
    const long long jend = p.j_lo + p.j_cnt;
    const long long jt1 = (jt0 + TILE < jend ? jt0 + TILE : jend) - 1;
    const int ntile_ = (int)(jt1 - jt0 + 1);
    if (WRITE_H) {
        const double2 z = make_double2(0.0, 0.0);
        const long long zero = p.g.zero;
#pragma unroll 4
        for (int lb = tid; lb < ntile_; lb += EMPTY_THREADS) {
            const long long j = jt0 + lb;
            if (p.mask_positive) {
                const long long o = out_off + (j - p.j_lo);
                p.hp[o] = z; p.hc[o] = z;
            } else {
                const long long o = out_off + zero;
                p.hp[o + j] = z; p.hc[o + j] = z;
                if (j > 0) { p.hp[o - j] = z; p.hc[o - j] = z; }
            }
        }
    }
    if (LIKE && tid < SUM_CW) { // the partial-sum slots of the tile's consumer warps
        double *o = p.partial + (((long long)blockIdx.y * p.ntiles + blockIdx.x) * SUM_CW + tid) * 3;
        o[0] = (tid == 0) ? p.tile_dd[p.j_lo / TILE + tglob] : 0.0; o[1] = 0.0; o[2] = 0.0; // (tile_dd is used only when j_lo is tile-aligned)
    }
}

// exact frequency of tile-local bin lb, as every kernel of the path computes it
__device__ __forceinline__ double tile_binf(const Grid &g, long long jt0, int lb) {
    const long long jj = jt0 + lb;
    return g.fpos ? g.fpos[jj] : rmul((double)(int)jj, g.val);
}

// knot frequency of a group exactly as segment_kernel rounds it
__device__ __forceinline__ double knot_F(const double *sK, int j, double dm, double dn) {
    return radd(rmul(dm, sK[3 * j + 1]), rmul(dn, sK[3 * j + 2]));
}

// ==========================================================================================
// Piece table: one Piece per (group record, spline segment) of every walker, built once per batch.  Row r = group * MAXBR +
// branch of walker w starts at ((mode_off * MAXBR + r) * lstride); entry j of a row is the piece on segment j (valid for
// ja <= j <= jb of a non-empty branch).  The inverse interpolant is fitted over the piece's whole x-range (the segment, cut
// at the branch ends) and VERIFIED here; where it is not good enough (turnover neighbourhoods: 0.3 % of the evaluations of
// the bench workload) tol < 0 sends the evaluation through the bracketed solver.
// ==========================================================================================
struct PieceParams {
    const emrifd_walker_t *w;
    const double *t, *coeff;
    const int *m, *n;
    const emrifd_branch_t *br;
    const int *leader, *gcount;
    const double *gq;
    Piece *pieces;
    int lstride;
};
#define PIECE_THREADS 128
__global__ void __launch_bounds__(PIECE_THREADS) piece_build_kernel(PieceParams p) {
    const emrifd_walker_t wd = p.w[blockIdx.y];
    const int G = p.gcount[blockIdx.y], L = wd.L, K = wd.K, R = 2 * K + 4;
    const int nseg = L - 1;
    const long long tot = (long long)G * nseg; // (group, segment) pairs; one thread per (branch slot, group, segment)
    const double *coeff = p.coeff + wd.coeff_off;
    const double *gq = p.gq + 16 * wd.teuk_off;
    const double *tk = p.t + wd.knot_off;
    for (long long idx = (long long)blockIdx.x * PIECE_THREADS + threadIdx.x; idx < tot * MAXBR; idx += (long long)gridDim.x * PIECE_THREADS) {
        // (group, segment) pairs first, branch slot last: the threads of a warp share the branch slot, and slot 0 -- the only
        // non-empty one of a monotone harmonic -- keeps whole warps busy while the warps of the empty slots leave at once
        const int bslot = (int)(idx / tot);
        const long long gj = idx - (long long)bslot * tot;
        const int gi = (int)(gj / nseg), j = (int)(gj - (long long)gi * nseg);
        const int r = gi * MAXBR + bslot;
        const int lead = p.leader[wd.mode_off + gi];
        const emrifd_branch_t *b = p.br + (wd.mode_off + lead) * MAXBR + bslot;
        if (b->end < b->start) continue;
        const int ja = b->ja, jb = b->jb, dir = b->dir;
        if (j < ja || j > jb) continue;
        const double dm = (double)p.m[wd.mode_off + lead], dn = (double)p.n[wd.mode_off + lead];
        const double4 qf = *reinterpret_cast<const double4 *>(coeff + ((long long)j * R + 2 * K) * 4);     // f_phi quad
        const double4 qr = *reinterpret_cast<const double4 *>(coeff + ((long long)j * R + 2 * K + 1) * 4); // f_r
        const double4 qP = *reinterpret_cast<const double4 *>(coeff + ((long long)j * R + 2 * K + 2) * 4); // Phi_phi
        const double4 qR = *reinterpret_cast<const double4 *>(coeff + ((long long)j * R + 2 * K + 3) * 4); // Phi_r
        const double4 *ga = reinterpret_cast<const double4 *>(gq + ((long long)j * G + gi) * 16);
        Piece S;
        S.amp[0] = ga[0]; S.amp[1] = ga[1]; S.amp[2] = ga[2]; S.amp[3] = ga[3];
        const double tj = tk[j], hj = tk[j + 1] - tj;
        S.c0 = radd(rmul(dm, qf.x), rmul(dn, qr.x));
        S.c1 = fma(dm, qf.y, dn * qr.y);
        S.c2 = fma(dm, qf.z, dn * qr.z);
        S.c3 = fma(dm, qf.w, dn * qr.w);
        S.d2 = 2.0 * S.c2; S.d3 = 3.0 * S.c3;
        const double xl0 = (j == ja) ? b->xa : 0.0, xh0 = (j == jb) ? b->xb : hj;
        S.xlo = xl0 - 1e-5 * hj; S.xhi = xh0 + 1e-5 * hj;
        S.tj = tj;
        { // ---- inverse interpolant x(f) over [xl0, xh0]: the consumers' root = interpolant + ONE Newton step ----
            const double sdir = dir > 0 ? 1.0 : -1.0;
            const double c0 = S.c0, c1 = S.c1, c2 = S.c2, c3 = S.c3;
            const double xm = 0.5 * (xl0 + xh0), xr = 0.5 * (xh0 - xl0);
            const double fa = c0 + xl0 * fma(xl0, fma(xl0, c3, c2), c1), fb = c0 + xh0 * fma(xh0, fma(xh0, c3, c2), c1);
            S.fmid = 0.5 * (fa + fb);
            const double fh = 0.5 * (fb - fa);
            S.finv = fh != 0.0 ? 1.0 / fh : 0.0;
            if (!(fabs(S.finv) < 1e300)) S.finv = 0.0;
            // curvature bound K = max |fddot / (2 fdot)| over the piece: one Newton step of size d leaves an error ~ K d^2
            double Kc = 0.0;
            bool bad = !(xr > 0.0) || S.finv == 0.0;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const double xq = q == 0 ? xl0 : (q == 1 ? xh0 : xm);
                const double fd = fma(xq, fma(S.d3, xq, S.d2), c1), fdd = fma(2.0 * S.d3, xq, S.d2);
                if (!(fd * sdir > 0.0)) bad = true;
                Kc = fmax(Kc, fabs(fdd / (2.0 * fd)));
            }
            double tol = Kc > 0.0 ? sqrt(1e-7 / Kc) : hj;
            if (!(tol <= hj)) tol = hj;
            double q[PIECE_DEG + 1];
#pragma unroll
            for (int d = 0; d <= PIECE_DEG; d++) q[d] = 0.0;
            q[0] = xm;
            if (!bad) {
                double uk[PIECE_DEG + 1], a[PIECE_DEG + 1];
#pragma unroll
                for (int k = 0; k <= PIECE_DEG; k++) {
                    const double xk = fma(xr, cospi((2 * k + 1) / (2.0 * (PIECE_DEG + 1))), xm);
                    a[k] = xk;
                    uk[k] = (c0 + xk * fma(xk, fma(xk, c3, c2), c1) - S.fmid) * S.finv;
                }
#pragma unroll
                for (int jj = 1; jj <= PIECE_DEG; jj++) // Newton divided differences in u
#pragma unroll
                    for (int k = PIECE_DEG; k >= jj; k--) a[k] = (a[k] - a[k - 1]) / (uk[k] - uk[k - jj]);
                q[0] = a[PIECE_DEG]; // monomial coefficients, built by Horner on the Newton form
#pragma unroll
                for (int k = PIECE_DEG - 1; k >= 0; k--) {
#pragma unroll
                    for (int d = PIECE_DEG; d >= 1; d--) q[d] = q[d - 1] - uk[k] * q[d];
                    q[0] = a[k] - uk[k] * q[0];
                }
                // verify between the nodes (Chebyshev extrema and the midpoints between them, ends included)
                double err = 0.0;
#pragma unroll
                for (int k = 0; k <= 2 * (PIECE_DEG + 1); k++) {
                    const double xt = fma(xr, cospi(k / (2.0 * (PIECE_DEG + 1))), xm);
                    const double ut = (c0 + xt * fma(xt, fma(xt, c3, c2), c1) - S.fmid) * S.finv;
                    double xq = q[PIECE_DEG];
#pragma unroll
                    for (int d = PIECE_DEG - 1; d >= 0; d--) xq = fma(xq, ut, q[d]);
                    err = fmax(err, fabs(xq - xt));
                }
                if (!(err <= 0.25 * tol)) bad = true;
            }
            S.q0 = q[0]; S.q1 = q[1]; S.q2 = q[2]; S.q3 = q[3]; S.q4 = q[4]; S.q5 = q[5];
            S.tol = bad ? -1.0 : tol;
        }
        double u[4]; // Phi/(2 pi) mod 1 as double-doubles (hi, lo) for Phi_phi, Phi_r
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double ph = q == 0 ? qP.x : qR.x;
            double a = ph * EMRIFD_INV2PI_HI;
            double er = fma(ph, EMRIFD_INV2PI_HI, -a);
            a -= rint(a);
            er = fma(ph, EMRIFD_INV2PI_LO, er);
            const double hi2 = a + er;
            u[2 * q] = hi2; u[2 * q + 1] = er - (hi2 - a);
        }
        S.mu_hi = fma(dm, u[0], dn * u[2]) - (dir < 0 ? 0.25 : 0.0);
        S.mu_lo = fma(dm, u[1], dn * u[3]);
        S.p1 = -EMRIFD_INV2PI_HI * fma(dm, qP.y, dn * qR.y);
        S.p2 = -EMRIFD_INV2PI_HI * fma(dm, qP.z, dn * qR.z);
        S.p3 = -EMRIFD_INV2PI_HI * fma(dm, qP.w, dn * qR.w);
        S.pad = 0.0;
        p.pieces[(wd.mode_off * MAXBR + r) * (long long)p.lstride + j] = S;
    }
}

// Producer warp, step 1: lane i < gcount turns overlapping group record s_list[i] into a FillEntry (tile-local bin ranges and the
// spline segments they fall in, per side) and the warp numbers the sub-entries of the round (exclusive scan).  Returns their
// total.  rc: the walker's cached records (shared memory) or NULL (read the branch records from global memory).
__device__ __noinline__ int fill_entries(const RecC *rc, const emrifd_branch_t *br, const int *marr, const int *narr, int include_minus_m,
                                         long long zero, const int *s_list, const int *s_rec, int gcount, long long jt0,
                                         long long jt1, Grid g, const double *sK, FillEntry *ent) {
    // (br, marr, narr: this walker's blocks.  Scalars by value: a reference to the kernel parameters would force a
    //  local-memory copy of them)
    const int lane = threadIdx.x & 31;
    int nsub = 0;
    if (lane < gcount) {
        const int r = s_list[lane];
        FillEntry e;
        long long bstart, bend;
        if (rc) {
            const RecC c = rc[r];
            bstart = c.start; bend = c.end; e.ja = c.ja; e.jb = c.jb; e.dir = c.dir; e.mirror = c.mirror; e.dm = c.dm; e.dn = c.dn;
        } else {
            const int rec = s_rec[lane], k = rec / MAXBR;
            const emrifd_branch_t b = br[rec];
            const int mi = marr[k];
            bstart = b.start; bend = b.end; e.ja = b.ja; e.jb = b.jb; e.dir = b.dir;
            e.mirror = (mi > 0) && include_minus_m; e.dm = (double)mi; e.dn = (double)narr[k];
        }
        e.r = r; e.pad = 0;
        // tile-local covered ranges: +f bins have full-grid index zero + jt0 + lb, -f bins zero - jt0 - lb
        const long long ntile = jt1 - jt0 + 1;
        long long lo = bstart - (zero + jt0), hi = bend - (zero + jt0);
        e.s[0] = (int)(lo < 0 ? 0 : (lo > ntile ? ntile : lo));
        e.e[0] = (int)(hi > ntile - 1 ? ntile - 1 : (hi < -1 ? -1 : hi));
        lo = (zero - jt0) - bend; hi = (zero - jt0) - bstart;
        e.s[1] = (int)(lo < 0 ? 0 : (lo > ntile ? ntile : lo));
        if (jt0 == 0 && e.s[1] == 0) e.s[1] = 1; // f = 0 is handled on the + side
        e.e[1] = (int)(hi > ntile - 1 ? ntile - 1 : (hi < -1 ? -1 : hi));
#pragma unroll
        for (int sd = 0; sd < 2; sd++) {
            e.jlo[sd] = 1; e.jhi[sd] = 0;
            if (e.s[sd] > e.e[sd]) continue;
            int jj2[2];
            { // segment of the side's first bin: largest j whose knot frequency is not beyond f (binary search) ...
                const double fb = tile_binf(g, jt0, e.s[sd]);
                const double f = sd == 0 ? fb : -fb;
                int l2 = e.ja, h2 = e.jb;
                while (l2 < h2) {
                    const int mid = (l2 + h2 + 1) >> 1;
                    const double Fk = knot_F(sK, mid, e.dm, e.dn);
                    if (e.dir > 0 ? (Fk <= f) : (Fk >= f)) l2 = mid; else h2 = mid - 1;
                }
                jj2[0] = l2;
            }
            { // ... and of its last bin: a short walk from there (a tile rarely spans more than a few segments)
                const double fb = tile_binf(g, jt0, e.e[sd]);
                const double f = sd == 0 ? fb : -fb;
                int l2 = jj2[0];
                while (l2 < e.jb) { const double Fk = knot_F(sK, l2 + 1, e.dm, e.dn); if (e.dir > 0 ? (Fk <= f) : (Fk >= f)) l2++; else break; }
                while (l2 > e.ja) { const double Fk = knot_F(sK, l2, e.dm, e.dn); if (e.dir > 0 ? (Fk <= f) : (Fk >= f)) break; l2--; }
                jj2[1] = l2;
            }
            e.jlo[sd] = jj2[0] < jj2[1] ? jj2[0] : jj2[1];
            e.jhi[sd] = jj2[0] < jj2[1] ? jj2[1] : jj2[0];
            nsub += e.jhi[sd] - e.jlo[sd] + 1;
        }
        e.nsub = nsub;
        e.off = 0;
        ent[lane] = e;
    }
    int incl = nsub;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane < gcount) ent[lane].off = incl - nsub;
    __syncwarp();
    return __shfl_sync(0xffffffffu, incl, 31);
}

// ---- mbarrier / TMA helpers (CTA-local producer/consumer hand-off of the sub-entry ring) ----
__device__ __forceinline__ unsigned smem_u32(const void *ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// the producer warp's wait for a free slot: it is ahead of the consumers most of the time, so it sleeps between polls instead of
// taking issue slots from the consumer warps that share its scheduler
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(400);
    }
}
// global -> shared bulk copy by the TMA; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_copy(void *dst_smem, const void *src_global, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Producer warp, step 2: lane t < nsw builds sub-entry w0 + t of the round in `sub`: finds its (record, side, segment) from the
// entries' offsets, the tile-local bins that fall on that segment (the 16-byte header), and has the TMA copy the piece itself.
__device__ __noinline__ void fill_subs(const Piece *pieces, int lstride, int gcount, int w0, int nsw, Grid g, long long jt0,
                                       const double *sK, const FillEntry *ent, SubEntry *sub, unsigned long long *full) {
    // (pieces: this walker's block of the piece table)
    const int t = threadIdx.x & 31;
    if (t >= nsw) return;
    const int idx = w0 + t;
    int ei = 0;
    while (ei < gcount - 1 && idx >= ent[ei].off + ent[ei].nsub) ei++;
    const FillEntry e = ent[ei];
    const int loc = idx - e.off;
    const int n0 = e.jhi[0] >= e.jlo[0] ? e.jhi[0] - e.jlo[0] + 1 : 0;
    const int sd = loc >= n0 ? 1 : 0;
    const int j = e.jlo[sd] + (sd ? loc - n0 : loc);
    tma_copy(&sub[t].P, pieces + ((long long)e.r * lstride + j), (unsigned)sizeof(Piece), full);
    const bool fwd = (e.dir > 0) == (sd == 0); // bin index and segment index grow together
    const double sg = sd == 0 ? 1.0 : -1.0;
    // ---- bins of this side that fall on segment j: P(lb, jj) = "bin lb lies on a segment >= jj" is monotone in lb ----
    int bnd[2]; // first bin with P(., j) [fwd] / first bin without P(., j + 1) [!fwd]; then the bin after the last one
#pragma unroll
    for (int w = 0; w < 2; w++) {
        const int jj = fwd ? j + w : j + 1 - w;
        const bool edge = fwd ? (w == 0 ? j == e.jlo[sd] : j == e.jhi[sd]) : (w == 0 ? j == e.jhi[sd] : j == e.jlo[sd]);
        if (edge) { bnd[w] = w == 0 ? e.s[sd] : e.e[sd] + 1; continue; }
        const double Fk = knot_F(sK, jj, e.dm, e.dn);
        int lo = e.s[sd], hi = e.e[sd] + 1; // first lb in [lo, hi] where P flips
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double f = sg * tile_binf(g, jt0, mid);
            const bool P = e.dir > 0 ? (Fk <= f) : (Fk >= f);
            if (P == fwd) hi = mid; else lo = mid + 1;
        }
        bnd[w] = lo;
    }
    int4 hdr;
    hdr.x = bnd[0]; hdr.y = bnd[1] - 1;
    hdr.z = sd == 0 ? 0 : (int)0x80000000u;
    hdr.w = (sd ? SE_SIDE : 0) | (e.mirror ? SE_MIRROR : 0) | (e.dir < 0 ? SE_FALL : 0);
    *reinterpret_cast<int4 *>(&sub[t]) = hdr;
}

#ifdef SUM_STATS
__device__ unsigned long long g_stats[16]; // debug build only: 0 tiles, 1 passes, 2 sub-entries, 3 (warp, sub-entry) visits, 4 pair evaluations, 5 bins accumulated, 6 robust pair solves, 7 empty passes
#define STAT_ADD(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define STAT_ADD(i, v) do { } while (0)
#endif
#define PASS_FIRST 1 /* first pass of its tile: the consumers clear their accumulators and compute their bin frequencies */
#define PASS_LAST 2  /* last pass of its tile: the consumers read out after evaluating it */
#define PASS_QUIT 4  /* no more tiles */
struct SumShared { // static shared memory of the mode-sum CTA
    unsigned long long full[SUM_RING], empty[SUM_RING]; // mbarriers: pass published and its pieces landed / pass released by all consumer warps
    int4 hdr[SUM_RING];                                 // (tile, walker, sub-entries in the pass, PASS_* flags)
    int list[64], rec[64];                              // producer: overlapping group records waiting for a fill round
    double2 rot[64];                                    // (cos, sin)(2 pi k / 64): the consumers' sincos table
};

// ------------------------------------------------------------------------------------------------------------------
// Producer warp.  For every tile it owns (persistent launch: pulled from the queue empty_tile_kernel filled; direct launch: the
// CTA's own tile) it scans the walker's group records in order, turns the overlapping ones into sub-entries (fill_entries /
// fill_subs, 32 per pass: a header each, the piece bodies by TMA) and publishes the passes through the ring.  A walker's
// knots and group records are cached in shared memory when the walker changes (the queue hands out a walker's tiles
// consecutively), so a tile costs no dependent global load.  The producer runs ahead of the consumers by up to SUM_RING passes
// -- across tile boundaries.  A pass is published one step late (when the next one has been built, or the tile ends), so that
// the last pass of a tile can carry PASS_LAST.
// ------------------------------------------------------------------------------------------------------------------
template <bool LIKE, int BPT, bool PERSISTENT>
__device__ __forceinline__ void sum_producer(const SumParams &p, SumShared &sh, SubEntry *ring, FillEntry *ent, RecC *sR, double *sK) {
    constexpr int TILE = SUM_CT * BPT;
    const int lane = threadIdx.x & 31;
    const long long zero = p.g.zero;
    const long long jend = p.j_lo + p.j_cnt; // exclusive
    int it = 0;                              // passes acquired so far
    int pending = -1;                        // slot built but not yet published
    int staged_walker = -1;
    const unsigned long long none = ~0ull;
    const unsigned int nq = PERSISTENT ? p.qctl[0] : 1u;
    unsigned long long q = none, qn = none;
    if (PERSISTENT) {
        if (lane == 0) { const unsigned int i0 = atomicAdd(&p.qctl[1], 1u); q = i0 < nq ? p.queue[i0] : none; }
        q = __shfl_sync(0xffffffffu, q, 0);
    } else {
        q = ((unsigned long long)blockIdx.y << 32) | blockIdx.x;
    }
    // per-walker state, reloaded only when the walker changes
    emrifd_walker_t wd;
    memset(&wd, 0, sizeof(wd));
    int G = 0, nrec = 0, wbad = 0;
    bool cached = false; // the walker's group records are in shared memory
    while (q != none) {
        if (PERSISTENT) { // the next item: in flight during this tile
            if (lane == 0) { const unsigned int i1 = atomicAdd(&p.qctl[1], 1u); qn = i1 < nq ? p.queue[i1] : none; }
        }
        const int tile_x = (int)(q & 0xffffffffu), walker_y = (int)(q >> 32);
        const long long jt0 = p.j_lo + (p.tile_first + (long long)tile_x * p.tile_stride) * TILE;
        const long long jt1 = (jt0 + TILE < jend ? jt0 + TILE : jend) - 1; // inclusive
        const long long pos_lo = zero + jt0, pos_hi = zero + jt1, neg_lo = zero - jt1, neg_hi = zero - jt0;
        if (!PERSISTENT) { // one tile per CTA: look at the chunk hulls before staging anything
            const int nrec_ = p.gcount[walker_y] * MAXBR;
            const long long *crng = p.chunk_rng + (long long)walker_y * p.cpw * 2;
            bool any_ = false;
            for (int ch = 0; ch * SUM_CHUNK < nrec_; ch++) any_ |= !(crng[2 * ch] > jt1 || crng[2 * ch + 1] < jt0);
            if (p.wstatus[walker_y]) any_ = false;
            if (!any_ && !(LIKE && (p.no_empty || tile_truncated(p, jt0, TILE)))) break; // empty_tile_kernel has dealt with this tile
        }
        if (staged_walker != walker_y) {
            staged_walker = walker_y;
            wd = p.w[walker_y];
            G = p.gcount[walker_y];
            nrec = G * MAXBR;
            wbad = p.wstatus[walker_y]; // failed walker: zeros (and a NaN likelihood from like_finalize_kernel)
            const emrifd_branch_t *br_ = p.br + wd.mode_off * MAXBR;
            const int *leader_ = p.leader + wd.mode_off;
            const double *coeff_ = p.coeff + wd.coeff_off;
            // knots (time, f_phi, f_r)
            const double *t = p.t + wd.knot_off;
            const int R = 2 * wd.K + 4;
            for (int i = lane; i < wd.L; i += 32) {
                sK[3 * i] = t[i];
                sK[3 * i + 1] = coeff_[((long long)i * R + 2 * wd.K) * 4];
                sK[3 * i + 2] = coeff_[((long long)i * R + 2 * wd.K + 1) * 4];
            }
            // group records (a one-tile CTA reads the few it needs straight from global memory instead)
            cached = PERSISTENT && nrec <= SUM_RCAP;
            for (int r = lane; cached && r < nrec; r += 32) {
                const int lead = leader_[r / MAXBR];
                const emrifd_branch_t b = br_[lead * MAXBR + (r % MAXBR)];
                const int mi = p.m[wd.mode_off + lead];
                RecC c;
                c.start = b.start; c.end = b.end; c.ja = b.ja; c.jb = b.jb; c.dir = b.dir;
                c.mirror = (mi > 0) && p.include_minus_m; c.dm = (double)mi; c.dn = (double)p.n[wd.mode_off + lead];
                long long lo = 0x7fffffffLL, hi = -1;
                if (b.end >= b.start) {
                    if (b.end >= zero) { lo = (b.start > zero ? b.start : zero) - zero; hi = b.end - zero; }
                    if (b.start <= zero) {
                        const long long a_ = zero - (b.end < zero ? b.end : zero), c_ = zero - b.start;
                        lo = a_ < lo ? a_ : lo; hi = c_ > hi ? c_ : hi;
                    }
                }
                c.lo = (int)lo; c.hi = (int)hi; c.pad0 = c.pad1 = 0;
                sR[r] = c;
            }
            __syncwarp();
        }
        const bool per_bin = LIKE && (p.no_empty || tile_truncated(p, jt0, TILE));
        bool any = !wbad;
        if (any && !cached) { // the chunk hulls tell whether anything touches the tile
            const long long *crng = p.chunk_rng + (long long)walker_y * p.cpw * 2;
            any = false;
            for (int ch = 0; ch * SUM_CHUNK < nrec; ch++) any |= !(crng[2 * ch] > jt1 || crng[2 * ch + 1] < jt0);
        }
        if (any || per_bin) { // (else: direct-grid launch on a tile empty_tile_kernel has dealt with)
            const emrifd_branch_t *br = p.br + wd.mode_off * MAXBR;
            const int *leader = p.leader + wd.mode_off;
            const Piece *pieces = p.pieces + wd.mode_off * MAXBR * (long long)p.lstride;
            if (lane == 0) STAT_ADD(0, 1);
            int pflags = PASS_FIRST; // flags of the next pass built for this tile
            int count = 0;           // overlapping records waiting in sh.list
            // a fill round: the first (up to) 32 waiting records, one per lane, become sub-entries; passes of SUM_SUBCAP go out
            auto fill_round = [&]() {
                const int gcount = count < 32 ? count : 32;
                const int tot = fill_entries(cached ? sR : nullptr, br, p.m + wd.mode_off, p.n + wd.mode_off, p.include_minus_m, zero,
                                             sh.list, sh.rec, gcount, jt0, jt1, p.g, sK, ent);
                for (int w0 = 0; w0 < tot; w0 += SUM_SUBCAP) {
                    const int nsw = tot - w0 < SUM_SUBCAP ? tot - w0 : SUM_SUBCAP;
                    if (pending >= 0) { if (lane == 0) mbar_arrive(&sh.full[pending]); pending = -1; }
                    const int slot = it % SUM_RING;
                    mbar_wait_relaxed(&sh.empty[slot], ((it / SUM_RING) & 1) ^ 1);
                    it++;
                    if (lane == 0) {
                        mbar_expect_tx(&sh.full[slot], (unsigned)(nsw * sizeof(Piece)));
                        sh.hdr[slot] = make_int4(tile_x, walker_y, nsw, pflags);
                        STAT_ADD(1, 1); STAT_ADD(2, nsw);
                    }
                    __syncwarp();
                    fill_subs(pieces, p.lstride, gcount, w0, nsw, p.g, jt0, sK, ent, ring + slot * SUM_SUBCAP, &sh.full[slot]);
                    pflags = 0;
                    __syncwarp();
                    pending = slot;
                }
                // drop the records of this round from the list
                const int left = count - gcount;
                int mv_l = 0, mv_r = 0;
                if (lane < left) { mv_l = sh.list[32 + lane]; mv_r = sh.rec[32 + lane]; }
                __syncwarp();
                if (lane < left) { sh.list[lane] = mv_l; sh.rec[lane] = mv_r; }
                __syncwarp();
                count = left;
            };
            if (cached) { // ordered compaction of the group records that overlap the tile, from the shared-memory table
                const int t0 = (int)jt0, t1 = (int)jt1;
                for (int base = 0; any && base < nrec; base += 32) {
                    const int r = base + lane;
                    bool pred = false;
                    if (r < nrec) pred = !(sR[r].lo > t1 || sR[r].hi < t0);
                    const unsigned bal = __ballot_sync(0xffffffffu, pred);
                    if (pred) { const int pos = count + __popc(bal & ((1u << lane) - 1)); sh.list[pos] = r; sh.rec[pos] = 0; }
                    count += __popc(bal);
                    __syncwarp();
                    if (count >= 32) fill_round();
                }
            } else {
                const long long *crng = p.chunk_rng + (long long)walker_y * p.cpw * 2;
                for (int base = 0; any && base < nrec; base += 32) {
                    const int ch = base / SUM_CHUNK;
                    if (crng[2 * ch] > jt1 || crng[2 * ch + 1] < jt0) { base = (ch + 1) * SUM_CHUNK - 32; continue; } // chunk outside the tile
                    const int r = base + lane;
                    bool pred = false;
                    int rec = 0;
                    if (r < nrec) {
                        rec = leader[r / MAXBR] * MAXBR + (r % MAXBR);
                        const long long s0 = br[rec].start, e0 = br[rec].end;
                        pred = (e0 >= s0) && ((s0 <= pos_hi && e0 >= pos_lo) || (s0 <= neg_hi && e0 >= neg_lo));
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, pred);
                    if (pred) { const int pos = count + __popc(bal & ((1u << lane) - 1)); sh.list[pos] = r; sh.rec[pos] = rec; }
                    count += __popc(bal);
                    __syncwarp();
                    if (count >= 32) fill_round();
                }
            }
            if (count > 0) fill_round();
            // ---- end of the tile: its last pass carries PASS_LAST (a tile without sub-entries gets an empty pass) ----
            if (pending < 0) {
                const int slot = it % SUM_RING;
                mbar_wait_relaxed(&sh.empty[slot], ((it / SUM_RING) & 1) ^ 1);
                it++;
                if (lane == 0) { sh.hdr[slot] = make_int4(tile_x, walker_y, 0, PASS_FIRST | PASS_LAST); STAT_ADD(7, 1); mbar_arrive(&sh.full[slot]); }
            } else {
                if (lane == 0) { sh.hdr[pending].w |= PASS_LAST; mbar_arrive(&sh.full[pending]); }
                pending = -1;
            }
        }
        if (PERSISTENT) q = __shfl_sync(0xffffffffu, qn, 0); else q = none;
    }
    { // no more tiles
        const int slot = it % SUM_RING;
        mbar_wait_relaxed(&sh.empty[slot], ((it / SUM_RING) & 1) ^ 1);
        if (lane == 0) { sh.hdr[slot] = make_int4(0, 0, 0, PASS_QUIT); mbar_arrive(&sh.full[slot]); }
    }
}

// ---- A6/A7: S = -flip(W); h+ = (S + conj flip S)/2; hx = i (S - conj flip S)/2; scale; rotate ----
// Straight from the accumulator registers: a thread finalises its own two adjacent bins (32 contiguous bytes per array,
// a warp covers 1 KB); the data loads of both bins are issued together, ahead of the stores.
template <bool WRITE_H, bool LIKE>
__device__ __forceinline__ void sum_readout(const SumParams &p, const int tile_x, const int walker_y, const long long j0,
                                            const int nb, const double (&wp_r)[2], const double (&wp_i)[2],
                                            const double (&wm_r)[2], const double (&wm_i)[2]) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const emrifd_walker_t *wp = p.w + walker_y;
    const double sc = wp->scale, c2 = wp->cos2psi, s2 = wp->sin2psi;
    const long long out_off = wp->out_off;
    const long long zero = p.g.zero;
    double a0 = 0, a1 = 0, a2 = 0;
    double2 d0[2], d1[2];
    double w0[2], w1[2];
    if (LIKE) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const long long j = j0 + (b < nb ? b : 0);
            const long long jj = nb > 0 ? j : 0;
            d0[b] = p.dw[jj]; d1[b] = p.dw[p.n_data + jj];
            w0[b] = p.wf[jj]; w1[b] = p.wf[p.n_data + jj];
        }
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {
        if (b >= nb) continue;
        const long long j = j0 + b;
        double wpr = wp_r[b], wpi = wp_i[b], wmr = wm_r[b], wmi = wm_i[b];
        if (j == 0) { wpr += wmr; wpi += wmi; wmr = wpr; wmi = wpi; }
        const double pr_ = 0.5 * (-wmr - wpr), pi_ = 0.5 * (-wmi + wpi);
        const double xr_ = 0.5 * (wmi + wpi), xi_ = 0.5 * (-wmr + wpr);
        const double hpr = sc * (c2 * pr_ - s2 * xr_), hpi = sc * (c2 * pi_ - s2 * xi_);
        const double hxr = sc * (s2 * pr_ + c2 * xr_), hxi = sc * (s2 * pi_ + c2 * xi_);
        if (WRITE_H) {
            if (p.mask_positive) {
                const long long o = out_off + (j - p.j_lo);
                p.hp[o] = make_double2(hpr, hpi);
                p.hc[o] = make_double2(hxr, hxi);
            } else {
                const long long o = out_off + zero;
                p.hp[o + j] = make_double2(hpr, hpi);
                p.hc[o + j] = make_double2(hxr, hxi);
                if (j > 0) { // Hermitian mirror: h(-f) = conj h(f)
                    p.hp[o - j] = make_double2(hpr, -hpi);
                    p.hc[o - j] = make_double2(hxr, -hxi);
                }
            }
        }
        if (LIKE) {
            const double h0r = hpr * w0[b], h0i = hpi * w0[b], h1r = hxr * w1[b], h1i = hxi * w1[b];
            const double r0 = d0[b].x - h0r, q0 = d0[b].y - h0i, r1 = d1[b].x - h1r, q1 = d1[b].y - h1i;
            a0 += r0 * r0 + q0 * q0 + r1 * r1 + q1 * q1;
            a1 += d0[b].x * h0r + d0[b].y * h0i + d1[b].x * h1r + d1[b].y * h1i;
            a2 += h0r * h0r + h0i * h0i + h1r * h1r + h1i * h1i;
        }
    }
    if (LIKE) { // per-warp partial sums; like_finalize_kernel adds them in a fixed order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, o);
            a1 += __shfl_down_sync(0xffffffffu, a1, o);
            a2 += __shfl_down_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) {
            double *o = p.partial + (((long long)walker_y * p.ntiles + tile_x) * SUM_CW + wid) * 3;
            o[0] = a0; o[1] = a1; o[2] = a2;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Consumer warp: owns 64 consecutive bins of every tile the CTA handles; a thread owns two adjacent (+f, -f) bin pairs whose
// four complex accumulators W(+-f) live in registers for the whole tile.  It waits for the next pass, evaluates its two bins
// on each of the pass's cubic pieces that covers them -- root from the piece's inverse interpolant + one Newton step (no state
// carried from bin to bin, both bins in straight-line code), SPA factor, phase, amplitude -- releases the slot, and reads out
// after a tile's last pass.  Consumer warps never wait for one another: a warp with less work in a tile runs ahead, up to
// SUM_RING passes, into the following tiles.
// ------------------------------------------------------------------------------------------------------------------
template <bool WRITE_H, bool LIKE>
__device__ __forceinline__ void sum_consumer(const SumParams &p, SumShared &sh, const SubEntry *ring) {
    constexpr int TILE = SUM_CT * 2;
    const int tid = threadIdx.x, lane = tid & 31;
    const long long jend = p.j_lo + p.j_cnt; // exclusive
    double fb[2] = {0.0, 0.0};               // this thread's two bin frequencies
    double wp_r[2] = {0.0, 0.0}, wp_i[2] = {0.0, 0.0}, wm_r[2] = {0.0, 0.0}, wm_i[2] = {0.0, 0.0}; // W(+f), W(-f)
    for (int it = 0;; it++) {
        const int slot = it % SUM_RING;
        mbar_wait(&sh.full[slot], (it / SUM_RING) & 1);
        const int4 hd = sh.hdr[slot];
        if (hd.w & PASS_QUIT) return;
        const int nsw = hd.z;
        int nb;
        {
            const long long jt0 = p.j_lo + (p.tile_first + (long long)hd.x * p.tile_stride) * TILE;
            const long long jt1 = (jt0 + TILE < jend ? jt0 + TILE : jend) - 1; // inclusive
            const long long j0 = jt0 + 2 * tid;                                 // this thread's first bin
            nb = (int)(jt1 - j0 + 1);                                           // its number of valid bins
            nb = nb < 0 ? 0 : (nb > 2 ? 2 : nb);
            if (hd.w & PASS_FIRST) {
                if (LIKE && nb > 0) { // the read-out's data lines: start them towards L2 now (a third of them come from DRAM)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dw + j0));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dw + p.n_data + j0));
                    if ((lane & 1) == 0) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.wf + j0));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.wf + p.n_data + j0));
                    }
                }
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    wp_r[b] = wp_i[b] = wm_r[b] = wm_i[b] = 0.0;
                    const long long jj = j0 + b;
                    fb[b] = (b < nb) ? (p.g.fpos ? p.g.fpos[jj] : rmul((double)(int)jj, p.g.val)) : 0.0;
                }
                if (nb == 1) fb[1] = fb[0]; // truncated tile: the missing second bin shadows the first
            }
        }
        if (nb > 0) {
            // the thread's bin range as one opaque register pair (kept live instead of being re-derived from %tid every pass)
            int tb0 = 2 * tid, tb1 = 2 * tid + nb - 1;
            asm volatile("" : "+r"(tb0), "+r"(tb1));
            const SubEntry *Ep = ring + slot * SUM_SUBCAP;
            const SubEntry *const Eend = Ep + nsw;
            for (; Ep != Eend; ++Ep) {
                const SubEntry &E = *Ep;
                const Piece &S = E.P;
                const int es = E.s, ee = E.e;
                if (es > tb1 || ee < tb0) continue; // the piece does not touch this thread's bins
                bool in[2];
                in[0] = es <= tb0;                  // (then ee >= tb0 holds)
                in[1] = ee > tb0 && nb > 1;         // (then es <= tb0 + 1 holds)
                STAT_ADD(4, 1); STAT_ADD(5, (int)in[0] + (int)in[1]);
                const unsigned int smask = E.fmask;
                const int fl = E.flags;
                double f[2], x[2];
                // (a bin just outside the piece is evaluated all the same -- one bin beyond the piece's end is still a perfectly
                //  regular point of its cubic and interpolant -- and not accumulated)
                f[0] = flip_sign(fb[0], smask);
                f[1] = flip_sign(fb[1], smask);
                bool ok;
                {
                    const double c0 = S.c0, c1 = S.c1, c2 = S.c2, c3 = S.c3, d2 = S.d2, d3 = S.d3, tol = S.tol;
                    const double fmid = S.fmid, finv = S.finv, q0 = S.q0, q1 = S.q1, q2 = S.q2, q3 = S.q3, q4 = S.q4, q5 = S.q5;
                    double dx[2];
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const double u = (f[i] - fmid) * finv;
                        const double x0 = fma(u, fma(u, fma(u, fma(u, fma(u, q5, q4), q3), q2), q1), q0);
                        const double g0 = x0 * fma(x0, fma(x0, c3, c2), c1) - (f[i] - c0);
                        dx[i] = g0 * fast_rcp(fma(x0, fma(d3, x0, d2), c1));
                        x[i] = x0 - dx[i];
                    }
                    ok = fabs(dx[0]) <= tol && fabs(dx[1]) <= tol; // (tol < 0: the piece has no usable interpolant)
                }
                if (!ok) { // turnover neighbourhood, or an interpolant that missed: bracketed solver
                    STAT_ADD(6, 1);
                    const double sdir = (fl & SE_FALL) ? -1.0 : 1.0;
                    const double xlo = S.xlo, xhi = S.xhi, tolr = 1e-6 * (xhi - xlo);
                    x[0] = solve_slow(S.c1, S.c2, S.c3, f[0] - S.c0, xlo, xhi, tolr, sdir);
                    x[1] = solve_slow(S.c1, S.c2, S.c3, f[1] - S.c0, xlo, xhi, tolr, sdir);
                }
                if (fl & SE_SIDE) eval_sub<2>(x, f, in, S, fl, p.k13_few, sh.rot, wm_r, wm_i, wp_r, wp_i); // bins at -f: direct -> W(-f)
                else eval_sub<2>(x, f, in, S, fl, p.k13_few, sh.rot, wp_r, wp_i, wm_r, wm_i);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty[slot]); // the warp is done with this pass's sub-entries
        if (hd.w & PASS_LAST) {
            int tile_x = hd.x;
            asm volatile("" : "+r"(tile_x)); // recompute the tile geometry here rather than keep it live through the evaluation
            const long long jt0 = p.j_lo + (p.tile_first + (long long)tile_x * p.tile_stride) * TILE;
            sum_readout<WRITE_H, LIKE>(p, tile_x, hd.y, jt0 + 2 * tid, nb, wp_r, wp_i, wm_r, wm_i);
        }
    }
}

// Persistent CTAs (one grid of #SM x resident CTAs) work through the non-empty tiles queued by empty_tile_kernel: the heavy
// kernel is never launched on the > 90 % of the band that a sparse system leaves empty, and tiles of very different cost
// balance dynamically.  PERSISTENT = false: a plain (tile, walker) grid for launches of about one wave (a bin-sharded slice of
// a single long waveform, a single short waveform), where the hardware's breadth-first CTA placement balances the SMs better
// than queue order (measured on the 8-GPU bin-sharded configs[3] slices).
template <bool WRITE_H, bool LIKE, int BPT, bool PERSISTENT>
__global__ void SUM_BOUNDS mode_sum_kernel(SumParams p) {
    static_assert(BPT == 2, "a consumer thread owns one pair of bins");
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ SumShared sh;
    // dynamic smem: sub-entry ring | fill entries | group records [SUM_RCAP] | knots (t, f_phi, f_r)[L]
    SubEntry *ring = reinterpret_cast<SubEntry *>(smraw);
    FillEntry *ent = reinterpret_cast<FillEntry *>(ring + SUM_RING * SUM_SUBCAP);
    RecC *sR = reinterpret_cast<RecC *>(ent + SUM_ECAP);
    double *sK = reinterpret_cast<double *>(sR + SUM_RCAP);
    if (threadIdx.x < 64) sh.rot[threadIdx.x] = make_double2(c_rot64[2 * threadIdx.x], c_rot64[2 * threadIdx.x + 1]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < SUM_RING; i++) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], SUM_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads(); // the only CTA-wide barrier: from here on the warps synchronise through the ring's mbarriers alone
    if (threadIdx.x >= SUM_CT) sum_producer<LIKE, BPT, PERSISTENT>(p, sh, ring, ent, sR, sK);
    else sum_consumer<WRITE_H, LIKE>(p, sh, ring);
}

// deterministic second stage: one CTA per walker
#define FIN_THREADS 1024
__global__ void __launch_bounds__(FIN_THREADS) like_finalize_kernel(const double *__restrict__ partial, long long ntiles,
                                                                    double *__restrict__ out, const int *__restrict__ wstatus) {
    __shared__ double s[3][FIN_THREADS];
    const double *pp = partial + (long long)blockIdx.x * ntiles * 3;
    double a0 = 0, a1 = 0, a2 = 0;
    for (long long i = threadIdx.x; i < ntiles; i += FIN_THREADS) { a0 += pp[3 * i]; a1 += pp[3 * i + 1]; a2 += pp[3 * i + 2]; }
    s[0][threadIdx.x] = a0; s[1][threadIdx.x] = a1; s[2][threadIdx.x] = a2;
    __syncthreads();
    for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s[0][threadIdx.x] += s[0][threadIdx.x + o];
            s[1][threadIdx.x] += s[1][threadIdx.x + o];
            s[2][threadIdx.x] += s[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // a walker whose trajectory or work-list was refused reports NaN (Eryn maps NaN to -1e300, red_blue.py:282-284)
        const bool bad = wstatus && wstatus[blockIdx.x];
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        out[3 * blockIdx.x + 0] = bad ? nan : -0.5 * 4.0 * s[0][0]; // ll = -1/2 * 4 * sum |d~ - h~|^2 (likelihood.py:270-274)
        out[3 * blockIdx.x + 1] = bad ? nan : 4.0 * s[1][0];
        out[3 * blockIdx.x + 2] = bad ? nan : 4.0 * s[2][0];
    }
}

// ==========================================================================================
// A10 / A11 on materialised arrays
// ==========================================================================================
__global__ void __launch_bounds__(256) inner_partial_kernel(const double2 *__restrict__ a, const double2 *__restrict__ b,
                                                            long long nch, long long n, const double *__restrict__ f,
                                                            const double *__restrict__ psd, double *__restrict__ partial) {
    __shared__ double s[2][8];
    double re = 0, im = 0;
    const long long tot = nch * n;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < tot; i += (long long)gridDim.x * 256) {
        const long long k = i % n;
        const double dx = (k == 0) ? (f[1] - f[0]) : (f[k] - f[k - 1]);
        const double2 x = a[i], y = b[i];
        const double wgt = psd ? dx / psd[k] : dx;
        // conj(a) * b
        re += wgt * (x.x * y.x + x.y * y.y);
        im += wgt * (x.x * y.y - x.y * y.x);
    }
    for (int o = 16; o > 0; o >>= 1) { re += __shfl_down_sync(0xffffffffu, re, o); im += __shfl_down_sync(0xffffffffu, im, o); }
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = re; s[1][threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0;
        for (int q = 0; q < 8; q++) { t0 += s[0][q]; t1 += s[1][q]; }
        partial[2 * blockIdx.x] = t0; partial[2 * blockIdx.x + 1] = t1;
    }
}
__global__ void inner_final_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out) {
    __shared__ double s[2][256];
    double a = 0, b = 0;
    for (int i = threadIdx.x; i < nb; i += 256) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    s[0][threadIdx.x] = a; s[1][threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s[0][threadIdx.x] += s[0][threadIdx.x + o]; s[1][threadIdx.x] += s[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = 4.0 * s[0][0]; out[1] = 4.0 * s[1][0]; }
}

// templates [B][2][n] complex vs stored data -> partial [B][nblk][3]
__global__ void __launch_bounds__(256) loglike_partial_kernel(const double2 *__restrict__ tmpl, const double2 *__restrict__ dw,
                                                              const double *__restrict__ wf, long long n,
                                                              double *__restrict__ partial) {
    __shared__ double s[3][8];
    const double2 *h = tmpl + (long long)blockIdx.y * 2 * n;
    double a0 = 0, a1 = 0, a2 = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < 2 * n; i += (long long)gridDim.x * 256) {
        const double2 hv = h[i], d = dw[i];
        const double w = wf[i];
        const double hr = hv.x * w, hi = hv.y * w;
        const double r = d.x - hr, q = d.y - hi;
        a0 += r * r + q * q; a1 += d.x * hr + d.y * hi; a2 += hr * hr + hi * hi;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o); a1 += __shfl_down_sync(0xffffffffu, a1, o); a2 += __shfl_down_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = a0; s[1][threadIdx.x >> 5] = a1; s[2][threadIdx.x >> 5] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0, t2 = 0;
        for (int q = 0; q < 8; q++) { t0 += s[0][q]; t1 += s[1][q]; t2 += s[2][q]; }
        double *o = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 3;
        o[0] = t0; o[1] = t1; o[2] = t2;
    }
}

// ==========================================================================================
// Ylm, mode selection by power and mode compaction on the device.
//
// Selection (few.utils.modeselector.ModeSelector semantics), frozen spec shared with the numpy
// restatement oracle/oracle.py::mode_select_ref:
//   power_i = |A_i Y_i|^2 over the M stored modes plus the -m copies (conj(A) Y_{l,-m}) of the m > 0 ones, every
//   operation individually rounded; total = sum in the fixed order (256 strided partials, shuffle tree, 8 warps);
//   order: descending power, ties by ascending index; entry s of that order is kept iff s == 0 or the sequential
//   cumulative sum of the entries before it is < total * (1 - eps); a kept -m copy keeps its +m partner; a walker's
//   kept set is the union over its time samples.
// One CTA per time sample.  Entries with power < total * eps / (2 Mtot) cannot be reached by the cumulative sum before
// it crosses the threshold (together they hold < eps/2 of the total), so only the survivors are sorted: a few hundred
// at eps = 1e-2 instead of 7137.
// ==========================================================================================
#define SEL_THREADS 256
#define SEL_CAP 8192 /* >= M + Mneg (7137 for l <= 10, |n| <= 30) */
#define SEL_CAP_SMALL 1024 /* sort capacity of the first launch */

struct SelParams {
    const double2 *teuk;     // [nsamp][M]
    const int *samp_walker;  // [nsamp]
    const double2 *ylm;      // [B][M + Mneg]
    const int *neg_src;      // [Mneg] index of the +m mode each -m copy belongs to
    int M, Mneg, B;
    double eps;
    unsigned char *flags;    // [B][M], zero-initialised by the caller
    int cap;                 // sort capacity of this launch (entries of shared memory)
    int pass;                // 0: every sample, survivors beyond cap -> ovf[samp] = 1 and nothing written; 1: only the samples with ovf set
    unsigned char *ovf;      // [nsamp]
};

__device__ __forceinline__ double sel_power(const SelParams &p, const double2 *A, const double2 *Y, int i) {
    const double2 a = A[i < p.M ? i : p.neg_src[i - p.M]];
    const double ai = i < p.M ? a.y : -a.y; // -m copy: conj(A)
    const double2 y = Y[i];
    const double re = rsub(rmul(a.x, y.x), rmul(ai, y.y)), im = radd(rmul(a.x, y.y), rmul(ai, y.x));
    return radd(rmul(re, re), rmul(im, im));
}

__global__ void __launch_bounds__(SEL_THREADS) mode_select_kernel(SelParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double *key = reinterpret_cast<double *>(smraw);                       // [cap]
    unsigned short *idx = reinterpret_cast<unsigned short *>(key + p.cap); // [cap]
    __shared__ double s_red[SEL_THREADS / 32];
    __shared__ double s_total;
    __shared__ int s_count, s_end;
    const int samp = blockIdx.x, w = p.samp_walker[samp], Mtot = p.M + p.Mneg;
    if (w < 0 || w >= p.B) return; // a sample that names no walker of the batch is ignored (never an out-of-bounds flag write)
    if (p.pass == 1 && !p.ovf[samp]) return; // handled by the small-capacity launch
    const double2 *A = p.teuk + (long long)samp * p.M;
    const double2 *Y = p.ylm + (long long)w * Mtot;
    double part = 0.0;
    for (int i = threadIdx.x; i < Mtot; i += SEL_THREADS) part = radd(part, sel_power(p, A, Y, i));
    for (int o = 16; o > 0; o >>= 1) part = radd(part, __shfl_down_sync(0xffffffffu, part, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int q = 0; q < SEL_THREADS / 32; q++) t = radd(t, s_red[q]); s_total = t; }
    __syncthreads();
    const double total = s_total;
    const double cut = p.eps >= 1e-9 ? rmul(total, rdiv(p.eps, 2.0 * (double)Mtot)) : 0.0;
    // survivors of the pre-filter, compacted in arbitrary order (the sort below is a total order)
    for (int i = threadIdx.x; i < Mtot; i += SEL_THREADS) {
        const double pw = sel_power(p, A, Y, i);
        if (!(pw < cut)) { const int s = atomicAdd(&s_count, 1); if (s < p.cap) { key[s] = pw; idx[s] = (unsigned short)i; } }
    }
    __syncthreads();
    const int cnt = s_count;
    if (p.pass == 0) { // more survivors than this launch can sort: leave the sample to the full-capacity launch
        if (threadIdx.x == 0) p.ovf[samp] = cnt > p.cap ? 1 : 0;
        if (cnt > p.cap) return;
    }
    int n2 = 2;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + threadIdx.x; i < n2; i += SEL_THREADS) { key[i] = -1.0; idx[i] = 0xffff; } // padding sorts last
    // bitonic sort: descending key, ascending index among equal keys
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n2; i += SEL_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double a = key[i], b = key[ixj];
                    const unsigned short ia = idx[i], ib = idx[ixj];
                    const bool a_first = a > b || (a == b && ia < ib);
                    if (((i & k) == 0) != a_first) { key[i] = b; key[ixj] = a; idx[i] = ib; idx[ixj] = ia; }
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { // sequential cumulative sum (numpy.cumsum order): first position that is NOT kept
        const double thresh = rmul(total, rsub(1.0, p.eps));
        double cs = 0.0;
        int s = 0;
        for (; s < cnt; s++) {
            if (s > 0 && !(cs < thresh)) break;
            cs = radd(cs, key[s]);
        }
        s_end = s;
    }
    __syncthreads();
    unsigned char *fl = p.flags + (long long)w * p.M;
    for (int s = threadIdx.x; s < s_end; s += SEL_THREADS) {
        const int i = idx[s];
        fl[i < p.M ? i : p.neg_src[i - p.M]] = 1; // benign race: every writer stores 1
    }
}

// Spin-weight -2 spherical harmonics for a batch of viewing angles (few.utils.ylm.GetYlms(assume_positive_m=True)
// followed by the per-mode expansion of FastSchwarzschildEccentricFlux.__call__): out[w][i] = Y_{l_i, m_i}(theta_w, phi_w)
// for i < M and Y_{l, -m} of mode neg_src[i - M] for the Mneg copies.  -2Y_lm = sqrt((2l+1)/4pi) d^l_{m,2}(theta) e^{i m phi}
// with the explicit finite Wigner sum (exact at theta = 0, pi).  One CTA per walker: the <= 2 (lmax+1)^2 distinct
// (l, +-m) values are computed once into shared memory, then gathered per mode.
#define YLM_LMAX 12
__global__ void __launch_bounds__(256) ylm_kernel(const int *__restrict__ l_arr, const int *__restrict__ m_arr, int M,
                                                  const int *__restrict__ neg_src, int Mneg, const double *__restrict__ theta,
                                                  const double *__restrict__ phi, double2 *__restrict__ out) {
    __shared__ double2 tab[(YLM_LMAX + 1) * (2 * YLM_LMAX + 1)];
    __shared__ double fact[2 * YLM_LMAX + 3];
    const int w = blockIdx.x;
    if (threadIdx.x == 0) { double f = 1.0; fact[0] = 1.0; for (int i = 1; i < 2 * YLM_LMAX + 3; i++) { f *= (double)i; fact[i] = f; } }
    __syncthreads();
    const double th = theta[w], ph = phi[w];
    double sb, cb;
    sincos(0.5 * th, &sb, &cb);
    for (int q = threadIdx.x; q < (YLM_LMAX + 1) * (2 * YLM_LMAX + 1); q += blockDim.x) {
        const int l = q / (2 * YLM_LMAX + 1), mp = q % (2 * YLM_LMAX + 1) - YLM_LMAX;
        double2 v = make_double2(0.0, 0.0);
        if (l >= 2 && mp >= -l && mp <= l) {
            const int m = 2; // second index of d^l_{mp, m}: m = -s = 2
            const double pref = sqrt(fact[l + mp] * fact[l - mp] * fact[l + m] * fact[l - m]);
            const int k0 = max(0, m - mp), k1 = min(l + m, l - mp);
            double tot = 0.0;
            for (int k = k0; k <= k1; k++) {
                const double den = fact[l + m - k] * fact[k] * fact[l - k - mp] * fact[k - m + mp];
                double term = ((k - m + mp) & 1) ? -1.0 / den : 1.0 / den;
                const int ec = 2 * l - 2 * k + m - mp, es = 2 * k - m + mp;
                for (int a = 0; a < ec; a++) term *= cb;
                for (int a = 0; a < es; a++) term *= sb;
                tot += term;
            }
            const double d = sqrt((2.0 * l + 1.0) / (4.0 * 3.14159265358979323846)) * pref * tot;
            double sn, cs;
            sincos((double)mp * ph, &sn, &cs);
            v = make_double2(d * cs, d * sn);
        }
        tab[q] = v;
    }
    __syncthreads();
    double2 *o = out + (long long)w * (M + Mneg);
    for (int i = threadIdx.x; i < M + Mneg; i += blockDim.x) {
        const int src = i < M ? i : neg_src[i - M];
        const int l = l_arr[src], m = i < M ? m_arr[src] : -m_arr[src];
        o[i] = tab[l * (2 * YLM_LMAX + 1) + m + YLM_LMAX];
    }
}

// Compaction step 1: per walker, the ascending list of kept mode indices and its length (block-wide scan of the flags).
__global__ void __launch_bounds__(256) compact_count_kernel(const unsigned char *__restrict__ flags, int M, int *__restrict__ keep_idx,
                                                            int *__restrict__ K_out) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int w = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned char *fl = flags + (long long)w * M;
    int *ki = keep_idx + (long long)w * M;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < M; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const int f = i < M && fl[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int q = 0; q < wid; q++) off += s_warp[q];
        if (f) ki[off + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; q++) t += s_warp[q]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) K_out[w] = s_base;
}

// Compaction step 2: gather the kept modes of every walker into the packed layout the mode-sum pipeline consumes
// (teuk [L][K], m/n [K], ylm [2K] = +m block then -m block; m = 0 modes reuse their +m value in the -m block).
// grid (Lmax + 1, B): block row j < L gathers knot j of teuk; block row Lmax gathers m, n, ylm.
__global__ void __launch_bounds__(256) compact_gather_kernel(const emrifd_walker_t *__restrict__ wk, int Lmax, int M, int Mneg,
                                                             const double2 *__restrict__ teuk_full, const int *__restrict__ keep_idx,
                                                             const int *__restrict__ m_basis, const int *__restrict__ n_basis,
                                                             const int *__restrict__ neg_pos, const double2 *__restrict__ ylm_full,
                                                             double2 *__restrict__ teuk_out, int *__restrict__ m_out,
                                                             int *__restrict__ n_out, double2 *__restrict__ ylm_out) {
    const int w = blockIdx.y, j = blockIdx.x;
    const emrifd_walker_t W = wk[w];
    const int *ki = keep_idx + (long long)w * M;
    if (j < Lmax) {
        if (j >= W.L) return;
        const double2 *src = teuk_full + (W.knot_off + j) * (long long)M;
        double2 *dst = teuk_out + W.teuk_off + (long long)j * W.K;
        for (int k = threadIdx.x; k < W.K; k += blockDim.x) dst[k] = src[ki[k]];
    } else {
        const double2 *Y = ylm_full + (long long)w * (M + Mneg);
        for (int k = threadIdx.x; k < W.K; k += blockDim.x) {
            const int i = ki[k];
            m_out[W.mode_off + k] = m_basis[i];
            n_out[W.mode_off + k] = n_basis[i];
            ylm_out[2 * W.mode_off + k] = Y[i];
            ylm_out[2 * W.mode_off + W.K + k] = Y[neg_pos[i] >= 0 ? M + neg_pos[i] : i];
        }
    }
}

// Stand-in amplitude producer on the device (amplitude/synthetic.py::SyntheticAmplitude, the offline substitute for few's
// RomanAmplitude whose weights are a Zenodo download): A_lmn(p, e) = c_lmn p^{-l/2} exp(-(n - n0)^2 / (2 sigma^2))
// exp(i (e n / 6 + 8 / p)), n0 = 2.5 e (1 + 0.2 m) / sqrt(1 - e), sigma = 0.35 + 3.5 e.  One CTA per trajectory point:
// the separable factors go through small shared-memory tables ([l], [m][n], [n]); every mode is then three look-ups.
#define AMP_LMAX 12
#define AMP_NMAX 32
__global__ void __launch_bounds__(256) synth_amplitude_kernel(const double *__restrict__ p_arr, const double *__restrict__ e_arr,
                                                              const int *__restrict__ l_arr, const int *__restrict__ m_arr,
                                                              const int *__restrict__ n_arr, const double2 *__restrict__ cmode,
                                                              int M, int lmax, int nmax, double2 *__restrict__ out) {
    __shared__ double sP[AMP_LMAX + 1];
    __shared__ double sEnv[(AMP_LMAX + 1) * (2 * AMP_NMAX + 1)];
    __shared__ double2 sPh[2 * AMP_NMAX + 1];
    const long long row = blockIdx.x;
    const double p = p_arr[row], e = e_arr[row];
    const int nn = 2 * nmax + 1;
    for (int l = 2 + threadIdx.x; l <= lmax; l += blockDim.x) sP[l] = pow(p, -0.5 * (double)l);
    const double sig = 0.35 + 3.5 * e, n0b = 2.5 * e / sqrt(1.0 - e), inv2s2 = 1.0 / (2.0 * sig * sig);
    for (int q = threadIdx.x; q < (lmax + 1) * nn; q += blockDim.x) {
        const int m = q / nn, n = q % nn - nmax;
        const double d = (double)n - n0b * (1.0 + 0.2 * (double)m);
        sEnv[q] = exp(-(d * d) * inv2s2);
    }
    for (int q = threadIdx.x; q < nn; q += blockDim.x) {
        double sn, cs;
        sincos((e / 6.0) * (double)(q - nmax) + 8.0 / p, &sn, &cs);
        sPh[q] = make_double2(cs, sn);
    }
    __syncthreads();
    double2 *o = out + row * (long long)M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const int l = l_arr[i], m = m_arr[i], n = n_arr[i] + nmax;
        const double a = sP[l] * sEnv[m * nn + n];
        const double2 ph = sPh[n], c = cmode[i];
        const double pr = ph.x * c.x - ph.y * c.y, pi = ph.x * c.y + ph.y * c.x; // (EPH * cmode), then times the real envelope
        o[i] = make_double2(a * pr, a * pi);
    }
}

// FP64 FMA peak micro-benchmark: 8 independent chains per thread
// ==========================================================================================
// FD window convolution (FDutils.py:35-47,66-101).  The reference convolves the FD channels with the
// conjugated DFT of a time-domain window: out[k] = (1/N) sum_i a[i] b[(k - i) mod N], a = conj(fft(window)).  For the windows the
// scripts use (hann, blackman, hamming, nuttall, ...: check_mode_by_mode.py:43,269) the DFT is concentrated in a few taps
// around i = 0 (mod N), so the convolution is a banded stencil: window_taps_kernel evaluates the 2H + 1 central taps of the
// window's DFT directly (and sum w^2, from which Parseval gives the energy left outside the band = the truncation bound), and
// band_convolve_kernel applies them: one read and one write of the signal, no FFT.
// ==========================================================================================
#define WIN_THREADS 256
// taps W_j = sum_n w[n] e^{-2 pi i j n / N} for j = 0..H (a real window: W_{-j} = conj W_j) as per-block partial sums, plus sum w^2
// in row H + 1.  grid (chunks, H + 2).  The phase j n / N is reduced exactly in integers before the sincos.
__global__ void __launch_bounds__(WIN_THREADS) window_taps_kernel(const double *__restrict__ w, long long N, int H,
                                                                  double *__restrict__ partial) {
    __shared__ double s[2][WIN_THREADS / 32];
    const int j = blockIdx.y;
    double re = 0.0, im = 0.0;
    for (long long n = (long long)blockIdx.x * WIN_THREADS + threadIdx.x; n < N; n += (long long)gridDim.x * WIN_THREADS) {
        const double wn = w[n];
        if (j > H) { re = fma(wn, wn, re); continue; }
        const long long r = ((long long)j * n) % N;
        double sn, cs;
        sincos_cycles((double)r / (double)N, sn, cs);
        re = fma(wn, cs, re); im = fma(-wn, sn, im);
    }
    for (int o = 16; o > 0; o >>= 1) { re += __shfl_down_sync(0xffffffffu, re, o); im += __shfl_down_sync(0xffffffffu, im, o); }
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = re; s[1][threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0;
        for (int q = 0; q < WIN_THREADS / 32; q++) { t0 += s[0][q]; t1 += s[1][q]; }
        double *o = partial + ((long long)j * gridDim.x + blockIdx.x) * 2;
        o[0] = t0; o[1] = t1;
    }
}
// fixed-order sum of the partials: taps[j] (complex) for j = 0..H, taps[H + 1] = (sum w^2, 0)
__global__ void window_taps_final_kernel(const double *__restrict__ partial, int nchunk, double *__restrict__ taps) {
    __shared__ double s[2][256];
    const double *pp = partial + (long long)blockIdx.x * nchunk * 2;
    double a = 0, b = 0;
    for (int i = threadIdx.x; i < nchunk; i += 256) { a += pp[2 * i]; b += pp[2 * i + 1]; }
    s[0][threadIdx.x] = a; s[1][threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s[0][threadIdx.x] += s[0][threadIdx.x + o]; s[1][threadIdx.x] += s[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { taps[2 * blockIdx.x] = s[0][0]; taps[2 * blockIdx.x + 1] = s[1][0]; }
}
// out[c] = sum over blocks of partial[block][c] (fixed order), c = blockIdx.x < ncol
__global__ void inner_strided_sum_kernel(const double *__restrict__ partial, int nblk, int ncol, double *__restrict__ out) {
    __shared__ double s[256];
    double a = 0;
    for (int i = threadIdx.x; i < nblk; i += 256) a += partial[(long long)i * ncol + blockIdx.x];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}
// energy of a full-length DFT array split by circular distance from index 0: band[h] = sum over min(i, N - i) == h of |a_i|^2 for
// h = 0..H, band[H + 1] = everything farther out (one pass; per-block partials [H + 2])
__global__ void __launch_bounds__(WIN_THREADS) band_energy_kernel(const double2 *__restrict__ a, long long N, int H,
                                                                  double *__restrict__ partial) {
    extern __shared__ double sb[]; // [H + 2]
    for (int i = threadIdx.x; i < H + 2; i += WIN_THREADS) sb[i] = 0.0;
    __syncthreads();
    double far = 0.0;
    for (long long i = (long long)blockIdx.x * WIN_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * WIN_THREADS) {
        const double2 v = a[i];
        const double e = v.x * v.x + v.y * v.y;
        const long long d = i < N - i ? i : N - i;
        if (d <= H) atomicAdd(&sb[(int)d], e); else far += e;
    }
    for (int o = 16; o > 0; o >>= 1) far += __shfl_down_sync(0xffffffffu, far, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sb[H + 1], far);
    __syncthreads();
    for (int i = threadIdx.x; i < H + 2; i += WIN_THREADS) partial[(long long)blockIdx.x * (H + 2) + i] = sb[i];
}
// out[c][k - out_lo] = (1/N) sum_{i = -H..H} taps[i + H] b[c][(k - i) mod N] for k in [out_lo, out_lo + out_n): the signal slice
// (with its circular halo) is staged in shared memory once; grid (tiles of 1024 outputs, channels)
#define CONV_TILE 1024
__global__ void __launch_bounds__(WIN_THREADS) band_convolve_kernel(const double2 *__restrict__ taps, int H, const double2 *__restrict__ b,
                                                                    long long N, long long out_lo, long long out_n,
                                                                    double2 *__restrict__ out) {
    extern __shared__ double2 sc[]; // signal [CONV_TILE + 2 H] | taps [2 H + 1]
    double2 *st = sc + CONV_TILE + 2 * H;
    const double2 *bc = b + (long long)blockIdx.y * N;
    const long long k0 = out_lo + (long long)blockIdx.x * CONV_TILE; // first output of the tile
    for (int i = threadIdx.x; i < CONV_TILE + 2 * H; i += WIN_THREADS) {
        long long src = (k0 - H + i) % N;  // b index k0 - H + i, wrapped
        if (src < 0) src += N;
        sc[i] = bc[src];
    }
    for (int i = threadIdx.x; i < 2 * H + 1; i += WIN_THREADS) st[i] = taps[i];
    __syncthreads();
    const double inv = 1.0 / (double)N;
#pragma unroll
    for (int q = 0; q < CONV_TILE / WIN_THREADS; q++) {
        const int lk = q * WIN_THREADS + threadIdx.x;
        const long long k = k0 + lk;
        if (k >= out_lo + out_n) continue;
        double re = 0.0, im = 0.0;
        // b[k - i] sits at sc[lk + H - i]; i runs over -H..H in a fixed order
        for (int i = -H; i <= H; i++) {
            const double2 a = st[i + H], v = sc[lk + H - i];
            re = fma(a.x, v.x, fma(-a.y, v.y, re));
            im = fma(a.x, v.y, fma(a.y, v.x, im));
        }
        out[(long long)blockIdx.y * out_n + (k - out_lo)] = make_double2(re * inv, im * inv);
    }
}

__global__ void __launch_bounds__(256) fma_bench_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
}

// ==========================================================================================
// host side
// ==========================================================================================
static size_t spline_smem_bytes(int L, bool tiled) {
    return sizeof(double) * (size_t)L * (5 + (tiled ? 2 * SPL_ROWS : 0));
}

static size_t sum_smem_bytes(int L) {
    return sizeof(SubEntry) * SUM_RING * SUM_SUBCAP + sizeof(FillEntry) * SUM_ECAP + sizeof(RecC) * SUM_RCAP +
           sizeof(double) * SMEM_PER_KNOT * (size_t)L;
}

static int ensure_bytes(emrifd_handle *h, void **ptr, int64_t *cap, int64_t need, bool pinned_host = false) {
    if (need <= *cap) return 0;
    int64_t ncap = need + need / 4 + 256;
    if (*ptr) {
        if (pinned_host) { CUDA_TRY(h, cudaStreamSynchronize(h->stream)); CUDA_TRY(h, cudaFreeHost(*ptr)); }
        else { CUDA_TRY(h, cudaStreamSynchronize(h->stream)); CUDA_TRY(h, cudaFree(*ptr)); }
        *ptr = nullptr; *cap = 0;
    }
    if (pinned_host) CUDA_TRY(h, cudaMallocHost(ptr, (size_t)ncap));
    else CUDA_TRY(h, cudaMalloc(ptr, (size_t)ncap));
    *cap = ncap;
    return 0;
}

// upload walker descriptors (host -> device) through a pinned staging ring
static int upload_walkers(emrifd_handle *h, const emrifd_walker_t *walkers, int64_t B) {
    if (B > h->stage_cap) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 4; i++) {
            if (h->h_stage[i]) CUDA_TRY(h, cudaFreeHost(h->h_stage[i]));
            CUDA_TRY(h, cudaMallocHost((void **)&h->h_stage[i], sizeof(emrifd_walker_t) * (size_t)(B + B / 4 + 16)));
        }
        h->stage_cap = B + B / 4 + 16;
    }
    int64_t wc = h->walkers_cap;
    int rc = ensure_bytes(h, (void **)&h->d_walkers, &wc, (int64_t)sizeof(emrifd_walker_t) * B);
    if (rc) return rc;
    h->walkers_cap = wc;
    if ((int64_t)sizeof(int) * B > h->wstatus_cap) {
        if ((rc = ensure_bytes(h, (void **)&h->d_wstatus, &h->wstatus_cap, (int64_t)sizeof(int) * B))) return rc;
        CUDA_TRY(h, cudaMemsetAsync(h->d_wstatus, 0, (size_t)h->wstatus_cap, h->stream));
    }
    const int slot = h->stage_next;
    h->stage_next = (slot + 1) & 3;
    CUDA_TRY(h, cudaEventSynchronize(h->stage_ev[slot]));
    memcpy(h->h_stage[slot], walkers, sizeof(emrifd_walker_t) * (size_t)B);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_walkers, h->h_stage[slot], sizeof(emrifd_walker_t) * (size_t)B,
                                cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->stage_ev[slot], h->stream));
    return 0;
}

static int validate_walkers(emrifd_handle *h, const emrifd_walker_t *w, int64_t B, int *Lmax, int *Kmax) {
    if (!w || B <= 0 || B > 65535) return set_err(h, EMRIFD_ERR_INVALID, "walkers NULL or batch size outside [1, 65535]");
    int lm = 0, km = 0;
    int64_t nm = 0, nte = 0;
    for (int64_t i = 0; i < B; i++) {
        if (w[i].L < 4) return set_err(h, EMRIFD_ERR_TOO_FEW_KNOTS, "not-a-knot spline needs at least 4 knots");
        if (w[i].L > EMRIFD_MAX_KNOTS) return set_err(h, EMRIFD_ERR_TOO_MANY_KNOTS, "trajectory longer than EMRIFD_MAX_KNOTS");
        if (w[i].K < 1) return set_err(h, EMRIFD_ERR_INVALID, "walker with no modes");
        if (w[i].knot_off < 0 || w[i].teuk_off < 0 || w[i].mode_off < 0 || w[i].coeff_off < 0 || w[i].out_off < 0)
            return set_err(h, EMRIFD_ERR_INVALID, "walker descriptor with a negative offset");
        if (w[i].L > lm) lm = w[i].L;
        if (w[i].K > km) km = w[i].K;
        if (w[i].mode_off + w[i].K > nm) nm = w[i].mode_off + w[i].K;
        if (w[i].teuk_off + (int64_t)w[i].L * w[i].K > nte) nte = w[i].teuk_off + (int64_t)w[i].L * w[i].K;
    }
    *Lmax = lm; *Kmax = km;
    h->tot_modes = nm; h->tot_teuk = nte;
    return 0;
}

static int check_grid(emrifd_handle *h, int64_t N, double val, const double *fpos) {
    if (N < 3 || (N & 1) == 0) return set_err(h, EMRIFD_ERR_INVALID, "frequency grid length must be odd and >= 3");
    if (!fpos && !(val > 0.0)) return set_err(h, EMRIFD_ERR_INVALID, "implicit grid needs val = 1/(N dt) > 0");
    return 0;
}

extern "C" {

int emrifd_version(void) { return EMRIFD_VERSION; }
int emrifd_sizeof_branch(void) { return (int)sizeof(emrifd_branch_t); }
int emrifd_sizeof_walker(void) { return (int)sizeof(emrifd_walker_t); }

int emrifd_create(int device, void *stream, emrifd_handle_t **out) {
    if (!out) return EMRIFD_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return EMRIFD_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return EMRIFD_ERR_CUDA;
    emrifd_handle *h = new (std::nothrow) emrifd_handle();
    if (!h) return EMRIFD_ERR_NOMEM;
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->stream = (cudaStream_t)stream;
    if (cudaMalloc((void **)&h->d_status, sizeof(int)) != cudaSuccess) { delete h; return EMRIFD_ERR_CUDA; }
    cudaMemset(h->d_status, 0, sizeof(int));
    bool ok = true; // every set-up call is checked: a handle is either fully usable or not created
    for (int i = 0; i < 4; i++) ok &= cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming) == cudaSuccess;
    ok &= cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking) == cudaSuccess;
    { const char *ov = getenv("EMRIFD_OVERLAP"); h->overlap_mode = ov ? atoi(ov) : 1; }
    ok &= cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok &= cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 64; i++)
        ok &= cudaEventCreate(&h->ev_a[i]) == cudaSuccess && cudaEventCreate(&h->ev_b[i]) == cudaSuccess && cudaEventCreate(&h->ev_m[i]) == cudaSuccess &&
              cudaEventCreate(&h->ev_s[i]) == cudaSuccess;
    int optin = 0;
    ok &= cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess;
    const int big = optin - 4096; // static smem of the kernels (< 4 KB) comes out of the same budget
    h->max_dyn_smem = big;
    ok &= cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && h->num_sms > 0;
#define SET_ATTR(F_, BYTES_) ok &= cudaFuncSetAttribute(F_, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES_) == cudaSuccess
    SET_ATTR((spline_build_kernel<false, true>), big);
    SET_ATTR((spline_build_kernel<false, false>), big);
    SET_ATTR((spline_build_kernel<true, true>), big);
    SET_ATTR((spline_build_kernel<true, false>), big);
    SET_ATTR(segment_kernel, big);
    SET_ATTR(group_index_kernel, big);
    SET_ATTR(mode_select_kernel, SEL_CAP * 10);
#define SET_SMEM(W_, L_) \
    SET_ATTR((mode_sum_kernel<W_, L_, SUM_BPT, true>), big); SET_ATTR((mode_sum_kernel<W_, L_, SUM_BPT, false>), big);
    SET_SMEM(true, false) SET_SMEM(true, true) SET_SMEM(false, true)
#undef SET_SMEM
#undef SET_ATTR
    if (!ok || cudaGetLastError() != cudaSuccess) { emrifd_destroy(h); return EMRIFD_ERR_CUDA; }
    *out = h;
    return 0;
}

int emrifd_destroy(emrifd_handle_t *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_status); cudaFree(h->d_walkers); cudaFree(h->d_queue); cudaFree(h->d_partial); cudaFree(h->d_ws); cudaFree(h->d_chunk); cudaFree(h->d_tiledd);
    cudaFree(h->d_wstatus); cudaFree(h->d_leader); cudaFree(h->d_gcount); cudaFree(h->d_gq); cudaFree(h->d_gmem); cudaFree(h->d_goff); cudaFree(h->d_pieces); cudaFree(h->d_selovf);
    if (h->h_ws) cudaFreeHost(h->h_ws);
    for (int i = 0; i < 4; i++) { if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]); if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]); }
    for (int i = 0; i < 64; i++) { if (h->ev_a[i]) cudaEventDestroy(h->ev_a[i]); if (h->ev_b[i]) cudaEventDestroy(h->ev_b[i]); if (h->ev_m[i]) cudaEventDestroy(h->ev_m[i]); if (h->ev_s[i]) cudaEventDestroy(h->ev_s[i]); }
    if (h->aux) { cudaStreamSynchronize(h->aux); cudaStreamDestroy(h->aux); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    cudaGetLastError();
    delete h;
    return 0;
}

int emrifd_set_stream(emrifd_handle_t *h, void *stream) {
    if (!h) return EMRIFD_ERR_INVALID;
    h->stream = (cudaStream_t)stream;
    return 0;
}
int emrifd_synchronize(emrifd_handle_t *h) {
    if (!h) return EMRIFD_ERR_INVALID;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}
const char *emrifd_last_error(emrifd_handle_t *h) { return h ? h->err : "null handle"; }
int64_t emrifd_launch_count(emrifd_handle_t *h) { return h ? h->launches : 0; }

int emrifd_spline_build(emrifd_handle_t *h, const double *t, const double *y, int64_t L, int64_t R,
                        int64_t row_stride, int64_t knot_stride, double *coeff) {
    if (!h || !t || !y || !coeff || R <= 0) return set_err(h, EMRIFD_ERR_INVALID, "spline_build: bad argument");
    if (L < 4) return set_err(h, EMRIFD_ERR_TOO_FEW_KNOTS, "not-a-knot spline needs at least 4 knots");
    cudaSetDevice(h->device);
    SplineParams p; memset(&p, 0, sizeof(p));
    p.t = t; p.coeff = coeff; p.ygen = y; p.rs = row_stride; p.ks = knot_stride; p.Lgen = (int)L; p.Rgen = (int)R;
    dim3 grid((unsigned)((R + SPL_ROWS - 1) / SPL_ROWS), 1);
    const size_t tiled = spline_smem_bytes((int)L, true), plain = spline_smem_bytes((int)L, false);
    if ((int64_t)tiled <= h->max_dyn_smem) spline_build_kernel<true, true><<<grid, SPL_CTA, tiled, h->stream>>>(p, h->d_status);
    else if ((int64_t)plain <= h->max_dyn_smem) spline_build_kernel<true, false><<<grid, SPL_CTA, plain, h->stream>>>(p, h->d_status);
    else return set_err(h, EMRIFD_ERR_TOO_MANY_KNOTS, "spline_build: too many knots for the shared-memory factorisation");
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_spline_eval(emrifd_handle_t *h, const double *t, const double *coeff, int64_t L, int64_t R,
                       const double *tnew, int64_t n, double *out) {
    if (!h || !t || !coeff || !tnew || !out || L < 2 || R <= 0 || n < 0) return set_err(h, EMRIFD_ERR_INVALID, "spline_eval: bad argument");
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(R < 64 ? R : 64));
    spline_eval_kernel<<<grid, 256, 0, h->stream>>>(t, coeff, (int)L, (int)R, tnew, n, out);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

static int batch_spline_dev(emrifd_handle *h, int64_t B, int Lmax, int Kmax, const double *t, const double *teuk,
                            const double *f_phi, const double *f_r, const double *Phi_phi, const double *Phi_r, double *coeff) {
    SplineParams p; memset(&p, 0, sizeof(p));
    p.w = h->d_walkers; p.t = t; p.teuk = teuk; p.trk0 = f_phi; p.trk1 = f_r; p.trk2 = Phi_phi; p.trk3 = Phi_r; p.coeff = coeff;
    const int R = 2 * Kmax + 4;
    dim3 grid((unsigned)((R + SPL_ROWS - 1) / SPL_ROWS), (unsigned)B);
    const size_t tiled = spline_smem_bytes(Lmax, true), plain = spline_smem_bytes(Lmax, false);
    if ((int64_t)tiled <= h->max_dyn_smem) spline_build_kernel<false, true><<<grid, SPL_CTA, tiled, h->stream>>>(p, h->d_status);
    else spline_build_kernel<false, false><<<grid, SPL_CTA, plain, h->stream>>>(p, h->d_status);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_batch_spline(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                        const double *t, const double *teuk, const double *f_phi, const double *f_r,
                        const double *Phi_phi, const double *Phi_r, double *coeff) {
    if (!h || !t || !teuk || !f_phi || !f_r || !Phi_phi || !Phi_r || !coeff) return set_err(h, EMRIFD_ERR_INVALID, "batch_spline: NULL argument");
    int Lmax, Kmax, rc;
    if ((rc = validate_walkers(h, walkers, B, &Lmax, &Kmax))) return rc;
    cudaSetDevice(h->device);
    if ((rc = upload_walkers(h, walkers, B))) return rc;
    return batch_spline_dev(h, B, Lmax, Kmax, t, teuk, f_phi, f_r, Phi_phi, Phi_r, coeff);
}

static int batch_segment_dev(emrifd_handle *h, int64_t B, int Lmax, int Kmax, const double *t, const double *coeff,
                             const int32_t *m_arr, const int32_t *n_arr, int64_t N, double val, const double *fpos,
                             emrifd_branch_t *branches, int64_t *n_eval) {
    SegParams p;
    p.w = h->d_walkers; p.t = t; p.coeff = coeff; p.m = m_arr; p.n = n_arr; p.br = branches;
    p.n_eval = (long long *)n_eval;
    p.wstatus = h->d_wstatus;
    p.g.N = N; p.g.zero = (N - 1) / 2; p.g.val = val; p.g.fpos = fpos;
    CUDA_TRY(h, cudaMemsetAsync(h->d_wstatus, 0, sizeof(int) * (size_t)B, h->stream));
    if (n_eval) CUDA_TRY(h, cudaMemsetAsync(n_eval, 0, sizeof(int64_t) * 2 * (size_t)B, h->stream));
    // modes per CTA: as many as fit next to the staged tracks (8 at L <= ~600, fewer for very long trajectories)
    int mpc = (int)((h->max_dyn_smem - 72 * (int64_t)Lmax) / (36 * (int64_t)Lmax));
    mpc = mpc > 8 ? 8 : mpc;
    if (mpc < 1) return set_err(h, EMRIFD_ERR_TOO_MANY_KNOTS, "trajectory too long for the segmentation kernel");
    dim3 grid((unsigned)((Kmax + mpc - 1) / mpc), (unsigned)B);
    segment_kernel<<<grid, SEG_THREADS, (size_t)(72 + 36 * mpc) * (size_t)Lmax, h->stream>>>(p, h->d_status, mpc);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_batch_segment(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                         const double *t, const double *coeff, const int32_t *m_arr, const int32_t *n_arr,
                         int64_t N, double val, const double *fpos, emrifd_branch_t *branches, int64_t *n_eval) {
    if (!h || !t || !coeff || !m_arr || !n_arr || !branches) return set_err(h, EMRIFD_ERR_INVALID, "batch_segment: NULL argument");
    int Lmax, Kmax, rc;
    if ((rc = validate_walkers(h, walkers, B, &Lmax, &Kmax))) return rc;
    if ((rc = check_grid(h, N, val, fpos))) return rc;
    cudaSetDevice(h->device);
    if ((rc = upload_walkers(h, walkers, B))) return rc;
    return batch_segment_dev(h, B, Lmax, Kmax, t, coeff, m_arr, n_arr, N, val, fpos, branches, n_eval);
}

static int batch_sum_dev(emrifd_handle *h, int64_t B, int Lmax, int Kmax, const double *t, const double *coeff,
                         const int32_t *m_arr, const int32_t *n_arr, const double *ylm, const emrifd_branch_t *branches,
                         int64_t N, double val, const double *fpos, int flags, int64_t j_lo, int64_t j_cnt,
                         double *hp, double *hc, double *like_out, int64_t tile_first = 0, int64_t tile_stride = 1) {
    const bool write_h = hp && hc, like = like_out != nullptr;
    if (!write_h && !like) return set_err(h, EMRIFD_ERR_INVALID, "batch_sum: nothing to compute (no output requested)");
    if (like && !h->d_data) return set_err(h, EMRIFD_ERR_NO_DATA, "likelihood requested before emrifd_set_data");
    const int64_t npos = (N + 1) / 2;
    if (j_lo < 0 || j_cnt <= 0 || j_lo + j_cnt > npos) return set_err(h, EMRIFD_ERR_INVALID, "batch_sum: bin slice outside [0, (N+1)/2)");
    const bool mask_pos = (flags & EMRIFD_MASK_POSITIVE) != 0;
    if (write_h && !mask_pos && !(j_lo == 0 && j_cnt == npos))
        return set_err(h, EMRIFD_ERR_INVALID, "two-sided output needs the full bin range; use EMRIFD_MASK_POSITIVE for slices");
    if (like && h->n_data != npos) return set_err(h, EMRIFD_ERR_INVALID, "data length does not match (N+1)/2");
    SumParams p; memset(&p, 0, sizeof(p));
    p.w = h->d_walkers; p.t = t; p.coeff = coeff; p.m = m_arr; p.n = n_arr; p.ylm = (const double2 *)ylm; p.br = branches;
    p.g.N = N; p.g.zero = (N - 1) / 2; p.g.val = val; p.g.fpos = fpos;
    p.j_lo = j_lo; p.j_cnt = j_cnt;
    p.include_minus_m = (flags & EMRIFD_INCLUDE_MINUS_M) != 0; p.mask_positive = mask_pos;
    p.hp = (double2 *)hp; p.hc = (double2 *)hc;
    p.dw = (const double2 *)h->d_data; p.wf = h->d_wfac; p.n_data = h->n_data;
    // one tile shape for every launch, so that a walker's result does not depend on which batch it is evaluated in
    const int64_t tile_bins = SUM_TILE;
    const int64_t ntiles_all = (j_cnt + tile_bins - 1) / tile_bins;
    if (tile_stride < 1 || tile_first < 0) return set_err(h, EMRIFD_ERR_INVALID, "batch_sum: bad tile_first / tile_stride");
    if (tile_first >= ntiles_all) { // this rank owns no tile: all sums are zero
        if (like) CUDA_TRY(h, cudaMemsetAsync(like_out, 0, sizeof(double) * 3 * (size_t)B, h->stream));
        return 0;
    }
    const int64_t ntiles = (ntiles_all - tile_first + tile_stride - 1) / tile_stride; // tiles of this launch
    p.tile_first = tile_first; p.tile_stride = tile_stride;
    if (like) {
        int rc = ensure_bytes(h, (void **)&h->d_partial, &h->partial_cap, (int64_t)sizeof(double) * 3 * ntiles * SUM_CW * B);
        if (rc) return rc;
        p.partial = h->d_partial;
    }
    // (m, n) groups: index + combined amplitude quads (one stationary point per (group, bin) in the sum)
    {
        if (Kmax > GRP_TAB) return set_err(h, EMRIFD_ERR_INVALID, "batch_sum: more than 8192 modes in one walker");
        int rc;
        if ((rc = ensure_bytes(h, (void **)&h->d_leader, &h->leader_cap, (int64_t)sizeof(int) * h->tot_modes))) return rc;
        if ((rc = ensure_bytes(h, (void **)&h->d_gcount, &h->gcount_cap, (int64_t)sizeof(int) * B))) return rc;
        if ((rc = ensure_bytes(h, (void **)&h->d_gq, &h->gq_cap, (int64_t)sizeof(double) * 16 * h->tot_teuk))) return rc;
        if ((rc = ensure_bytes(h, (void **)&h->d_gmem, &h->gmem_cap, (int64_t)sizeof(int) * h->tot_modes))) return rc;
        if ((rc = ensure_bytes(h, (void **)&h->d_goff, &h->goff_cap, (int64_t)sizeof(int) * (h->tot_modes + B)))) return rc;
        GroupParams gp;
        gp.w = h->d_walkers; gp.coeff = coeff; gp.m = m_arr; gp.n = n_arr; gp.ylm = (const double2 *)ylm;
        gp.leader = h->d_leader; gp.gcount = h->d_gcount; gp.gmem = h->d_gmem; gp.goff = h->d_goff; gp.gq = h->d_gq;
        const size_t gsmem = sizeof(int) * ((size_t)GRP_TAB + 3 * (size_t)Kmax);
        group_index_kernel<<<(unsigned)B, GRP_THREADS, gsmem, h->stream>>>(gp);
        int64_t nsl = ((int64_t)Lmax * Kmax + 2047) / 2048; // CTAs per walker: ~2048 (knot, mode) pairs each
        nsl = nsl < 1 ? 1 : (nsl > 1024 ? 1024 : nsl);
        group_combine_kernel<<<dim3((unsigned)nsl, (unsigned)B), GRP_THREADS, 0, h->stream>>>(gp);
        h->launches += 2;
        CUDA_TRY(h, cudaGetLastError());
        p.leader = h->d_leader; p.gcount = h->d_gcount; p.gq = h->d_gq; p.wstatus = h->d_wstatus; p.k13_few = h->k13_few;
        // piece table: every (group record, segment) with its constants and inverse interpolant, once per batch
        if ((rc = ensure_bytes(h, &h->d_pieces, &h->pieces_cap, (int64_t)sizeof(Piece) * h->tot_modes * MAXBR * Lmax))) return rc;
        PieceParams pp;
        pp.w = h->d_walkers; pp.t = t; pp.coeff = coeff; pp.m = m_arr; pp.n = n_arr; pp.br = branches;
        pp.leader = h->d_leader; pp.gcount = h->d_gcount; pp.gq = h->d_gq; pp.pieces = (Piece *)h->d_pieces; pp.lstride = Lmax;
        int64_t npc = ((int64_t)Kmax * MAXBR * Lmax + PIECE_THREADS - 1) / PIECE_THREADS;
        npc = npc < 1 ? 1 : (npc > 4096 ? 4096 : npc);
        piece_build_kernel<<<dim3((unsigned)npc, (unsigned)B), PIECE_THREADS, 0, h->stream>>>(pp);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        p.pieces = (const Piece *)h->d_pieces; p.lstride = Lmax;
    }
    const int cpw = (Kmax * MAXBR + SUM_CHUNK - 1) / SUM_CHUNK;
    {
        int rc = ensure_bytes(h, (void **)&h->d_chunk, &h->chunk_cap, (int64_t)sizeof(long long) * 2 * cpw * B);
        if (rc) return rc;
        dim3 cgrid((unsigned)cpw, (unsigned)B);
        chunk_range_kernel<<<cgrid, SUM_CHUNK, 0, h->stream>>>(h->d_walkers, branches, h->d_leader, h->d_gcount, (N - 1) / 2, h->d_chunk, cpw);
        h->launches++;
        p.chunk_rng = h->d_chunk; p.cpw = cpw;
        // per-tile sum |d~|^2 table of the injected data: usable when this launch's tiles coincide with the table's
        // (a truncated last tile is sent through the per-bin path by the kernels, see tile_truncated)
        p.tile_dd = (like && (j_lo % tile_bins) == 0) ? h->d_tiledd : nullptr;
    }
    const size_t smem = sum_smem_bytes(Lmax);
    if ((int64_t)smem > h->max_dyn_smem) return set_err(h, EMRIFD_ERR_TOO_MANY_KNOTS, "trajectory too long for the shared-memory staging of the mode-sum kernel");
    dim3 grid((unsigned)ntiles, (unsigned)B);
    int ev = -1;
    if (h->timing && h->ev_n < 64) { ev = h->ev_n++; cudaEventRecord(h->ev_a[ev], h->stream); }
    // pass 1 (full grid, light): empty tiles are written here, the others are queued; pass 2: persistent CTAs drain the queue
    p.no_empty = (like && !p.tile_dd) ? 1 : 0;
    p.ntiles = (int)ntiles;
    {
        int rc = ensure_bytes(h, (void **)&h->d_queue, &h->queue_cap, (int64_t)(sizeof(unsigned long long) + 1) * ntiles * B + 32);
        if (rc) return rc;
        p.qctl = (unsigned int *)h->d_queue;
        p.queue = (unsigned long long *)h->d_queue + 2;
        p.tile_flag = (unsigned char *)(p.queue + ntiles * B);
        CUDA_TRY(h, cudaMemsetAsync(h->d_queue, 0, 16, h->stream));
    }
    // dispatch on (WRITE_H, LIKE)
#define SUM_DISPATCH(KERNEL, THREADS, GRID, SMEM, STREAM, ...)                                        \
    do {                                                                                              \
        if (write_h && like) KERNEL<true, true, SUM_BPT, ##__VA_ARGS__><<<GRID, THREADS, SMEM, STREAM>>>(p);  \
        else if (write_h) KERNEL<true, false, SUM_BPT, ##__VA_ARGS__><<<GRID, THREADS, SMEM, STREAM>>>(p);    \
        else KERNEL<false, true, SUM_BPT, ##__VA_ARGS__><<<GRID, THREADS, SMEM, STREAM>>>(p);                 \
    } while (0)
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mode_sum_kernel<true, true, SUM_BPT, true>, SUM_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t pgrid = (int64_t)h->num_sms * per_sm;
    const bool persistent = ntiles * B > 4 * pgrid; // (about one wave or less: a plain grid, every CTA classifies its own tile)
    {
        dim3 cgrid((unsigned)((ntiles + CLASSIFY_THREADS - 1) / CLASSIFY_THREADS), (unsigned)B);
        if (like) classify_tiles_kernel<true, SUM_BPT><<<cgrid, CLASSIFY_THREADS, 0, h->stream>>>(p);
        else classify_tiles_kernel<false, SUM_BPT><<<cgrid, CLASSIFY_THREADS, 0, h->stream>>>(p);
        h->launches++;
        if (persistent) {
            queue_build_kernel<<<1, QBUILD_THREADS, 0, h->stream>>>(p.tile_flag, (long long)ntiles * B, (int)ntiles, p.queue, p.qctl);
            h->launches++;
        }
    }
    // The zero-fill runs on a stream of its own underneath the sum (fork after the classification, join before the finalize);
    // the sum is launched first so that its CTAs take their SMs and the zero-fill CTAs fill in around them.
    // (EMRIFD_OVERLAP=0: same stream, after the sum -- for A/B runs.)
    const bool overlap = h->overlap_mode != 0;
    cudaStream_t zs = overlap ? h->aux : h->stream;
    if (overlap) CUDA_TRY(h, cudaEventRecord(h->ev_fork, h->stream));
    if (ev >= 0) cudaEventRecord(h->ev_m[ev], h->stream);
    if (persistent) SUM_DISPATCH(mode_sum_kernel, SUM_THREADS, (unsigned)pgrid, smem, h->stream, true);
    else SUM_DISPATCH(mode_sum_kernel, SUM_THREADS, grid, smem, h->stream, false);
    if (ev >= 0) cudaEventRecord(h->ev_s[ev], h->stream);
    if (overlap) CUDA_TRY(h, cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
    SUM_DISPATCH(empty_tile_kernel, EMPTY_THREADS, grid, 0, zs);
    if (overlap) {
        CUDA_TRY(h, cudaEventRecord(h->ev_join, h->aux));
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0)); // join: everything after this call sees both kernels' results
    }
#undef SUM_DISPATCH
    h->launches++;
    if (ev >= 0) cudaEventRecord(h->ev_b[ev], h->stream);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if (like) {
        like_finalize_kernel<<<(unsigned)B, FIN_THREADS, 0, h->stream>>>(h->d_partial, ntiles * SUM_CW, like_out, h->d_wstatus);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
    }
    return 0;
}

int emrifd_batch_sum(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                     const double *t, const double *coeff, const int32_t *m_arr, const int32_t *n_arr,
                     const double *ylm, const emrifd_branch_t *branches,
                     int64_t N, double val, const double *fpos, int flags, int64_t j_lo, int64_t j_cnt,
                     double *hp, double *hc, double *like_out) {
    if (!h || !t || !coeff || !m_arr || !n_arr || !ylm || !branches) return set_err(h, EMRIFD_ERR_INVALID, "batch_sum: NULL argument");
    int Lmax, Kmax, rc;
    if ((rc = validate_walkers(h, walkers, B, &Lmax, &Kmax))) return rc;
    if ((rc = check_grid(h, N, val, fpos))) return rc;
    cudaSetDevice(h->device);
    if ((rc = upload_walkers(h, walkers, B))) return rc;
    return batch_sum_dev(h, B, Lmax, Kmax, t, coeff, m_arr, n_arr, ylm, branches, N, val, fpos, flags, j_lo, j_cnt, hp, hc, like_out);
}

int emrifd_tile_bins(void) { return SUM_TILE; }

int emrifd_batch_sum_cyclic(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                            const double *t, const double *coeff, const int32_t *m_arr, const int32_t *n_arr,
                            const double *ylm, const emrifd_branch_t *branches,
                            int64_t N, double val, const double *fpos, int flags, int64_t tile_first, int64_t tile_stride,
                            double *hp, double *hc, double *like_out) {
    if (!h || !t || !coeff || !m_arr || !n_arr || !ylm || !branches) return set_err(h, EMRIFD_ERR_INVALID, "batch_sum_cyclic: NULL argument");
    if ((hp || hc) && !(flags & EMRIFD_MASK_POSITIVE))
        return set_err(h, EMRIFD_ERR_INVALID, "batch_sum_cyclic: waveform output needs EMRIFD_MASK_POSITIVE");
    int Lmax, Kmax, rc;
    if ((rc = validate_walkers(h, walkers, B, &Lmax, &Kmax))) return rc;
    if ((rc = check_grid(h, N, val, fpos))) return rc;
    cudaSetDevice(h->device);
    if ((rc = upload_walkers(h, walkers, B))) return rc;
    return batch_sum_dev(h, B, Lmax, Kmax, t, coeff, m_arr, n_arr, ylm, branches, N, val, fpos, flags, 0, (N + 1) / 2, hp, hc, like_out,
                         tile_first, tile_stride); // ownership is defined on tiles of emrifd_tile_bins() bins
}

int emrifd_fd_waveform_batch(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                             const double *t, const double *teuk, const double *f_phi, const double *f_r,
                             const double *Phi_phi, const double *Phi_r,
                             const int32_t *m_arr, const int32_t *n_arr, const double *ylm,
                             int64_t N, double val, const double *fpos, int flags,
                             double *coeff, emrifd_branch_t *branches, double *hp, double *hc, double *like_out) {
    if (!h || !t || !teuk || !f_phi || !f_r || !Phi_phi || !Phi_r || !m_arr || !n_arr || !ylm || !coeff || !branches)
        return set_err(h, EMRIFD_ERR_INVALID, "fd_waveform_batch: NULL argument");
    int Lmax, Kmax, rc;
    if ((rc = validate_walkers(h, walkers, B, &Lmax, &Kmax))) return rc;
    if ((rc = check_grid(h, N, val, fpos))) return rc;
    cudaSetDevice(h->device);
    if ((rc = upload_walkers(h, walkers, B))) return rc;
    if ((rc = batch_spline_dev(h, B, Lmax, Kmax, t, teuk, f_phi, f_r, Phi_phi, Phi_r, coeff))) return rc;
    if ((rc = batch_segment_dev(h, B, Lmax, Kmax, t, coeff, m_arr, n_arr, N, val, fpos, branches, nullptr))) return rc;
    return batch_sum_dev(h, B, Lmax, Kmax, t, coeff, m_arr, n_arr, ylm, branches, N, val, fpos, flags, 0, (N + 1) / 2, hp, hc, like_out);
}

int emrifd_batch_status(emrifd_handle_t *h) {
    if (!h) return EMRIFD_ERR_INVALID;
    cudaSetDevice(h->device);
    int st = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&st, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (st != 0) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream));
        if (st == EMRIFD_ERR_KNOT_ORDER) return set_err(h, st, "trajectory knots are not strictly increasing");
        if (st == EMRIFD_ERR_BRANCHES) return set_err(h, st, "a mode has more monotone branches than EMRIFD_MAX_BRANCHES");
        return set_err(h, st, "device-side error");
    }
    return 0;
}

int emrifd_set_data(emrifd_handle_t *h, const double *d_whitened, const double *noise_factor, int64_t n) {
    if (!h || !d_whitened || !noise_factor || n <= 0) return set_err(h, EMRIFD_ERR_INVALID, "set_data: bad argument");
    h->d_data = d_whitened; h->d_wfac = noise_factor; h->n_data = n;
    cudaSetDevice(h->device);
    const int tile = SUM_TILE;
    const int64_t nt = (n + tile - 1) / tile;
    int rc = ensure_bytes(h, (void **)&h->d_tiledd, &h->tiledd_cap, (int64_t)sizeof(double) * nt);
    if (rc) return rc;
    tile_dd_kernel<<<(unsigned)nt, 256, 0, h->stream>>>((const double2 *)d_whitened, n, tile, h->d_tiledd);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_inner_product(emrifd_handle_t *h, const double *a, const double *b, int64_t nch, int64_t n,
                         const double *freqs, const double *psd, double *out) {
    if (!h || !a || !b || !freqs || !out || nch <= 0 || n < 2) return set_err(h, EMRIFD_ERR_INVALID, "inner_product: bad argument");
    cudaSetDevice(h->device);
    int64_t nb = (nch * n + 256 * 8 - 1) / (256 * 8);
    if (nb > (int64_t)h->num_sms * 8) nb = (int64_t)h->num_sms * 8; // 8 resident CTAs per SM
    if (nb < 1) nb = 1;
    int rc = ensure_bytes(h, (void **)&h->d_partial, &h->partial_cap, (int64_t)sizeof(double) * 2 * nb);
    if (rc) return rc;
    inner_partial_kernel<<<(unsigned)nb, 256, 0, h->stream>>>((const double2 *)a, (const double2 *)b, nch, n, freqs, psd, h->d_partial);
    inner_final_kernel<<<1, 256, 0, h->stream>>>(h->d_partial, (int)nb, out);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_loglike(emrifd_handle_t *h, const double *templates, int64_t B, double *out) {
    if (!h || !templates || !out || B <= 0 || B > 65535) return set_err(h, EMRIFD_ERR_INVALID, "loglike: bad argument");
    if (!h->d_data) return set_err(h, EMRIFD_ERR_NO_DATA, "likelihood requested before emrifd_set_data");
    cudaSetDevice(h->device);
    const int64_t n = h->n_data;
    int64_t nb = (2 * n + 256 * 4 - 1) / (256 * 4);
    const int64_t wave = (int64_t)h->num_sms * 8;
    const int64_t cap = (wave + B - 1) / B > 8 ? (wave + B - 1) / B : 8;
    if (nb > cap) nb = cap;
    int rc = ensure_bytes(h, (void **)&h->d_partial, &h->partial_cap, (int64_t)sizeof(double) * 3 * nb * B);
    if (rc) return rc;
    dim3 grid((unsigned)nb, (unsigned)B);
    loglike_partial_kernel<<<grid, 256, 0, h->stream>>>((const double2 *)templates, (const double2 *)h->d_data, h->d_wfac, n, h->d_partial);
    like_finalize_kernel<<<(unsigned)B, FIN_THREADS, 0, h->stream>>>(h->d_partial, nb, out, nullptr);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_loglike_batch_host(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                              const double *t, const double *teuk, const double *f_phi, const double *f_r,
                              const double *Phi_phi, const double *Phi_r,
                              const int32_t *m_arr, const int32_t *n_arr, const double *ylm,
                              int64_t N, double val, const double *fpos_dev, int flags,
                              double *hp_dev, double *hc_dev, double *like_out_host) {
    if (!h || !t || !teuk || !f_phi || !f_r || !Phi_phi || !Phi_r || !m_arr || !n_arr || !ylm || !like_out_host)
        return set_err(h, EMRIFD_ERR_INVALID, "loglike_batch_host: NULL argument");
    int Lmax, Kmax, rc;
    if ((rc = validate_walkers(h, walkers, B, &Lmax, &Kmax))) return rc;
    if ((rc = check_grid(h, N, val, fpos_dev))) return rc;
    if (!h->d_data) return set_err(h, EMRIFD_ERR_NO_DATA, "likelihood requested before emrifd_set_data");
    cudaSetDevice(h->device);
    // sizes of the packed inputs (walkers are packed back to back in the order given)
    int64_t nk = 0, nm = 0, nte = 0, nco = 0;
    for (int64_t i = 0; i < B; i++) {
        const int64_t L = walkers[i].L, K = walkers[i].K;
        if (walkers[i].knot_off + L > nk) nk = walkers[i].knot_off + L;
        if (walkers[i].mode_off + K > nm) nm = walkers[i].mode_off + K;
        if (walkers[i].teuk_off + L * K > nte) nte = walkers[i].teuk_off + L * K;
        if (walkers[i].coeff_off + L * (2 * K + 4) * 4 > nco) nco = walkers[i].coeff_off + L * (2 * K + 4) * 4;
    }
    // one packed staging buffer: [t f_phi f_r Phi_phi Phi_r | teuk | ylm | m n] + result
    auto al = [](int64_t x) { return (x + 255) & ~(int64_t)255; };
    const int64_t o_t = 0, o_fp = o_t + al(8 * nk), o_fr = o_fp + al(8 * nk), o_pp = o_fr + al(8 * nk), o_pr = o_pp + al(8 * nk);
    const int64_t o_te = o_pr + al(8 * nk), o_y = o_te + al(16 * nte), o_m = o_y + al(16 * 2 * nm), o_n = o_m + al(4 * nm);
    const int64_t in_bytes = o_n + al(4 * nm);
    const int64_t o_co = in_bytes, o_br = o_co + al(8 * nco), o_out = o_br + al((int64_t)sizeof(emrifd_branch_t) * nm * MAXBR);
    const int64_t dev_bytes = o_out + al(8 * 3 * B);
    if ((rc = ensure_bytes(h, (void **)&h->d_ws, &h->ws_cap, dev_bytes))) return rc;
    if ((rc = ensure_bytes(h, (void **)&h->h_ws, &h->h_ws_cap, in_bytes + al(8 * 3 * B), true))) return rc;
    char *hs = h->h_ws;
    // Staging into pinned memory and the H2D copy, pipelined: the small arrays first, then the amplitudes (the bulk: 3.4 MB of the
    // bench batch's 4 MB) in slices -- a slice's DMA runs while the next one is being staged, and a few host threads share the
    // staging copies (one thread moves ~10 GB/s out of pageable memory: the copy would cost as much as 6 % of the step).
    memcpy(hs + o_t, t, 8 * nk); memcpy(hs + o_fp, f_phi, 8 * nk); memcpy(hs + o_fr, f_r, 8 * nk);
    memcpy(hs + o_pp, Phi_phi, 8 * nk); memcpy(hs + o_pr, Phi_r, 8 * nk);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_ws + o_t, hs + o_t, (size_t)(o_te - o_t), cudaMemcpyHostToDevice, h->stream));
    {
        const int64_t te_bytes = 16 * nte;
        const int nsl = te_bytes > (1 << 20) ? 4 : 1;
        const int64_t per = ((te_bytes + nsl - 1) / nsl + 4095) & ~(int64_t)4095;
        for (int sl = 0; sl < nsl; sl++) {
            const int64_t lo = sl * per, hi = lo + per < te_bytes ? lo + per : te_bytes;
            if (hi <= lo) break;
            const int nth = (hi - lo) > (256 << 10) ? 4 : 1;
            const int64_t part = ((hi - lo + nth - 1) / nth + 63) & ~(int64_t)63;
#pragma omp parallel for num_threads(nth) schedule(static)
            for (int q = 0; q < nth; q++) {
                const int64_t a = lo + q * part, b = a + part < hi ? a + part : hi;
                if (b > a) memcpy(hs + o_te + a, (const char *)teuk + a, (size_t)(b - a));
            }
            CUDA_TRY(h, cudaMemcpyAsync(h->d_ws + o_te + lo, hs + o_te + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, h->stream));
        }
    }
    memcpy(hs + o_y, ylm, 16 * 2 * nm);
    memcpy(hs + o_m, m_arr, 4 * nm); memcpy(hs + o_n, n_arr, 4 * nm);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_ws + o_y, hs + o_y, (size_t)(in_bytes - o_y), cudaMemcpyHostToDevice, h->stream));
    if ((rc = upload_walkers(h, walkers, B))) return rc;
    char *d = h->d_ws;
    double *coeff = (double *)(d + o_co);
    emrifd_branch_t *br = (emrifd_branch_t *)(d + o_br);
    double *dout = (double *)(d + o_out);
    if ((rc = batch_spline_dev(h, B, Lmax, Kmax, (double *)(d + o_t), (double *)(d + o_te), (double *)(d + o_fp), (double *)(d + o_fr),
                               (double *)(d + o_pp), (double *)(d + o_pr), coeff))) return rc;
    if ((rc = batch_segment_dev(h, B, Lmax, Kmax, (double *)(d + o_t), coeff, (int32_t *)(d + o_m), (int32_t *)(d + o_n), N, val, fpos_dev, br, nullptr))) return rc;
    if ((rc = batch_sum_dev(h, B, Lmax, Kmax, (double *)(d + o_t), coeff, (int32_t *)(d + o_m), (int32_t *)(d + o_n), (double *)(d + o_y), br,
                            N, val, fpos_dev, flags | EMRIFD_MASK_POSITIVE, 0, (N + 1) / 2, hp_dev, hc_dev, dout))) return rc;
    double *hres = (double *)(hs + in_bytes);
    CUDA_TRY(h, cudaMemcpyAsync(hres, dout, sizeof(double) * 3 * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    int st = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&st, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    memcpy(like_out_host, hres, sizeof(double) * 3 * (size_t)B);
    if (st != 0) {
        // data-dependent failure of some walker(s): their rows are NaN (the per-walker contract of the reference's samplers,
        // Eryn/eryn/moves/red_blue.py:282-284); the rest of the batch is valid.  emrifd_walker_status tells which and why.
        cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream);
        set_err(h, st, st == EMRIFD_ERR_BRANCHES ? "a mode has more monotone branches than EMRIFD_MAX_BRANCHES (walker reported as NaN)"
                                                 : "trajectory knots are not strictly increasing (walker reported as NaN)");
    }
    return 0;
}

int emrifd_walker_status(emrifd_handle_t *h, int64_t B, int32_t *status_host) {
    if (!h || !status_host || B <= 0) return set_err(h, EMRIFD_ERR_INVALID, "walker_status: bad argument");
    if ((int64_t)sizeof(int) * B > h->wstatus_cap) return set_err(h, EMRIFD_ERR_INVALID, "walker_status: no batch of that size has run on this handle");
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaMemcpyAsync(status_host, h->d_wstatus, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int emrifd_walker_status_dev(emrifd_handle_t *h, int64_t B, int32_t *status_dev) {
    if (!h || !status_dev || B <= 0) return set_err(h, EMRIFD_ERR_INVALID, "walker_status_dev: bad argument");
    if ((int64_t)sizeof(int) * B > h->wstatus_cap) return set_err(h, EMRIFD_ERR_INVALID, "walker_status_dev: no batch of that size has run on this handle");
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaMemcpyAsync(status_dev, h->d_wstatus, sizeof(int) * (size_t)B, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

int emrifd_set_overlap(emrifd_handle_t *h, int enable) {
    if (!h) return EMRIFD_ERR_INVALID;
    h->overlap_mode = enable ? 1 : 0;
    return 0;
}

int emrifd_set_k13_mode(emrifd_handle_t *h, int mode) {
    if (!h || (mode != EMRIFD_K13_EXACT && mode != EMRIFD_K13_FEW)) return set_err(h, EMRIFD_ERR_INVALID, "set_k13_mode: mode must be EMRIFD_K13_EXACT or EMRIFD_K13_FEW");
    h->k13_few = mode == EMRIFD_K13_FEW;
    return 0;
}

int emrifd_mode_select(emrifd_handle_t *h, const double *teuk, int64_t nsamp, int64_t M, const int32_t *samp_walker,
                       const double *ylm, const int32_t *neg_src, int64_t Mneg, int64_t B, double eps, uint8_t *flags) {
    if (!h || !teuk || !samp_walker || !ylm || !flags || nsamp <= 0 || M <= 0 || Mneg < 0 || B <= 0 || (Mneg > 0 && !neg_src))
        return set_err(h, EMRIFD_ERR_INVALID, "mode_select: bad argument");
    if (M + Mneg > SEL_CAP) return set_err(h, EMRIFD_ERR_INVALID, "mode_select: more than 8192 modes (incl. -m copies)");
    if (!(eps >= 0.0 && eps < 1.0)) return set_err(h, EMRIFD_ERR_INVALID, "mode_select: eps must be in [0, 1)");
    cudaSetDevice(h->device);
    SelParams p;
    p.teuk = (const double2 *)teuk; p.samp_walker = samp_walker; p.ylm = (const double2 *)ylm; p.neg_src = neg_src;
    p.M = (int)M; p.Mneg = (int)Mneg; p.B = (int)B; p.eps = eps; p.flags = flags;
    CUDA_TRY(h, cudaMemsetAsync(flags, 0, (size_t)(B * M), h->stream));
    // Two launches: a small sort capacity first (10 KB of shared memory instead of 80: eight resident CTAs per SM instead of
    // two -- at eps = 1e-2 a few hundred modes survive the pre-filter), then the full capacity for the samples that overflowed
    // (every other CTA of that launch exits at once).  No host round trip, identical results.
    int rc = ensure_bytes(h, (void **)&h->d_selovf, &h->selovf_cap, nsamp);
    if (rc) return rc;
    p.ovf = (unsigned char *)h->d_selovf;
    auto launch = [&](int pass, int cap) {
        p.pass = pass; p.cap = cap;
        mode_select_kernel<<<(unsigned)nsamp, SEL_THREADS, (size_t)cap * (sizeof(double) + sizeof(unsigned short)), h->stream>>>(p);
        h->launches++;
    };
    if (M + Mneg <= SEL_CAP_SMALL) launch(0, SEL_CAP_SMALL);   // cannot overflow
    else if (eps < 1e-3) launch(0, SEL_CAP);                    // most modes survive the pre-filter: full capacity straight away
    else { launch(0, SEL_CAP_SMALL); launch(1, SEL_CAP); }
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_ylm_batch(emrifd_handle_t *h, const int32_t *l_arr, const int32_t *m_arr, int64_t M, const int32_t *neg_src,
                     int64_t Mneg, const double *theta, const double *phi, int64_t B, int lmax, double *ylm_out) {
    if (!h || !l_arr || !m_arr || !theta || !phi || !ylm_out || M <= 0 || Mneg < 0 || B <= 0 || (Mneg > 0 && !neg_src))
        return set_err(h, EMRIFD_ERR_INVALID, "ylm_batch: bad argument");
    if (lmax < 2 || lmax > YLM_LMAX) return set_err(h, EMRIFD_ERR_INVALID, "ylm_batch: lmax must be in [2, 12]");
    cudaSetDevice(h->device);
    ylm_kernel<<<(unsigned)B, 256, 0, h->stream>>>(l_arr, m_arr, (int)M, neg_src, (int)Mneg, theta, phi, (double2 *)ylm_out);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_synth_amplitude(emrifd_handle_t *h, const double *p, const double *e, int64_t nsamp, const int32_t *l_arr,
                           const int32_t *m_arr, const int32_t *n_arr, const double *cmode, int64_t M, int lmax, int nmax,
                           double *teuk_out) {
    if (!h || !p || !e || !l_arr || !m_arr || !n_arr || !cmode || !teuk_out || nsamp <= 0 || M <= 0)
        return set_err(h, EMRIFD_ERR_INVALID, "synth_amplitude: bad argument");
    if (lmax < 2 || lmax > AMP_LMAX || nmax < 0 || nmax > AMP_NMAX) return set_err(h, EMRIFD_ERR_INVALID, "synth_amplitude: lmax <= 12, nmax <= 32");
    cudaSetDevice(h->device);
    synth_amplitude_kernel<<<(unsigned)nsamp, 256, 0, h->stream>>>(p, e, l_arr, m_arr, n_arr, (const double2 *)cmode, (int)M, lmax, nmax,
                                                                    (double2 *)teuk_out);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_mode_compact_count(emrifd_handle_t *h, const uint8_t *flags, int64_t B, int64_t M, int32_t *keep_idx, int32_t *K_out) {
    if (!h || !flags || !keep_idx || !K_out || B <= 0 || M <= 0) return set_err(h, EMRIFD_ERR_INVALID, "mode_compact_count: bad argument");
    cudaSetDevice(h->device);
    compact_count_kernel<<<(unsigned)B, 256, 0, h->stream>>>(flags, (int)M, keep_idx, K_out);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_mode_compact_gather(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B, const double *teuk_full, int64_t M,
                               int64_t Mneg, const int32_t *keep_idx, const int32_t *m_basis, const int32_t *n_basis,
                               const int32_t *neg_pos, const double *ylm_full, double *teuk_out, int32_t *m_out, int32_t *n_out,
                               double *ylm_out) {
    if (!h || !teuk_full || !keep_idx || !m_basis || !n_basis || !neg_pos || !ylm_full || !teuk_out || !m_out || !n_out || !ylm_out ||
        M <= 0 || Mneg < 0)
        return set_err(h, EMRIFD_ERR_INVALID, "mode_compact_gather: bad argument");
    int Lmax, Kmax, rc;
    if ((rc = validate_walkers(h, walkers, B, &Lmax, &Kmax))) return rc;
    if (Kmax > M) return set_err(h, EMRIFD_ERR_INVALID, "mode_compact_gather: walker K exceeds the mode basis");
    cudaSetDevice(h->device);
    if ((rc = upload_walkers(h, walkers, B))) return rc;
    compact_gather_kernel<<<dim3((unsigned)(Lmax + 1), (unsigned)B), 256, 0, h->stream>>>(
        h->d_walkers, Lmax, (int)M, (int)Mneg, (const double2 *)teuk_full, keep_idx, m_basis, n_basis, neg_pos,
        (const double2 *)ylm_full, (double2 *)teuk_out, m_out, n_out, (double2 *)ylm_out);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_window_taps(emrifd_handle_t *h, const double *window, int64_t N, int H, double *taps) {
    if (!h || !window || !taps || N < 2 || H < 0 || H > EMRIFD_WINDOW_MAX_TAPS || 2 * (int64_t)H + 1 > N)
        return set_err(h, EMRIFD_ERR_INVALID, "window_taps: bad argument (0 <= H <= EMRIFD_WINDOW_MAX_TAPS, 2 H + 1 <= N)");
    cudaSetDevice(h->device);
    int64_t nchunk = (N + WIN_THREADS * 16 - 1) / (WIN_THREADS * 16);
    const int64_t cap = (int64_t)h->num_sms * 8 / (H + 2) + 1;
    if (nchunk > cap) nchunk = cap;
    if (nchunk < 1) nchunk = 1;
    int rc = ensure_bytes(h, (void **)&h->d_partial, &h->partial_cap, (int64_t)sizeof(double) * 2 * nchunk * (H + 2));
    if (rc) return rc;
    window_taps_kernel<<<dim3((unsigned)nchunk, (unsigned)(H + 2)), WIN_THREADS, 0, h->stream>>>(window, N, H, h->d_partial);
    window_taps_final_kernel<<<(unsigned)(H + 2), 256, 0, h->stream>>>(h->d_partial, (int)nchunk, taps);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_band_energy(emrifd_handle_t *h, const double *a, int64_t N, int H, double *energy) {
    if (!h || !a || !energy || N < 2 || H < 0 || H > EMRIFD_WINDOW_MAX_TAPS || 2 * (int64_t)H + 1 > N)
        return set_err(h, EMRIFD_ERR_INVALID, "band_energy: bad argument");
    cudaSetDevice(h->device);
    int64_t nb = (N + WIN_THREADS * 16 - 1) / (WIN_THREADS * 16);
    if (nb > (int64_t)h->num_sms * 4) nb = (int64_t)h->num_sms * 4;
    if (nb < 1) nb = 1;
    int rc = ensure_bytes(h, (void **)&h->d_partial, &h->partial_cap, (int64_t)sizeof(double) * (nb * (H + 2) + 2 * (H + 2)));
    if (rc) return rc;
    band_energy_kernel<<<(unsigned)nb, WIN_THREADS, sizeof(double) * (H + 2), h->stream>>>((const double2 *)a, N, H, h->d_partial);
    // fixed-order sum over the blocks: reuse the taps finaliser on a transposed view is not possible (layout [block][H + 2]), so
    // one small kernel launch per call would be needed; the table is tiny -- sum it with the generic strided finaliser below
    inner_strided_sum_kernel<<<(unsigned)(H + 2), 256, 0, h->stream>>>(h->d_partial, (int)nb, H + 2, energy);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_band_convolve(emrifd_handle_t *h, const double *taps, int H, const double *signal, int64_t nch, int64_t N,
                         int64_t out_lo, int64_t out_n, double *out) {
    if (!h || !taps || !signal || !out || nch <= 0 || nch > 65535 || N < 2 || H < 0 || H > EMRIFD_WINDOW_MAX_TAPS || 2 * (int64_t)H + 1 > N ||
        out_lo < 0 || out_n <= 0 || out_lo + out_n > N)
        return set_err(h, EMRIFD_ERR_INVALID, "band_convolve: bad argument");
    cudaSetDevice(h->device);
    const size_t smem = sizeof(double2) * (size_t)(CONV_TILE + 4 * H + 1);
    dim3 grid((unsigned)((out_n + CONV_TILE - 1) / CONV_TILE), (unsigned)nch);
    band_convolve_kernel<<<grid, WIN_THREADS, smem, h->stream>>>((const double2 *)taps, H, (const double2 *)signal, N, out_lo, out_n, (double2 *)out);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int emrifd_bench_fp64_fma(emrifd_handle_t *h, int iters, double *gflops) {
    if (!h || !gflops || iters <= 0) return set_err(h, EMRIFD_ERR_INVALID, "bench_fp64_fma: bad argument");
    cudaSetDevice(h->device);
    double *d = nullptr;
    CUDA_TRY(h, cudaMalloc((void **)&d, 8));
    const int blocks = h->num_sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fma_bench_kernel<<<blocks, 256, 0, h->stream>>>(d, 16, 1.0000001, 1e-9); // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a, h->stream);
        fma_bench_kernel<<<blocks, 256, 0, h->stream>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(b, h->stream);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    h->launches += 6;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    CUDA_TRY(h, cudaGetLastError());
    const double flops = 2.0 * 64.0 * (double)iters * 256.0 * blocks;
    *gflops = flops / (best * 1e-3) * 1e-9;
    return 0;
}

#ifdef SUM_STATS
int emrifd_debug_stats(uint64_t *out) { // debug build only: read and clear the counters
    unsigned long long z[16] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_stats, sizeof(z));
    cudaMemcpyToSymbol(g_stats, z, sizeof(z));
    return 0;
}
#endif

int emrifd_sum_kernel_times(emrifd_handle_t *h, int enable, double *ms_pair, double *ms_mode_sum, int64_t *launches) {
    if (!h) return EMRIFD_ERR_INVALID;
    cudaSetDevice(h->device);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < h->ev_n; i++) {
        float e = 0, m = 0;
        if (cudaEventElapsedTime(&e, h->ev_a[i], h->ev_b[i]) == cudaSuccess && cudaEventElapsedTime(&m, h->ev_m[i], h->ev_s[i]) == cudaSuccess) {
            h->sum_ms += e; h->sum_ms_main += m; h->sum_launches++;
        }
    }
    h->ev_n = 0;
    if (ms_pair) *ms_pair = h->sum_ms;
    if (ms_mode_sum) *ms_mode_sum = h->sum_ms_main;
    if (launches) *launches = h->sum_launches;
    if (ms_pair || ms_mode_sum || launches) { h->sum_ms = 0; h->sum_ms_main = 0; h->sum_launches = 0; }
    h->timing = enable;
    return 0;
}

int emrifd_sum_kernel_time(emrifd_handle_t *h, int enable, double *ms, int64_t *launches) {
    return emrifd_sum_kernel_times(h, enable, ms, nullptr, launches);
}

} // extern "C"

/*
 * emrifd_oracle.c -- CPU ORACLE for the FD EMRI mode-sum + likelihood hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (emri_frequencydomainwaveforms_b200/) never imports, links or calls anything here.
 *
 * PARITY STATUS: "parity unpinned" against FastEMRIWaveforms itself.  The arithmetic of
 * this path lives in the third-party dependency `few` (FastEMRIWaveforms, v1.5.x API,
 * unpinned, not vendored, not installable here).  This file
 * restates the published algorithm from the in-repo statements of it:
 *   - per-harmonic SPA construction:  Tutorial_FD_construction_single_mode.ipynb cell 26
 *     (JSON lines 548-623): arg = -2*pi*i*fdot^3/(3*fddot^2),
 *     amp = A*Ylm * i*fdot/|fddot| * K_{1/3}(arg)*exp(arg) * 2/sqrt(3),
 *     h(+f) = amp*exp(i(2 pi f t - Phi_mn)),  h(-f) = conj(A)*Y_{l,-m}*[fdot,fddot -> -fdot,-fddot]*exp(i(-2 pi f t + Phi_mn))
 *   - not-a-knot spline algebra: SciPy CubicSpline (the notebook uses it interchangeably
 *     with few's CubicSplineInterpolant, cells 8/11/20) -- pinned against SciPy in tests
 *   - sign/flip + h+/hx split:  SURVEY.md A.3 (notebook cell 25/28-32 for the -f, -Re, -Im convention)
 *   - inner product:  LISAanalysistools/lisatools/diagnostic.py:95-110
 *   - likelihood:     LISAanalysistools/lisatools/sampling/likelihood.py:178-180,213-220,257-274
 * It is pinned against: SciPy CubicSpline (spline), scipy.special.kv + mpmath (K_{1/3} factor),
 * a NumPy transcription of lisatools inner_product / Likelihood.get_ll executed from
 * /root/reference (tests/golden/make_golden.py), and first-principles identities.
 *
 * Structure follows the reference formulation (mode-major SCATTER into W[], then a
 * flip/split pass) -- deliberately different from the product's bin-owner GATHER kernel.
 *
 * Compile twice:  -DORC_QUAD  -> evaluation in __float128 ("truth")
 *                 (default)   -> evaluation in double (FEW-equivalent CPU baseline, OpenMP)
 * Spline build and segmentation are ALWAYS plain double with no FMA contraction
 * (compile with -ffp-contract=off): they define bit-exact index sets.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <quadmath.h>
typedef __float128 qreal; /* the K_{1/3} ascending series always runs in binary128 (cancellation) */
#ifdef ORC_QUAD
typedef __float128 real;
#define R_(x) x##Q
#define r_sqrt sqrtq
#define r_cbrt cbrtq
#define r_fabs fabsq
#define r_sin sinq
#define r_cos cosq
#define ASYM_TERMS 24
#else
typedef double real;
#define R_(x) x
#define r_sqrt sqrt
#define r_cbrt cbrt
#define r_fabs fabs
#define r_sin sin
#define r_cos cos
#define ASYM_TERMS 16
#endif

#define SERIES_XMAX 30.0
#define ORC_MAXBR 4
#define ORC_PI R_(3.14159265358979323846264338327950288419716939937510)

typedef struct {
    int32_t mode, dir, ja, jb;
    int32_t closed_end, pad;
    int64_t start, end;
    double xa, xb, Fa, Fb;
} orc_branch_t;

int orc_is_quad(void) {
#ifdef ORC_QUAD
    return 1;
#else
    return 0;
#endif
}
int orc_sizeof_branch(void) { return (int)sizeof(orc_branch_t); }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* the timed CPU-baseline legs use every host thread even when the launcher (torchrun) exported OMP_NUM_THREADS=1 */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ */
/* A3: not-a-knot cubic spline, SciPy CubicSpline algebra              */
/* y is [R][L] row-major; coeff out is [L][R][4] = (y, c1, c2, c3)      */
/* ------------------------------------------------------------------ */
int orc_spline_build(const double *t, const double *y, int L, int R, double *coeff) {
    /* Frozen operation order (the CUDA kernel reproduces it bit for bit):
         rh_j = 1/h_j;  delta_j = (y_{j+1}-y_j)*rh_j;
         LU without pivoting: inv_0 = 1/d_0;  w_i = a_i*inv_{i-1};  d'_i = d_i - w_i*c_{i-1};  inv_i = 1/d'_i;
         forward  b'_i = b_i - w_i*b'_{i-1};  back  s_i = (b'_i - c_i*s_{i+1})*inv_i;
         tau = (s_i + s_{i+1} - 2 delta_i)*rh_i;  c2 = (delta_i - s_i)*rh_i - tau;  c3 = tau*rh_i. */
    if (L < 4) return -2;
    double *h = (double *)malloc(sizeof(double) * L * 5);
    double *w = h + L, *inv = w + L, *cup = inv + L, *rh = cup + L;
    for (int j = 0; j < L - 1; j++) {
        h[j] = t[j + 1] - t[j];
        if (!(h[j] > 0.0)) { free(h); return -3; }
        rh[j] = 1.0 / h[j];
    }
    const double dd0 = t[2] - t[0], ddn = t[L - 1] - t[L - 3];
    double dprev = h[1]; /* d_0 */
    cup[0] = dd0;
    inv[0] = 1.0 / dprev;
    w[0] = 0.0;
    for (int i = 1; i < L; i++) {
        double a, d;
        if (i < L - 1) { a = h[i]; d = 2.0 * (h[i - 1] + h[i]); cup[i] = h[i - 1]; }
        else           { a = ddn;  d = h[L - 3];                 cup[i] = 0.0; }
        w[i] = a * inv[i - 1];
        dprev = d - w[i] * cup[i - 1];
        inv[i] = 1.0 / dprev;
    }
#pragma omp parallel for schedule(static)
    for (int r = 0; r < R; r++) {
        const double *yr = y + (size_t)r * L;
        double *s = (double *)malloc(sizeof(double) * L);
        double dm = (yr[1] - yr[0]) * rh[0], dp = (yr[2] - yr[1]) * rh[1];
        s[0] = ((h[0] + 2.0 * dd0) * h[1] * dm + h[0] * h[0] * dp) / dd0;
        for (int i = 1; i < L - 1; i++) {
            dp = (yr[i + 1] - yr[i]) * rh[i];
            double b = 3.0 * (h[i] * dm + h[i - 1] * dp);
            s[i] = b - w[i] * s[i - 1];
            if (i < L - 2) dm = dp; /* at the end: dm = delta_{L-3}, dp = delta_{L-2} */
        }
        {
            double b = (h[L - 2] * h[L - 2] * dm + (2.0 * ddn + h[L - 2]) * h[L - 3] * dp) / ddn;
            s[L - 1] = b - w[L - 1] * s[L - 2];
        }
        s[L - 1] = s[L - 1] * inv[L - 1];
        for (int i = L - 2; i >= 0; i--) s[i] = (s[i] - cup[i] * s[i + 1]) * inv[i];
        for (int i = 0; i < L - 1; i++) {
            double dl = (yr[i + 1] - yr[i]) * rh[i];
            double tau = (s[i] + s[i + 1] - 2.0 * dl) * rh[i];
            double *c = coeff + ((size_t)i * R + r) * 4;
            c[0] = yr[i];
            c[1] = s[i];
            c[2] = (dl - s[i]) * rh[i] - tau;
            c[3] = tau * rh[i];
        }
        double *c = coeff + ((size_t)(L - 1) * R + r) * 4;
        c[0] = yr[L - 1]; c[1] = s[L - 1]; c[2] = 0.0; c[3] = 0.0;
        free(s);
    }
    free(h);
    return 0;
}

/* spline evaluation with end-segment extrapolation (SciPy extrapolate=True semantics) */
int orc_spline_eval(const double *t, const double *coeff, int L, int R,
                    const double *tnew, int64_t n, double *out /* [R][n] */) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < n; q++) {
        double tq = tnew[q];
        int lo = 0, hi = L - 2;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (t[mid] <= tq) lo = mid; else hi = mid - 1; }
        double x = tq - t[lo];
        for (int r = 0; r < R; r++) {
            const double *c = coeff + ((size_t)lo * R + r) * 4;
            out[(size_t)r * n + q] = c[0] + x * (c[1] + x * (c[2] + x * c[3]));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* frequency grid helpers: symmetric odd grid, zero at (N-1)/2         */
/* implicit: f_i = (i-zero)*val ; explicit: fpos[j], j=0..(N-1)/2      */
/* ------------------------------------------------------------------ */
typedef struct { int64_t N, zero; double val; const double *fpos; } orc_grid_t;

static inline double grid_f(const orc_grid_t *g, int64_t i) {
    int64_t k = i - g->zero;
    if (g->fpos) return k >= 0 ? g->fpos[k] : -g->fpos[-k];
    return (double)k * g->val;
}
/* smallest i in [0,N] with f_i >= F (strict=0) or f_i > F (strict=1) */
static int64_t grid_lower(const orc_grid_t *g, double F, int strict) {
    double df = g->fpos ? (g->fpos[g->zero] / (double)g->zero) : g->val;
    double e = F / df + (double)g->zero;
    int64_t i;
    if (!(e > 0.0)) i = 0; else if (e >= (double)g->N) i = g->N; else i = (int64_t)e;
#define COND(ii) (strict ? (grid_f(g, (ii)) > F) : (grid_f(g, (ii)) >= F))
    while (i > 0 && COND(i - 1)) i--;
    while (i < g->N && !COND(i)) i++;
#undef COND
    return i;
}

/* ------------------------------------------------------------------ */
/* A4: per-mode monotone-branch segmentation (plain double, no FMA)    */
/* coeff [L][R][4], rows: ReA[0..K), ImA[K..2K), f_phi, f_r, Phi_phi, Phi_r */
/* out: branches [K][ORC_MAXBR]; nbr[K]; returns 0 or -4 on overflow   */
/* ------------------------------------------------------------------ */
static void push_sub(orc_branch_t *br, int *nb, int *overflow, int k, int j,
                     double xa, double Fa, double xb, double Fb) {
    int dir = (Fb > Fa) - (Fb < Fa);
    if (dir == 0 || !(xb > xa)) return;
    if (*nb > 0 && br[*nb - 1].dir == dir) {
        br[*nb - 1].jb = j; br[*nb - 1].xb = xb; br[*nb - 1].Fb = Fb;
        return;
    }
    if (*nb >= ORC_MAXBR) { *overflow = 1; return; }
    orc_branch_t *b = &br[*nb];
    b->mode = k; b->dir = dir; b->ja = j; b->jb = j; b->closed_end = 0; b->pad = 0;
    b->xa = xa; b->xb = xb; b->Fa = Fa; b->Fb = Fb; b->start = 0; b->end = -1;
    (*nb)++;
}

int orc_segment_build(const double *t, const double *coeff, int L, int K,
                      const int32_t *m_arr, const int32_t *n_arr,
                      int64_t N, double val, const double *fpos,
                      orc_branch_t *branches, int32_t *nbr) {
    const int R = 2 * K + 4;
    orc_grid_t g = { N, (N - 1) / 2, val, fpos };
    int any_overflow = 0;
#pragma omp parallel for schedule(static) reduction(| : any_overflow)
    for (int k = 0; k < K; k++) {
        const double dm = (double)m_arr[k], dn = (double)n_arr[k];
        orc_branch_t *br = branches + (size_t)k * ORC_MAXBR;
        for (int q = 0; q < ORC_MAXBR; q++) {
            memset(&br[q], 0, sizeof(orc_branch_t));
            br[q].mode = k; br[q].start = 0; br[q].end = -1;
        }
        int nb = 0, overflow = 0;
        for (int j = 0; j < L - 1; j++) {
            const double *cp = coeff + ((size_t)j * R + 2 * K) * 4;     /* f_phi quad */
            const double *cr = cp + 4;                                   /* f_r quad  */
            const double *cp1 = coeff + ((size_t)(j + 1) * R + 2 * K) * 4;
            const double hj = t[j + 1] - t[j];
            const double c0 = dm * cp[0] + dn * cr[0];
            const double c1 = dm * cp[1] + dn * cr[1];
            const double c2 = dm * cp[2] + dn * cr[2];
            const double c3 = dm * cp[3] + dn * cr[3];
            const double Fnext = dm * cp1[0] + dn * cp1[4];
            /* roots of fdot(x) = c1 + 2 c2 x + 3 c3 x^2 in (0,hj) */
            double xr[2]; int nr = 0;
            const double qa = 3.0 * c3, qb = 2.0 * c2, qc = c1;
            if (qa == 0.0) {
                if (qb != 0.0) { double r0 = -qc / qb; if (r0 > 0.0 && r0 < hj) xr[nr++] = r0; }
            } else {
                double disc = qb * qb - 4.0 * qa * qc;
                if (disc >= 0.0) {
                    double sq = sqrt(disc);
                    double qq = (qb >= 0.0) ? -0.5 * (qb + sq) : -0.5 * (qb - sq);
                    double r0 = qq / qa;
                    double r1 = (qq != 0.0) ? qc / qq : r0;
                    if (r0 > r1) { double tmp = r0; r0 = r1; r1 = tmp; }
                    if (r0 > 0.0 && r0 < hj) xr[nr++] = r0;
                    if (r1 > 0.0 && r1 < hj && r1 != r0) xr[nr++] = r1;
                }
            }
            double xa = 0.0, Fa = c0;
            for (int q = 0; q < nr; q++) {
                double x = xr[q];
                double Fx = c0 + x * (c1 + x * (c2 + x * c3));
                push_sub(br, &nb, &overflow, k, j, xa, Fa, x, Fx);
                xa = x; Fa = Fx;
            }
            push_sub(br, &nb, &overflow, k, j, xa, Fa, hj, Fnext);
        }
        if (nb > 0) br[nb - 1].closed_end = 1;
        for (int q = 0; q < nb; q++) {
            orc_branch_t *b = &br[q];
            /* time-start end inclusive, time-end exclusive unless closed_end */
            double Flo = b->dir > 0 ? b->Fa : b->Fb, Fhi = b->dir > 0 ? b->Fb : b->Fa;
            int lo_strict = (b->dir > 0) ? 0 : !b->closed_end;
            int hi_strict = (b->dir > 0) ? !b->closed_end : 0; /* strict: f < Fhi */
            int64_t s = grid_lower(&g, Flo, lo_strict);
            /* end = (smallest i with f_i > Fhi [or >= Fhi if strict]) - 1 */
            int64_t e = grid_lower(&g, Fhi, hi_strict ? 0 : 1) - 1;
            b->start = s; b->end = e;
        }
        nbr[k] = nb;
        any_overflow |= overflow;
    }
    return any_overflow ? -4 : 0;
}

/* ------------------------------------------------------------------ */
/* SPA factor: R(X) = K_{1/3}(-iX) e^{-iX} sqrt(2X/pi) e^{-i pi/4}      */
/* and S(X) = R(X)/X^{1/6} for the small-X (turnover) regime            */
/* ------------------------------------------------------------------ */
/* K_{1/3} evaluation mode: 0 = math-exact (default); 1 = FastEMRIWaveforms-compatible ([UPSTREAM-MEMORY]):
   14-term ascending series (to X^26) for |X| <= 7 and the 9-term asymptotic series above -- 2.5e-7 off at the seam. */
static int g_k13_few = 0;
void orc_set_k13_mode(int few) { g_k13_few = few != 0; }
int orc_get_k13_mode(void) { return g_k13_few; }

static void k13_series_S(real Xr, real *Sre, real *Sim) {
    /* S = sqrt(2pi/3) e^{-i pi/4} e^{-iX} [2^{1/3} e^{i pi/6} B - 2^{-1/3} X^{2/3} e^{-i pi/6} A]
       A = sum (-X^2/4)^k/(k! Gamma(k+4/3)),  B = sum (-X^2/4)^k/(k! Gamma(k+2/3))
       (K_nu = (pi/2)(I_{-nu} - I_nu)/sin(nu pi), I_nu ascending series, z = -iX) */
    const qreal X = (qreal)Xr;
    const qreal PIq = 3.14159265358979323846264338327950288419716939937510Q;
    const qreal G43 = 0.892979511569249211234577932735323975Q; /* Gamma(4/3) */
    const qreal G23 = 1.354117939426400416945288028154513785Q; /* Gamma(2/3) */
    qreal q = -X * X / 4.0Q;
    qreal ta = 1.0Q / G43, tb = 1.0Q / G23, A = ta, B = tb;
    const int kmax = g_k13_few ? 14 : 400; /* FEW mode: exactly 14 terms of each series */
    for (int k = 1; k < kmax; k++) {
        ta *= q / ((qreal)k * ((qreal)k + 1.0Q / 3.0Q));
        tb *= q / ((qreal)k * ((qreal)k - 1.0Q / 3.0Q));
        A += ta; B += tb;
        if (!g_k13_few && k > (int)(Xr) + 2 && fabsq(ta) < 1e-40Q * fabsq(A) && fabsq(tb) < 1e-40Q * fabsq(B)) break;
    }
    qreal c13 = cbrtq(2.0Q);
    qreal x23 = cbrtq(X); x23 *= x23;
    qreal cb = c13 * B, ca = x23 * A / c13;
    const qreal s3h = sqrtq(3.0Q) / 2.0Q; /* e^{+-i pi/6} = (sqrt3/2, +-1/2) */
    qreal ure = s3h * cb - s3h * ca;
    qreal uim = 0.5Q * cb + 0.5Q * ca;
    qreal ang = -(X + PIq / 4.0Q);
    qreal cs = cosq(ang), sn = sinq(ang);
    qreal pref = sqrtq(2.0Q * PIq / 3.0Q);
    *Sre = (real)(pref * (ure * cs - uim * sn));
    *Sim = (real)(pref * (ure * sn + uim * cs));
}

static void k13_asym_R(real X, real *Rre, real *Rim) {
    /* R ~ sum a_k (i/X)^k, a_k = a_{k-1} (4/9 - (2k-1)^2)/(8k) */
    real a = R_(1.0), u = R_(1.0) / X, p = R_(1.0);
    real re = R_(1.0), im = R_(0.0);
    real last = R_(1.0);
    const int nterms = g_k13_few ? 9 : ASYM_TERMS; /* FEW mode: a_0 .. a_8, no optimal truncation */
    for (int k = 1; k < nterms; k++) {
        a *= (R_(4.0) / R_(9.0) - (real)((2 * k - 1) * (2 * k - 1))) / (R_(8.0) * (real)k);
        p *= u;
        real term = a * p;
        if (!g_k13_few && r_fabs(term) > r_fabs(last)) break; /* optimal truncation */
        last = term;
        switch (k & 3) {
            case 0: re += term; break;
            case 1: im += term; break;
            case 2: re -= term; break;
            case 3: im -= term; break;
        }
    }
    *Rre = re; *Rim = im;
}

/* returns R(X) for X >= 1 regime and S(X) = R/X^{1/6} always; which!=0 => small branch used */
void orc_spa_R(double Xd, double *Rre, double *Rim) {
    real X = (real)Xd, re, im;
    if (X < (g_k13_few ? (real)7.0 : (real)SERIES_XMAX)) {
        k13_series_S(X, &re, &im);
        real x16 = r_sqrt(r_cbrt(X));
        re *= x16; im *= x16;
    }
    else k13_asym_R(X, &re, &im);
    *Rre = (double)re; *Rim = (double)im;
}
void orc_spa_S(double Xd, double *Sre, double *Sim) {
    real re, im;
    k13_series_S((real)Xd, &re, &im);
    *Sre = (double)re; *Sim = (double)im;
}

/* G(fdot, fddot) = i fdot/|fddot| (2/sqrt3) K_{1/3}(-iX) e^{-iX}, X = 2 pi fdot^3/(3 fddot^2) */
static void spa_G(real fdot, real fddot, real *Gre, real *Gim) {
    real af = r_fabs(fdot), re, im;
    const real r2 = R_(0.70710678118654752440084436210484903928); /* 1/sqrt2 */
    if (fddot == R_(0.0)) { re = R_(1.0) / r_sqrt(af); im = R_(0.0); }
    else {
        real X = R_(2.0) * ORC_PI * af * af * af / (R_(3.0) * fddot * fddot);
        if (X < R_(1.0)) {
            /* G0 = e^{i3pi/4} S(X) (2pi/(3 fddot^2))^{1/6} : finite as fdot -> 0 */
            k13_series_S(X, &re, &im);
            real sc = r_cbrt(r_sqrt(R_(2.0) * ORC_PI / R_(3.0)) / r_fabs(fddot));
            re *= sc; im *= sc;
        } else {
            if (X < (g_k13_few ? (real)7.0 : (real)SERIES_XMAX)) {
                k13_series_S(X, &re, &im);
                real x16 = r_sqrt(r_cbrt(X)); re *= x16; im *= x16;
            }
            else k13_asym_R(X, &re, &im);
            real sc = R_(1.0) / r_sqrt(af);
            re *= sc; im *= sc;
        }
    }
    /* times e^{i 3pi/4} = (-1/sqrt2, 1/sqrt2) */
    real gre = (-re - im) * r2, gim = (re - im) * r2;
    if (fdot < R_(0.0)) gim = -gim; /* G(-fdot,-fddot) = conj G */
    *Gre = gre; *Gim = gim;
}

/* ------------------------------------------------------------------ */
/* A5+A6+A7: mode sum (scatter into W), flip/split, scale, rotate       */
/* ------------------------------------------------------------------ */
typedef struct { real re, im; } cplx;

int orc_mode_sum(const double *t, const double *coeff, int L, int K,
                 const int32_t *m_arr, const int32_t *n_arr,
                 const double *ylm /* [2K] complex interleaved: +m block then -m block */,
                 int64_t N, double val, const double *fpos,
                 const orc_branch_t *branches, const int32_t *nbr,
                 int include_minus_m, double scale, double cos2psi, double sin2psi,
                 int64_t out_lo, int64_t out_n,   /* output slice of full-grid indices [out_lo, out_lo+out_n) */
                 double *hp /* [out_n] complex interleaved */, double *hc,
                 int64_t *n_eval /* optional: number of (mode,bin) root evaluations */) {
    const int R = 2 * K + 4;
    orc_grid_t g = { N, (N - 1) / 2, val, fpos };
    cplx *W = (cplx *)calloc((size_t)N, sizeof(cplx));
    if (!W) return -5;
    int64_t evals = 0;
    const real twopi = R_(2.0) * ORC_PI;
    /* mode-major scatter, as in the reference formulation.  Serial over modes so that the
       accumulation order into W is deterministic; OpenMP parallelism is over bins inside a branch
       (distinct bins of one branch never collide unless the branch straddles f=0, see below). */
    for (int k = 0; k < K; k++) {
        const real dm = (real)m_arr[k], dn = (real)n_arr[k];
        const int mirror = (m_arr[k] > 0) && include_minus_m;
        const real ypr = (real)ylm[2 * k], ypi = (real)ylm[2 * k + 1];
        const real ymr = (real)ylm[2 * (K + k)], ymi = (real)ylm[2 * (K + k) + 1];
        for (int q = 0; q < nbr[k]; q++) {
            const orc_branch_t *b = &branches[(size_t)k * ORC_MAXBR + q];
            int64_t s = b->start < 0 ? 0 : b->start, e = b->end > N - 1 ? N - 1 : b->end;
            if (e < s) continue;
            evals += (e - s + 1);
            /* a branch that straddles f=0 has +f and mirrored -f targets inside one range: run it serially */
            const int cross = (s <= g.zero && g.zero <= e);
#pragma omp parallel for schedule(static) if (!cross)
            for (int64_t i = s; i <= e; i++) {
                const double f = grid_f(&g, i);
                /* locate the segment inside the branch (knot frequencies in double, as in segmentation) */
                int lo = b->ja, hi = b->jb;
                while (lo < hi) {
                    int mid = (lo + hi + 1) >> 1;
                    const double *cq = coeff + ((size_t)mid * R + 2 * K) * 4;
                    double Fk = (double)m_arr[k] * cq[0] + (double)n_arr[k] * cq[4];
                    int ok = b->dir > 0 ? (Fk <= f) : (Fk >= f);
                    if (ok) lo = mid; else hi = mid - 1;
                }
                const int j = lo;
                const double hj = t[j + 1] - t[j];
                const real xlo0 = (j == b->ja) ? (real)b->xa : R_(0.0);
                const real xhi0 = (j == b->jb) ? (real)b->xb : (real)hj;
                const double *cA = coeff + ((size_t)j * R + k) * 4;
                const double *cB = coeff + ((size_t)j * R + K + k) * 4;
                const double *cp = coeff + ((size_t)j * R + 2 * K) * 4;
                const double *cr = cp + 4, *cP = cp + 8, *cR = cp + 12;
                const real c0 = dm * (real)cp[0] + dn * (real)cr[0];
                const real c1 = dm * (real)cp[1] + dn * (real)cr[1];
                const real c2 = dm * (real)cp[2] + dn * (real)cr[2];
                const real c3 = dm * (real)cp[3] + dn * (real)cr[3];
                const real delta = (real)f - c0;
                /* bracketed Newton on g(x) = c1 x + c2 x^2 + c3 x^3 - delta, monotone with sign dir */
                real xl = xlo0, xh = xhi0;
                real gl = xl * (c1 + xl * (c2 + xl * c3)) - delta;
                real gh = xh * (c1 + xh * (c2 + xh * c3)) - delta;
                real x;
                if (gh == gl) x = R_(0.5) * (xl + xh); else x = xl - gl * (xh - xl) / (gh - gl);
                if (!(x >= xl)) x = xl;
                if (!(x <= xh)) x = xh;
                const real sdir = (real)b->dir;
                for (int it = 0; it < 200; it++) {
                    real gx = x * (c1 + x * (c2 + x * c3)) - delta;
                    real dg = c1 + x * (R_(2.0) * c2 + R_(3.0) * c3 * x);
                    if (gx * sdir > R_(0.0)) xh = x; else xl = x;
                    real xn;
                    if (dg != R_(0.0)) xn = x - gx / dg; else xn = R_(0.5) * (xl + xh);
                    if (!(xn >= xl && xn <= xh)) xn = R_(0.5) * (xl + xh);
                    real dx = r_fabs(xn - x);
                    x = xn;
#ifdef ORC_QUAD
                    if (dx <= R_(1e-28) * (real)hj) break;
#else
                    if (dx <= 4e-16 * hj) break;
#endif
                }
                /* evaluate splines */
                const real ReA = (real)cA[0] + x * ((real)cA[1] + x * ((real)cA[2] + x * (real)cA[3]));
                const real ImA = (real)cB[0] + x * ((real)cB[1] + x * ((real)cB[2] + x * (real)cB[3]));
                const real Pp = (real)cP[0] + x * ((real)cP[1] + x * ((real)cP[2] + x * (real)cP[3]));
                const real Pr = (real)cR[0] + x * ((real)cR[1] + x * ((real)cR[2] + x * (real)cR[3]));
                const real fdot = c1 + x * (R_(2.0) * c2 + R_(3.0) * c3 * x);
                const real fddot = R_(2.0) * c2 + R_(6.0) * c3 * x;
                real Gre, Gim;
                spa_G(fdot, fddot, &Gre, &Gim);
                const real tstar = (real)t[j] + x;
                real phase = twopi * (real)f * tstar - (dm * Pp + dn * Pr);
                const real cs = r_cos(phase), sn = r_sin(phase);
                /* C = A * G * e^{i phase} */
                const real agr = ReA * Gre - ImA * Gim, agi = ReA * Gim + ImA * Gre;
                const real Cr = agr * cs - agi * sn, Ci = agr * sn + agi * cs;
                W[i].re += ypr * Cr - ypi * Ci;
                W[i].im += ypr * Ci + ypi * Cr;
                if (mirror) {
                    const int64_t im_ = N - 1 - i; /* == i at f=0: both terms land on the same bin */
                    W[im_].re += ymr * Cr + ymi * Ci; /* Y_{l,-m} * conj(C) */
                    W[im_].im += ymi * Cr - ymr * Ci;
                }
            }
        }
    }
    /* A6/A7: S = -flip(W); h+ = (S + conj(flip S))/2; hx = i (S - conj(flip S))/2; scale; rotate */
    const real sc = (real)scale, c2p = (real)cos2psi, s2p = (real)sin2psi;
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < out_n; o++) {
        int64_t i = out_lo + o;
        cplx wi = W[i], wm = W[N - 1 - i];
        /* S_i = -W[N-1-i];  S_{N-1-i} = -W[i] */
        real pr = R_(0.5) * (-wm.re - wi.re), pi_ = R_(0.5) * (-wm.im + wi.im);
        /* hx = i/2 (S_i - conj(S_mirror)) = i/2 ((-wm.re + wi.re) + i(-wm.im - wi.im)) */
        real xr = R_(0.5) * (wm.im + wi.im), xi = R_(0.5) * (-wm.re + wi.re);
        real opr = sc * (c2p * pr - s2p * xr), opi = sc * (c2p * pi_ - s2p * xi);
        real oxr = sc * (s2p * pr + c2p * xr), oxi = sc * (s2p * pi_ + c2p * xi);
        hp[2 * o] = (double)opr; hp[2 * o + 1] = (double)opi;
        hc[2 * o] = (double)oxr; hc[2 * o + 1] = (double)oxi;
    }
    free(W);
    if (n_eval) *n_eval = evals;
    return 0;
}

/* ------------------------------------------------------------------ */
/* A10: lisatools inner_product (diagnostic.py:95-110)                  */
/* sums 4 * sum_k dx_k * Re(conj(a) b)/S over nch channels               */
/* ------------------------------------------------------------------ */
double orc_inner_product(const double *a, const double *b, int nch, int64_t n,
                         const double *freqs, const double *psd /* may be NULL => 1 */) {
    real out = R_(0.0);
    for (int ch = 0; ch < nch; ch++) {
        const double *ac = a + (size_t)ch * n * 2, *bc = b + (size_t)ch * n * 2;
        real s = R_(0.0);
        for (int64_t k = 0; k < n; k++) {
            double dx = (k == 0) ? (freqs[1] - freqs[0]) : (freqs[k] - freqs[k - 1]);
            real re = (real)ac[2 * k] * (real)bc[2 * k] + (real)ac[2 * k + 1] * (real)bc[2 * k + 1];
            real y = re / (psd ? (real)psd[k] : R_(1.0));
            s += (real)dx * y;
        }
        out += R_(4.0) * s;
    }
    return (double)out;
}

/* A11: ll = -1/2 * 4 * sum_ch sum_k |dw - h*w|^2   (likelihood.py:257-274);
   also returns <d|h>-like and <h|h>-like whitened sums for diagnostics */
int orc_loglike(const double *dw /* [nch][n] complex, whitened */, const double *h /* [nch][n] complex */,
                const double *wfac /* [nch][n] */, int nch, int64_t n, double *out3) {
    real s = R_(0.0), sdh = R_(0.0), shh = R_(0.0);
    for (int ch = 0; ch < nch; ch++)
        for (int64_t k = 0; k < n; k++) {
            size_t o = ((size_t)ch * n + k);
            real hr = (real)h[2 * o] * (real)wfac[o], hi = (real)h[2 * o + 1] * (real)wfac[o];
            real dr = (real)dw[2 * o] - hr, di = (real)dw[2 * o + 1] - hi;
            s += dr * dr + di * di;
            sdh += (real)dw[2 * o] * hr + (real)dw[2 * o + 1] * hi;
            shh += hr * hr + hi * hi;
        }
    out3[0] = (double)(-R_(0.5) * R_(4.0) * s);
    out3[1] = (double)(R_(4.0) * sdh);
    out3[2] = (double)(R_(4.0) * shh);
    return 0;
}

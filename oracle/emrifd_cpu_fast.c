/*
 * emrifd_cpu_fast.c -- an OPTIMISED double-precision CPU implementation of the FD mode sum + likelihood,
 * used ONLY as the timed CPU baseline (bench.py `cpu_baseline` and `--impl reference`).
 *
 * TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE: the product path never loads it.  The checker stays
 * emrifd_oracle.c (binary128, cold bracketed Newton, serial over modes); this file exists because timing a checker
 * says nothing about a CPU implementation.  It computes the same quantities from the same inputs -- spline
 * coefficients and work-list of the oracle's orc_spline_build / orc_segment_build (bit-exact, cheap), then
 *   - one stationary point per ((m, n) group, bin): amplitude quads of a group's (l, m, n) members combined first
 *     (Tutorial_FD_construction_single_mode.ipynb:558-616: t*, SPA factor and phase depend on (m, n) only),
 *   - OpenMP over tiles of positive-frequency bins, each tile owned by one thread (no atomics, private accumulators),
 *   - along a branch the root of bin i+1 is Newton-iterated from the root of bin i (warm start), cold bracketed solve
 *     only on the first bin of a (tile, branch, segment),
 *   - K_{1/3} factor from the same double-precision tables as the CUDA kernel (k13_tables.h), phase in cycles with an exact
 *     f * t_j two-product, polynomial sincos, FMA everywhere (compile with -O3 -march=native -ffp-contract=fast),
 *   - flip / h+,hx split / scale / rotation and the |d - h|^2, <d|h>, <h|h> sums fused into the tile read-out
 *     (LISAanalysistools/lisatools/sampling/likelihood.py:257-274).
 * Validated against the binary128 oracle to <= 1e-9 of max|h| in tests/test_oracle_cpu.py::test_fast_cpu_baseline_matches_oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define __device__
#define __constant__
#include "../emri_frequencydomainwaveforms_b200/csrc/k13_tables.h" /* generated tables only (no code) */

#define MAXBR 4
#define TILE 2048

typedef struct {
    int32_t mode, dir, ja, jb;
    int32_t closed_end, pad;
    int64_t start, end;
    double xa, xb, Fa, Fb;
} branch_t;

static const double c_sin[8] = {6.283185307179586, -41.34170224039976, 81.60524927607506, -76.70585975306139,
                                42.058693944897655, -15.09464257682299, 3.819952584848282, -0.7181223017785006};
static const double c_cos[9] = {1.0, -19.739208802178716, 64.9393940226683, -85.45681720669373, 60.24464137187666,
                                -26.4262567833744, 7.903536371318469, -1.714390711088672, 0.28200596845579123};

static inline void sincos_cycles(double c, double *sn, double *cs) {
    const double q = rint(4.0 * c);
    const int qi = (int)q;
    const double r = fma(-0.25, q, c), r2 = r * r;
    double ps = c_sin[7], pc = c_cos[8];
    for (int k = 6; k >= 0; k--) { ps = fma(ps, r2, c_sin[k]); pc = fma(pc, r2, c_cos[k + 1]); }
    ps *= r;
    pc = fma(pc, r2, c_cos[0]);
    const double a = (qi & 1) ? pc : ps, b = (qi & 1) ? ps : pc;
    *sn = (qi & 2) ? -a : a;
    *cs = ((qi + 1) & 2) ? -b : b;
}

/* R(X)/sqrt|fdot| with R = K_{1/3}(-iX) e^{-iX} sqrt(2X/pi) e^{-i pi/4}; u = 1/X, s = 1/sqrt|fdot| */
static inline void spa_R(double fdot, double fddot, double s, double u, double *re, double *im) {
    const double w = u * u;
    if (u <= 0.0009765625) {
        *re = fma(w, fma(w, k13_asym_re[2], k13_asym_re[1]), 1.0) * s;
        *im = u * fma(w, fma(w, k13_asym_im[2], k13_asym_im[1]), k13_asym_im[0]) * s;
    } else if (u <= 0.03125) {
        double pr = k13_asym_re[6], pi = k13_asym_im[6];
        for (int k = 5; k >= 0; k--) { pr = fma(pr, w, k13_asym_re[k]); pi = fma(pi, w, k13_asym_im[k]); }
        *re = pr * s; *im = u * pi * s;
    } else if (u <= 1.0) {
        int ex;
        const double X = 1.0 / u, mant = frexp(X, &ex);
        int oct = ex - 1;
        oct = oct < 0 ? 0 : (oct > K13_NOCT - 1 ? K13_NOCT - 1 : oct);
        const double sv = 2.0 / mant - 3.0;
        double pr = k13_poly_re[oct][K13_DEG], pi = k13_poly_im[oct][K13_DEG];
        for (int k = K13_DEG - 1; k >= 0; k--) { pr = fma(pr, sv, k13_poly_re[oct][k]); pi = fma(pi, sv, k13_poly_im[oct][k]); }
        *re = pr * s; *im = pi * s;
    } else {
        const double af = fabs(fdot), X = 2.0943951023931953 * af * af * af / (fddot * fddot), q = -0.25 * X * X;
        double A = k13_ser_a[K13_NSER - 1], B = k13_ser_b[K13_NSER - 1];
        for (int k = K13_NSER - 2; k >= 0; k--) { A = fma(A, q, k13_ser_a[k]); B = fma(B, q, k13_ser_b[k]); }
        const double c13 = 1.2599210498948732, x13 = cbrt(X), cb = c13 * B, ca = x13 * x13 * A / c13;
        const double ure = 0.8660254037844386 * (cb - ca), uim = 0.5 * (cb + ca);
        const double ang = -(X + 0.7853981633974483), sn = sin(ang), cs = cos(ang);
        const double sc = 1.4472025091165353 * cbrt(1.4472025091165353 / fabs(fddot));
        *re = sc * (ure * cs - uim * sn); *im = sc * (ure * sn + uim * cs);
    }
}

static double solve_cold(double c1, double c2, double c3, double delta, double xl, double xh, double sdir, double hj) {
    double gl = xl * fma(xl, fma(xl, c3, c2), c1) - delta, gh = xh * fma(xh, fma(xh, c3, c2), c1) - delta;
    double x = (gh == gl) ? 0.5 * (xl + xh) : xl - gl * (xh - xl) / (gh - gl);
    x = fmin(fmax(x, xl), xh);
    for (int it = 0; it < 80; it++) {
        const double gx = x * fma(x, fma(x, c3, c2), c1) - delta, dg = fma(x, fma(3.0 * c3, x, 2.0 * c2), c1);
        if (gx * sdir > 0.0) xh = x; else xl = x;
        double xn = x - gx / dg;
        if (!(xn >= xl && xn <= xh)) xn = 0.5 * (xl + xh);
        const double dx = fabs(xn - x);
        x = xn;
        if (dx <= 1e-10 * hj) break;
    }
    return x;
}

int cpuf_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void cpuf_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* One walker on the f >= 0 half of the implicit grid f_i = (i - zero) * val.  coeff [L][2K+4][4], branches [K][4] / nbr [K] from
 * the oracle's spline / segmentation.  hp, hc [npos] complex interleaved (may be NULL); dw [2][npos] complex whitened data and
 * wf [2][npos] noise factor (may be NULL); like3 = (ll, <d|h>, <h|h>) as Likelihood.get_ll defines them; n_eval = stationary
 * points solved. */
int cpuf_sum(const double *t, const double *coeff, int L, int K, const int32_t *m_arr, const int32_t *n_arr, const double *ylm,
             int64_t N, double val, const branch_t *branches, const int32_t *nbr, int include_minus_m, double scale,
             double cos2psi, double sin2psi, const double *dw, const double *wf, double *hp, double *hc, double *like3,
             int64_t *n_eval) {
    const int R = 2 * K + 4;
    const int64_t zero = (N - 1) / 2, npos = zero + 1;
    /* ---- (m, n) groups and their combined amplitude quads ---- */
    int *grp = (int *)malloc(sizeof(int) * K), *lead = (int *)malloc(sizeof(int) * K);
    int G = 0;
    for (int k = 0; k < K; k++) {
        int g = -1;
        for (int q = 0; q < G; q++) if (m_arr[lead[q]] == m_arr[k] && n_arr[lead[q]] == n_arr[k]) { g = q; break; }
        if (g < 0) { g = G; lead[G++] = k; }
        grp[k] = g;
    }
    double *gq = (double *)calloc((size_t)L * G * 16, sizeof(double));
    const double r2 = 0.7071067811865476;
    for (int j = 0; j < L; j++)
        for (int k = 0; k < K; k++) {
            const double *a = coeff + ((size_t)j * R + k) * 4, *b = coeff + ((size_t)j * R + K + k) * 4;
            const double ypx = ylm[2 * k], ypy = ylm[2 * k + 1], ymx = ylm[2 * (K + k)], ymy = ylm[2 * (K + k) + 1];
            const double ypr = -r2 * (ypx + ypy), ypi = r2 * (ypx - ypy), ymr = r2 * (ymy - ymx), ymi = -r2 * (ymx + ymy);
            double *o = gq + ((size_t)j * G + grp[k]) * 16;
            for (int c = 0; c < 4; c++) {
                o[c] += ypr * a[c] - ypi * b[c];
                o[4 + c] += ypr * b[c] + ypi * a[c];
                o[8 + c] += ymr * a[c] + ymi * b[c];
                o[12 + c] += ymi * a[c] - ymr * b[c];
            }
        }
    /* knot phases in cycles as double-doubles */
    double *U = (double *)malloc(sizeof(double) * 4 * L);
    for (int j = 0; j < L; j++)
        for (int q = 0; q < 2; q++) {
            const double ph = coeff[((size_t)j * R + 2 * K + 2 + q) * 4];
            double a = ph * EMRIFD_INV2PI_HI, e = fma(ph, EMRIFD_INV2PI_HI, -a);
            a -= rint(a);
            e = fma(ph, EMRIFD_INV2PI_LO, e);
            const double hi = a + e;
            U[4 * j + 2 * q] = hi; U[4 * j + 2 * q + 1] = e - (hi - a);
        }
    const int64_t ntiles = (npos + TILE - 1) / TILE;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    int64_t evals = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : s0, s1, s2, evals)
    for (int64_t tile = 0; tile < ntiles; tile++) {
        const int64_t jt0 = tile * TILE, jt1 = (jt0 + TILE < npos ? jt0 + TILE : npos) - 1;
        const int nt = (int)(jt1 - jt0 + 1);
        double wpr[TILE], wpi[TILE], wmr[TILE], wmi[TILE];
        int touched = 0;
        for (int g = 0; g < G; g++) {
            const int kl = lead[g];
            const double dm = (double)m_arr[kl], dn = (double)n_arr[kl];
            const int mirror = (m_arr[kl] > 0) && include_minus_m;
            for (int q = 0; q < nbr[kl]; q++) {
                const branch_t *b = &branches[(size_t)kl * MAXBR + q];
                if (b->end < b->start) continue;
                for (int side = 0; side < 2; side++) {
                    /* tile-local bin range [s, e] of this side */
                    int64_t lo, hi;
                    if (side == 0) { lo = b->start - (zero + jt0); hi = b->end - (zero + jt0); }
                    else { lo = (zero - jt0) - b->end; hi = (zero - jt0) - b->start; }
                    int s = (int)(lo < 0 ? 0 : (lo > nt ? nt : lo)), e = (int)(hi > nt - 1 ? nt - 1 : (hi < -1 ? -1 : hi));
                    if (side == 1 && jt0 == 0 && s == 0) s = 1;
                    if (s > e) continue;
                    if (!touched) { memset(wpr, 0, sizeof(wpr)); memset(wpi, 0, sizeof(wpi)); memset(wmr, 0, sizeof(wmr)); memset(wmi, 0, sizeof(wmi)); touched = 1; }
                    double *dr = side == 0 ? wpr : wmr, *di = side == 0 ? wpi : wmi;
                    double *mr = side == 0 ? wmr : wpr, *mi = side == 0 ? wmi : wpi;
                    const double sg = side == 0 ? 1.0 : -1.0, sdir = (double)b->dir;
                    int j = -1;
                    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, d2 = 0, d3 = 0, segA = 0, segB = 0, xlo = 0, xhi = 0, tol = 0, tj = 0, hj = 0;
                    double mu_hi = 0, mu_lo = 0, p1 = 0, p2 = 0, p3 = 0, xprev = 0;
                    const double *amp = NULL;
                    for (int lb = s; lb <= e; lb++) {
                        const double f = sg * ((double)(jt0 + lb) * val);
                        const int inside = (j >= 0) && (b->dir > 0 ? (f >= segA && f < segB) : (f <= segA && f > segB));
                        double x;
                        if (!inside) {
                            int l2 = b->ja, h2 = b->jb;
                            if (j >= 0) { /* walk from the previous segment */
                                l2 = j;
                                if (sdir * sg > 0) { while (l2 < b->jb) { const double *cq = coeff + ((size_t)(l2 + 1) * R + 2 * K) * 4; const double Fk = dm * cq[0] + dn * cq[4]; if (b->dir > 0 ? (Fk <= f) : (Fk >= f)) l2++; else break; } }
                                else { while (l2 > b->ja) { const double *cq = coeff + ((size_t)l2 * R + 2 * K) * 4; const double Fk = dm * cq[0] + dn * cq[4]; if (b->dir > 0 ? (Fk > f) : (Fk < f)) l2--; else break; } }
                            } else {
                                while (l2 < h2) {
                                    const int mid = (l2 + h2 + 1) >> 1;
                                    const double *cq = coeff + ((size_t)mid * R + 2 * K) * 4;
                                    const double Fk = dm * cq[0] + dn * cq[4];
                                    if (b->dir > 0 ? (Fk <= f) : (Fk >= f)) l2 = mid; else h2 = mid - 1;
                                }
                            }
                            j = l2;
                            const double *cq = coeff + ((size_t)j * R + 2 * K) * 4;
                            tj = t[j]; hj = t[j + 1] - tj;
                            c0 = dm * cq[0] + dn * cq[4]; c1 = fma(dm, cq[1], dn * cq[5]); c2 = fma(dm, cq[2], dn * cq[6]); c3 = fma(dm, cq[3], dn * cq[7]);
                            d2 = 2.0 * c2; d3 = 3.0 * c3;
                            const double xl0 = (j == b->ja) ? b->xa : 0.0, xh0 = (j == b->jb) ? b->xb : hj;
                            segA = (j == b->ja) ? fma(xl0, fma(xl0, fma(xl0, c3, c2), c1), c0) : c0;
                            const double *cqn = coeff + ((size_t)(j + 1) * R + 2 * K) * 4;
                            segB = (j == b->jb) ? fma(xh0, fma(xh0, fma(xh0, c3, c2), c1), c0) : dm * cqn[0] + dn * cqn[4];
                            tol = 1e-10 * hj; xlo = xl0 - 1e-5 * hj; xhi = xh0 + 1e-5 * hj;
                            mu_hi = fma(dm, U[4 * j], dn * U[4 * j + 2]) - (b->dir < 0 ? 0.25 : 0.0);
                            mu_lo = fma(dm, U[4 * j + 1], dn * U[4 * j + 3]);
                            p1 = -EMRIFD_INV2PI_HI * fma(dm, cq[9], dn * cq[13]);
                            p2 = -EMRIFD_INV2PI_HI * fma(dm, cq[10], dn * cq[14]);
                            p3 = -EMRIFD_INV2PI_HI * fma(dm, cq[11], dn * cq[15]);
                            amp = gq + ((size_t)j * G + g) * 16;
                            x = solve_cold(c1, c2, c3, f - c0, xlo, xhi, sdir, hj);
                        } else { /* warm start: Newton from the previous bin's root */
                            x = xprev;
                            const double delta = f - c0;
                            int ok = 0;
                            for (int it = 0; it < 4; it++) {
                                const double gx = x * fma(x, fma(x, c3, c2), c1) - delta;
                                const double dx = gx / fma(x, fma(d3, x, d2), c1);
                                x -= dx;
                                if (fabs(dx) <= tol) { ok = 1; break; }
                            }
                            if (!ok || !(x >= xlo && x <= xhi)) x = solve_cold(c1, c2, c3, delta, xlo, xhi, sdir, hj);
                        }
                        xprev = x;
                        evals++;
                        const double fd = fma(x, fma(d3, x, d2), c1), fdd = fma(2.0 * d3, x, d2);
                        const double sv = 1.0 / sqrt(fabs(fd)), sv2 = sv * sv;
                        const double u = 0.477464829275686 * (fdd * fdd) * (sv2 * sv2 * sv2);
                        double re, im;
                        spa_R(fd, fdd, sv, u, &re, &im);
                        if (b->dir < 0) im = -im;
                        double p0 = f * tj;
                        const double e0 = fma(f, tj, -p0);
                        p0 -= rint(p0);
                        const double poly = fma(f, x, x * fma(x, fma(x, p3, p2), p1));
                        double sn, cs;
                        sincos_cycles(((p0 - mu_hi) + (e0 - mu_lo)) + poly, &sn, &cs);
                        const double er = re * cs - im * sn, ei = re * sn + im * cs;
                        const double cr = fma(x, fma(x, fma(x, amp[3], amp[2]), amp[1]), amp[0]);
                        const double ci = fma(x, fma(x, fma(x, amp[7], amp[6]), amp[5]), amp[4]);
                        dr[lb] += cr * er - ci * ei; di[lb] += cr * ei + ci * er;
                        if (mirror) {
                            const double qr = fma(x, fma(x, fma(x, amp[11], amp[10]), amp[9]), amp[8]);
                            const double qi = fma(x, fma(x, fma(x, amp[15], amp[14]), amp[13]), amp[12]);
                            mr[lb] += qr * er + qi * ei; mi[lb] += qi * er - qr * ei;
                        }
                    }
                }
            }
        }
        /* ---- read-out: S = -flip(W), Hermitian split, scale, rotation; likelihood sums ---- */
        for (int lb = 0; lb < nt; lb++) {
            const int64_t jj = jt0 + lb;
            double hpr = 0, hpi = 0, hxr = 0, hxi = 0;
            if (touched) {
                double a = wpr[lb], bq = wpi[lb], c = wmr[lb], d = wmi[lb];
                if (jj == 0) { a += c; bq += d; c = a; d = bq; }
                const double pr_ = 0.5 * (-c - a), pi_ = 0.5 * (-d + bq), xr_ = 0.5 * (d + bq), xi_ = 0.5 * (-c + a);
                hpr = scale * (cos2psi * pr_ - sin2psi * xr_); hpi = scale * (cos2psi * pi_ - sin2psi * xi_);
                hxr = scale * (sin2psi * pr_ + cos2psi * xr_); hxi = scale * (sin2psi * pi_ + cos2psi * xi_);
            }
            if (hp) { hp[2 * jj] = hpr; hp[2 * jj + 1] = hpi; hc[2 * jj] = hxr; hc[2 * jj + 1] = hxi; }
            if (dw) {
                const double w0 = wf[jj], w1 = wf[npos + jj];
                const double h0r = hpr * w0, h0i = hpi * w0, h1r = hxr * w1, h1i = hxi * w1;
                const double d0r = dw[2 * jj], d0i = dw[2 * jj + 1], d1r = dw[2 * (npos + jj)], d1i = dw[2 * (npos + jj) + 1];
                const double a0 = d0r - h0r, a1 = d0i - h0i, a2 = d1r - h1r, a3 = d1i - h1i;
                s0 += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
                s1 += d0r * h0r + d0i * h0i + d1r * h1r + d1i * h1i;
                s2 += h0r * h0r + h0i * h0i + h1r * h1r + h1i * h1i;
            }
        }
    }
    if (like3) { like3[0] = -0.5 * 4.0 * s0; like3[1] = 4.0 * s1; like3[2] = 4.0 * s2; }
    if (n_eval) *n_eval = evals;
    free(grp); free(lead); free(gq); free(U);
    return 0;
}

/*
 * emrihost.c -- native HOST-side producers for the FD path (gcc, OpenMP over walkers).
 *
 * Per BASELINE.json's north_star the trajectory ODE "stays on the host as the reference's sequential
 * integrator" (FastEMRIWaveforms integrates in C++: upstream src/Inspiral.cc, src/Utility.cc; SURVEY.md
 * section 2.2, section 8f rank 3).  This file is the native stand-in for that producer:
 *   - Schwarzschild fundamental frequencies Omega_phi, Omega_r from complete elliptic integrals
 *     (K, E by the arithmetic-geometric mean, Pi by Carlson's R_J duplication algorithm) -- row A2 of the scope table, called inside
 *     FDInterpolatedModeSum.sum for the L sparse points;
 *   - an adaptive Dormand-Prince 5(4) integrator of (p, e, Phi_phi, Phi_r) with the step control and dense
 *     output of SciPy's RK45, whose accepted steps ARE the sparse trajectory, stopping 0.1 outside the
 *     separatrix or at T (same ODE as trajectory/inspiral.py, which remains the pure-Python twin);
 *   - a batched entry point that integrates many walkers in parallel (one OpenMP thread per walker).
 * Nothing here runs on the GPU and nothing on the GPU path depends on it; it only feeds inputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DIST_TO_SEP 0.1

/* ---- Carlson symmetric forms R_C, R_J (duplication; Carlson 1995) ----------------------------- */
static double carlson_rc(double x, double y) { /* y > 0 */
    if (x == y) return 1.0 / sqrt(x);
    if (x < y) { double d = sqrt((y - x) / x); return atan(d) / sqrt(y - x); }
    double d = sqrt((x - y) / x);
    return atanh(d) / sqrt(x - y);
}

static double carlson_rj(double x, double y, double z, double p) { /* p > 0 */
    double A0 = (x + y + z + 2.0 * p) / 5.0, A = A0;
    double delta = (p - x) * (p - y) * (p - z);
    double Q = fmax(fmax(fmax(fabs(A0 - x), fabs(A0 - y)), fabs(A0 - z)), fabs(A0 - p)) / pow(0.25 * 1e-17, 1.0 / 6.0);
    double p4 = 1.0, sum = 0.0;
    double x0 = x, y0 = y, z0 = z, pp0 = p;
    for (int n = 0; n < 80 && p4 * Q >= fabs(A); n++) {
        double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z), sp = sqrt(p);
        double lam = sx * sy + sy * sz + sz * sx;
        double d = (sp + sx) * (sp + sy) * (sp + sz);
        double e = p4 * p4 * p4 * delta / (d * d);
        /* RC(1, 1+e) = atan(sqrt e)/sqrt e = sum_k (-e)^k/(2k+1): after the first iteration or two |e| (it shrinks 64x per
           iteration) is small enough for eleven terms to be exact to 1e-18; the closed form (atan / atanh, sqrt, divisions) was
           half of the time of the trajectory's right-hand side */
        double rc;
        if (fabs(e) < 0.02) {
            rc = 1.0 / 23.0;
            rc = 1.0 / 21.0 - e * rc; rc = 1.0 / 19.0 - e * rc; rc = 1.0 / 17.0 - e * rc; rc = 1.0 / 15.0 - e * rc;
            rc = 1.0 / 13.0 - e * rc; rc = 1.0 / 11.0 - e * rc; rc = 1.0 / 9.0 - e * rc; rc = 1.0 / 7.0 - e * rc;
            rc = 1.0 / 5.0 - e * rc; rc = 1.0 / 3.0 - e * rc; rc = 1.0 - e * rc;
        } else {
            rc = carlson_rc(1.0, 1.0 + e);
        }
        sum += p4 / d * rc;
        A = (A + lam) * 0.25; x = (x + lam) * 0.25; y = (y + lam) * 0.25; z = (z + lam) * 0.25; p = (p + lam) * 0.25;
        p4 *= 0.25;
    }
    double X = (A0 - x0) * p4 / A, Y = (A0 - y0) * p4 / A, Z = (A0 - z0) * p4 / A;
    double P = (-X - Y - Z) / 2.0;
    (void)pp0;
    double E2 = X * Y + X * Z + Y * Z - 3.0 * P * P, E3 = X * Y * Z + 2.0 * E2 * P + 4.0 * P * P * P;
    double E4 = (2.0 * X * Y * Z + E2 * P + 3.0 * P * P * P) * P, E5 = X * Y * Z * P * P;
    double ser = 1.0 - 3.0 * E2 / 14.0 + E3 / 6.0 + 9.0 * E2 * E2 / 88.0 - 3.0 * E4 / 22.0 - 9.0 * E2 * E3 / 52.0 + 3.0 * E5 / 26.0;
    return p4 * ser / (A * sqrt(A)) + 6.0 * sum;
}


/* Pi(n, m) = K(m) + n/3 R_J(0, 1 - m, 1, 1 - n); K and E come from the AGM in schw_freqs */

/* ---- A2: Schwarzschild Omega_phi, Omega_r (dimensionless) ----------------------------------- */
static void schw_freqs(double p, double e, double *om_phi, double *om_r) {
    double m = 4.0 * e / (p - 6.0 + 2.0 * e);
    /* complete K(m), E(m) by the arithmetic-geometric mean (quadratic convergence: five or six square roots in all, against
       ~35 for R_F + R_D): K = pi / (2 agm(1, sqrt(1 - m))), E = K (1 - sum_n 2^(n-1) c_n^2), c_0^2 = m */
    double K, E;
    {
        double a = 1.0, b = sqrt(1.0 - m), c2sum = 0.5 * m, pw = 0.5;
        for (int it = 0; it < 12; it++) {
            const double c = 0.5 * (a - b);
            if (fabs(c) <= 4e-16 * a) break; /* (c enters the next a only at second order) */
            const double an = 0.5 * (a + b);
            b = sqrt(a * b); a = an;
            pw *= 2.0; c2sum += pw * c * c;
        }
        K = M_PI / (a + b);           /* a = b to rounding: 2 agm = a + b */
        E = K * (1.0 - c2sum);
    }
    const double n1 = 16.0 * e / (12.0 + 8.0 * e - 4.0 * e * e - 8.0 * p + p * p);
    const double n2 = 2.0 * e * (p - 4.0) / ((1.0 + e) * (p - 6.0 + 2.0 * e));
    const double P1 = K + n1 / 3.0 * carlson_rj(0.0, 1.0 - m, 1.0, 1.0 - n1);
    const double P2 = K + n2 / 3.0 * carlson_rj(0.0, 1.0 - m, 1.0, 1.0 - n2);
    double p2 = p * p;
    double B = (-2.0 * P2 * (6.0 + 2.0 * e - p) * (3.0 + e * e - p) * p2) / ((-1.0 + e) * (1.0 + e) * (1.0 + e))
             - (E * (-4.0 + p) * p2 * (-6.0 + 2.0 * e + p)) / (-1.0 + e * e)
             + (K * p2 * (28.0 + 4.0 * e * e - 12.0 * p + p2)) / (-1.0 + e * e)
             + (4.0 * (-4.0 + p) * p * (2.0 * (1.0 + e) * K + P2 * (-6.0 - 2.0 * e + p))) / (1.0 + e)
             + 2.0 * (-4.0 + p) * (-4.0 + p) * (K * (-4.0 + p) + (P1 * p * (-6.0 - 2.0 * e + p)) / (2.0 + 2.0 * e - p));
    double D = (p - 2.0) * (p - 2.0) - 4.0 * e * e;
    *om_phi = 2.0 * p * sqrt(p) / (sqrt(D) * (8.0 + B / (K * (p - 4.0) * (p - 4.0))));
    *om_r = M_PI * p * sqrt((p - 6.0 + 2.0 * e) / D) / (8.0 * K + B / ((p - 4.0) * (p - 4.0)));
}

void emrihost_schwarzschild_frequencies(const double *p, const double *e, int64_t n, double *om_phi, double *om_r) {
    for (int64_t i = 0; i < n; i++) schw_freqs(p[i], e[i], &om_phi[i], &om_r[i]);
}

/* ---- trajectory: y = (p, e, Phi_phi, Phi_r), t in units of M ---------------------------------- */
static void rhs(const double *y, double q, double *f) {
    double p = y[0], e = y[1] > 0.0 ? y[1] : 0.0;
    double ome2 = 1.0 - e * e, s = ome2 * sqrt(ome2);
    f[0] = -(64.0 / 5.0) * q * s * (1.0 + 7.0 / 8.0 * e * e) / (p * p * p);
    f[1] = -(304.0 / 15.0) * q * e * s * (1.0 + 121.0 / 304.0 * e * e) / (p * p * p * p);
    schw_freqs(p, e, &f[2], &f[3]);
}

static const double A_[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5.0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0}};
static const double B_[6] = {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0};
static const double E_[7] = {-71.0 / 57600.0, 0.0, 71.0 / 16695.0, -71.0 / 1920.0, 17253.0 / 339200.0, -22.0 / 525.0, 1.0 / 40.0};
static const double P_[7][4] = {
    {1.0, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
    {0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};

#define NV 4
static double rms_norm(const double *v) {
    double s = 0.0;
    for (int i = 0; i < NV; i++) s += v[i] * v[i];
    return sqrt(s / NV);
}
static double gsep(const double *y) { return y[0] - (6.0 + 2.0 * y[1] + DIST_TO_SEP); }

/* dense output of the last step: y(t_old + theta*h) = y_old + h * theta * sum_s K[s] * (P[s] . theta^k) */
static void dense(const double *yold, double h, double K[7][NV], double theta, double *y) {
    double pw[4] = {theta, theta * theta, theta * theta * theta, theta * theta * theta * theta};
    for (int i = 0; i < NV; i++) {
        double acc = 0.0;
        for (int s = 0; s < 7; s++) {
            double b = P_[s][0] * pw[0] + P_[s][1] * pw[1] + P_[s][2] * pw[2] + P_[s][3] * pw[3];
            acc += K[s][i] * b;
        }
        y[i] = yold[i] + h * acc;
    }
}

/* returns number of points (>=2) or a negative error; outputs in seconds (t) and dimensionless p, e, phases */
int emrihost_trajectory(double M, double mu, double p0, double e0, double Phi_phi0, double Phi_r0, double T_years,
                        double rtol, double atol, int max_len, double *t_out, double *p_out, double *e_out,
                        double *Pp_out, double *Pr_out, double *fphi_out, double *fr_out) {
    const double MTSUN = 4.925491025873693e-06, YR = 31558149.763545603;
    if (!(M > 0 && mu > 0) || e0 < 0.0 || e0 >= 1.0 || p0 < 6.0 + 2.0 * e0 + DIST_TO_SEP) return -1;
    const double q = mu / M, Msec = M * MTSUN, t_end = T_years * YR / Msec;
    double y[NV] = {p0, e0, Phi_phi0, Phi_r0}, f0[NV], t = 0.0;
    double K[7][NV];
    rhs(y, q, f0);
    /* SciPy select_initial_step (order 4 error estimator) */
    double h;
    {
        double sc[NV], d0v[NV], d1v[NV], y1[NV], f1[NV], d2v[NV];
        for (int i = 0; i < NV; i++) { sc[i] = atol + fabs(y[i]) * rtol; d0v[i] = y[i] / sc[i]; d1v[i] = f0[i] / sc[i]; }
        double d0 = rms_norm(d0v), d1 = rms_norm(d1v);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        if (h0 > t_end) h0 = t_end;
        for (int i = 0; i < NV; i++) y1[i] = y[i] + h0 * f0[i];
        rhs(y1, q, f1);
        for (int i = 0; i < NV; i++) d2v[i] = (f1[i] - f0[i]) / sc[i];
        double d2 = rms_norm(d2v) / h0;
        double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h = fmin(fmin(100.0 * h0, h1), t_end);
    }
    int n = 0;
    double om_phi, om_r;
#define PUSH(tt, yy)                                                                  \
    do {                                                                              \
        if (n >= max_len) return -2;                                                  \
        t_out[n] = (tt) * Msec; p_out[n] = (yy)[0]; e_out[n] = (yy)[1] > 0 ? (yy)[1] : 0.0; \
        Pp_out[n] = (yy)[2]; Pr_out[n] = (yy)[3];                                     \
        schw_freqs(p_out[n], e_out[n], &om_phi, &om_r);                               \
        fphi_out[n] = om_phi / (2.0 * M_PI * Msec); fr_out[n] = om_r / (2.0 * M_PI * Msec); \
        n++;                                                                          \
    } while (0)
    PUSH(t, y);
    double fcur[NV];
    memcpy(fcur, f0, sizeof(fcur));
    for (int guard = 0; guard < 100000 && t < t_end; guard++) {
        double min_step = 10.0 * fabs(nextafter(t, INFINITY) - t);
        if (h < min_step) h = min_step;
        int accepted = 0, rejected = 0;
        double ynew[NV], fnew[NV], tnew = t, hstep = h;
        while (!accepted) {
            if (hstep < min_step) return -3;
            tnew = t + hstep;
            if (tnew > t_end) tnew = t_end;
            hstep = tnew - t;
            memcpy(K[0], fcur, sizeof(fcur));
            for (int s = 1; s < 6; s++) {
                double ys[NV];
                for (int i = 0; i < NV; i++) {
                    double acc = 0.0;
                    for (int j = 0; j < s; j++) acc += A_[s][j] * K[j][i];
                    ys[i] = y[i] + hstep * acc;
                }
                rhs(ys, q, K[s]);
            }
            for (int i = 0; i < NV; i++) {
                double acc = 0.0;
                for (int s = 0; s < 6; s++) acc += B_[s] * K[s][i];
                ynew[i] = y[i] + hstep * acc;
            }
            rhs(ynew, q, fnew);
            memcpy(K[6], fnew, sizeof(fnew));
            double errv[NV];
            for (int i = 0; i < NV; i++) {
                double acc = 0.0;
                for (int s = 0; s < 7; s++) acc += E_[s] * K[s][i];
                double sc = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol;
                errv[i] = hstep * acc / sc;
            }
            double en = rms_norm(errv);
            if (en < 1.0) {
                double fac = (en == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(en, -0.2));
                if (rejected) fac = fmin(1.0, fac);
                h = hstep * fac;
                accepted = 1;
            } else {
                hstep *= fmax(0.2, 0.9 * pow(en, -0.2));
                rejected = 1;
            }
        }
        /* terminal event: crossing of p = 6 + 2e + 0.1 from above */
        if (gsep(y) > 0.0 && gsep(ynew) <= 0.0) {
            double lo = 0.0, hi = 1.0, yv[NV];
            for (int it = 0; it < 200; it++) { /* bisection on the dense output (SciPy: brentq with xtol 4 eps) */
                double mid = 0.5 * (lo + hi);
                dense(y, tnew - t, K, mid, yv);
                if (gsep(yv) > 0.0) lo = mid; else hi = mid;
                if (hi - lo < 4.0 * 2.220446049250313e-16) break;
            }
            dense(y, tnew - t, K, hi, yv);
            PUSH(t + hi * (tnew - t), yv);
            return n;
        }
        t = tnew;
        memcpy(y, ynew, sizeof(y));
        memcpy(fcur, fnew, sizeof(fcur));
        PUSH(t, y);
    }
    return n;
}

/* batched: walker i writes to rows i of [nb][max_len] arrays; lens[i] = points or negative error code */
void emrihost_trajectory_batch(int64_t nb, const double *M, const double *mu, const double *p0, const double *e0,
                               const double *Phi_phi0, const double *Phi_r0, double T_years, double rtol, double atol,
                               int max_len, double *t, double *p, double *e, double *Pp, double *Pr, double *fphi,
                               double *fr, int32_t *lens) {
    /* serial on purpose: 0.3-1 ms per walker, and an OpenMP team spinning next to torch's own thread pools cost far more
       than it saved; callers that want parallelism run this from several host threads (ctypes releases the GIL) */
    for (int64_t i = 0; i < nb; i++) {
        size_t o = (size_t)i * max_len;
        lens[i] = emrihost_trajectory(M[i], mu[i], p0[i], e0[i], Phi_phi0[i], Phi_r0[i], T_years, rtol, atol, max_len,
                                      t + o, p + o, e + o, Pp + o, Pr + o, fphi + o, fr + o);
    }
}

/* multi-threaded batch: plain pthreads created per call (no resident spinning team), walkers handed out through an atomic
   counter because their cost varies (0.3-1 ms each) */
#include <pthread.h>
typedef struct {
    int64_t nb;
    const double *M, *mu, *p0, *e0, *Pp0, *Pr0;
    double T, rtol, atol;
    int max_len;
    double *t, *p, *e, *Pp, *Pr, *fphi, *fr;
    int32_t *lens;
    int64_t next;
} traj_job_t;

static void *traj_worker(void *arg) {
    traj_job_t *j = (traj_job_t *)arg;
    for (;;) {
        int64_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->nb) break;
        size_t o = (size_t)i * j->max_len;
        j->lens[i] = emrihost_trajectory(j->M[i], j->mu[i], j->p0[i], j->e0[i], j->Pp0[i], j->Pr0[i], j->T, j->rtol, j->atol,
                                         j->max_len, j->t + o, j->p + o, j->e + o, j->Pp + o, j->Pr + o, j->fphi + o, j->fr + o);
    }
    return NULL;
}

void emrihost_trajectory_batch_mt(int64_t nb, const double *M, const double *mu, const double *p0, const double *e0,
                                  const double *Phi_phi0, const double *Phi_r0, double T_years, double rtol, double atol,
                                  int max_len, double *t, double *p, double *e, double *Pp, double *Pr, double *fphi,
                                  double *fr, int32_t *lens, int nthreads) {
    traj_job_t job = {nb, M, mu, p0, e0, Phi_phi0, Phi_r0, T_years, rtol, atol, max_len, t, p, e, Pp, Pr, fphi, fr, lens, 0};
    if (nthreads > nb) nthreads = (int)nb;
    if (nthreads > 64) nthreads = 64;
    if (nthreads <= 1) { traj_worker(&job); return; }
    pthread_t th[64];
    int started = 0;
    for (int k = 0; k < nthreads - 1; k++) {
        if (pthread_create(&th[started], NULL, traj_worker, &job) == 0) started++;
    }
    traj_worker(&job);
    for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
}

int emrihost_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

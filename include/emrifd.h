/*
 * emrifd.h -- C-ABI of libemrifd.so: the B200-native FD EMRI mode-sum + likelihood hot path.
 *
 * The reference (lorenzsp/EMRI_FrequencyDomainWaveforms) has no native code of its own: its
 * scripts reach this arithmetic through FastEMRIWaveforms' Cython modules (pyinterp /
 * pyinterp_cpu, upstream src/interpolate.cu).  Each entry point below
 * names the reference-side interface it replaces.  All entry points:
 *   - take plain pointers and sizes (no torch / numpy types),
 *   - return 0 on success or a negative EMRIFD_ERR_* code; they never throw or abort,
 *   - launch on the handle's stream and do NOT synchronise unless stated ("sync"),
 *   - treat pointers as DEVICE pointers unless the name ends in _host,
 *   - use FP64 / interleaved complex128, C-contiguous.
 *
 * Frequency grid convention (FDInterpolatedModeSum.sum: fftshift(fftfreq(N, dt)) or a user
 * f_arr; emri_pe.py:237-239,339-349): N odd, symmetric, exactly one zero at index (N-1)/2.
 *   implicit grid:  f_i = (i - (N-1)/2) * val,   val = 1/(N*dt)      (fpos == NULL)
 *   explicit grid:  f_i = +-fpos[|i - (N-1)/2|], fpos[0..(N-1)/2]     (fpos != NULL)
 *
 * Spline coefficient layout ("knot-major quads"): coeff[L][R][4] = (y, c1, c2, c3) per
 * (knot, row); rows ordered [Re A_0..Re A_{K-1} | Im A_0..Im A_{K-1} | f_phi, f_r, Phi_phi, Phi_r]
 * (row order of few's FDInterpolatedModeSum y_all; R = 2K + 4).
 */
#ifndef EMRIFD_H
#define EMRIFD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMRIFD_VERSION 200
#define EMRIFD_MAX_BRANCHES 4 /* monotone branches per mode (turnovers + 1) */
#define EMRIFD_MAX_KNOTS 1024 /* sparse trajectory length limit (few: max_init_len = 1000) */

#define EMRIFD_OK 0
#define EMRIFD_ERR_INVALID (-1)       /* bad argument (NULL, size <= 0, even N, ...) */
#define EMRIFD_ERR_TOO_FEW_KNOTS (-2) /* not-a-knot needs L >= 4 */
#define EMRIFD_ERR_KNOT_ORDER (-3)    /* knots not strictly increasing */
#define EMRIFD_ERR_BRANCHES (-4)      /* a mode has more than EMRIFD_MAX_BRANCHES monotone branches */
#define EMRIFD_ERR_NOMEM (-5)
#define EMRIFD_ERR_CUDA (-6)          /* CUDA runtime error; see emrifd_last_error */
#define EMRIFD_ERR_TOO_MANY_KNOTS (-7)
#define EMRIFD_ERR_NO_DATA (-8)       /* likelihood requested before emrifd_set_data */

/* evaluation of the SPA factor's K_{1/3} (Tutorial_FD_construction_single_mode.ipynb:599-609 uses scipy.special.kv) */
#define EMRIFD_K13_EXACT 0 /* <= 3e-15 everywhere (default) */
#define EMRIFD_K13_FEW 1   /* FastEMRIWaveforms' SPAFunc: 14-term ascending series for |X| <= 7, 9-term asymptotic above
                              (2.5e-7 off at the seam) -- for bit-level comparisons against FEW's CPU backend */

/* flags for the mode-sum entry points */
#define EMRIFD_INCLUDE_MINUS_M 1 /* add the mirrored -m term (include_minus_m=True) */
#define EMRIFD_MASK_POSITIVE 2   /* output only f >= 0 bins (mask_positive=True) */

typedef struct emrifd_handle emrifd_handle_t;

/* One monotone branch of one mode's f_mn(t) (the work-list entry of the segmentation step).
 * Bins start..end (full-grid indices, inclusive; empty if end < start) have exactly one
 * stationary point on this branch.  (ja, xa) -> (jb, xb): time-ordered end points as
 * (segment, offset inside segment); Fa, Fb the mode frequency there; dir = sign(Fb - Fa). */
typedef struct {
    int32_t mode, dir, ja, jb;
    int32_t closed_end, pad;
    int64_t start, end;
    double xa, xb, Fa, Fb;
} emrifd_branch_t;

/* Per-waveform descriptor of a ragged batch ("walker").  Offsets index the packed arrays. */
typedef struct {
    int32_t L, K;
    int64_t knot_off;  /* into t / f_phi / f_r / Phi_phi / Phi_r           (doubles)            */
    int64_t teuk_off;  /* into teuk, complex elements; walker block is [L][K]                    */
    int64_t mode_off;  /* into m_arr / n_arr; x2 for ylm (complex: +m block then -m block)      */
    int64_t coeff_off; /* into coeff, in doubles; walker block is [L][2K+4][4]                   */
    int64_t out_off;   /* into hp / hc, complex elements                                          */
    double scale;      /* mu*MRSUN_SI/(dist*Gpc)  (Tutorial_FD_construction_single_mode.ipynb:623) */
    double cos2psi, sin2psi; /* SSB polarisation rotation (GenerateEMRIWaveform)                  */
} emrifd_walker_t;

/* ---- lifetime -------------------------------------------------------------------------- */
int emrifd_version(void);
int emrifd_sizeof_branch(void);
int emrifd_sizeof_walker(void);
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for the default */
int emrifd_create(int device, void *stream, emrifd_handle_t **out);
int emrifd_destroy(emrifd_handle_t *h);
int emrifd_set_stream(emrifd_handle_t *h, void *stream);
int emrifd_synchronize(emrifd_handle_t *h); /* sync */
const char *emrifd_last_error(emrifd_handle_t *h);

/* ---- A3: not-a-knot cubic spline ---------------------------------------------------------
 * Replaces few.summation.interpolatedmodesum.CubicSplineInterpolant(t, y_all)
 * (Tutorial_FD_construction_single_mode.ipynb:176,380; upstream pyinterp interpolate_arrays_wrap).
 * y(r, j) = y[r*row_stride + j*knot_stride]; coeff out [L][R][4]. */
int emrifd_spline_build(emrifd_handle_t *h, const double *t, const double *y, int64_t L, int64_t R,
                        int64_t row_stride, int64_t knot_stride, double *coeff);
/* CubicSplineInterpolant.__call__(t_new): out[R][n], end-segment extrapolation */
int emrifd_spline_eval(emrifd_handle_t *h, const double *t, const double *coeff, int64_t L, int64_t R,
                       const double *tnew, int64_t n, double *out);

/* ---- A3+A4+A5+A6+A7: FDInterpolatedModeSum.sum for a ragged batch --------------------------
 * Replaces few.summation.fdinterp.FDInterpolatedModeSum.sum (upstream pyinterp get_waveform_fd)
 * as called from GenerateEMRIWaveform(..., sum_kwargs={'output_type':'fd'}) (emri_pe.py:86-105,212).
 * Packed device inputs: t, f_phi, f_r, Phi_phi, Phi_r [sum L]; teuk [sum L*K] complex;
 * m_arr, n_arr [sum K] int32; ylm [sum 2K] complex; walkers [B] (device copy made internally
 * from the HOST array `walkers`).  Work buffers coeff [sum L*(2K+4)*4] and branches
 * [sum K * EMRIFD_MAX_BRANCHES] are caller-owned device memory (results of A3 / A4 stay
 * inspectable).  Output hp/hc complex: N per walker, or (N+1)/2 with EMRIFD_MASK_POSITIVE.
 * Steps can be run separately (spline -> segment -> sum) or through emrifd_fd_waveform_batch. */
int emrifd_batch_spline(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                        const double *t, const double *teuk, const double *f_phi, const double *f_r,
                        const double *Phi_phi, const double *Phi_r, double *coeff);
int emrifd_batch_segment(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                         const double *t, const double *coeff, const int32_t *m_arr, const int32_t *n_arr,
                         int64_t N, double val, const double *fpos, emrifd_branch_t *branches,
                         int64_t *n_eval /* [B][2] device: #stationary points, #MBE; may be NULL */);
int emrifd_batch_sum(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                     const double *t, const double *coeff, const int32_t *m_arr, const int32_t *n_arr,
                     const double *ylm, const emrifd_branch_t *branches,
                     int64_t N, double val, const double *fpos, int flags,
                     int64_t j_lo, int64_t j_cnt, /* positive-bin slice [j_lo, j_lo+j_cnt); 0,(N+1)/2 = all */
                     double *hp, double *hc,      /* may both be NULL (likelihood only) */
                     double *like_out /* [B][3] device: ll, <d|h>, <h|h>; NULL = no likelihood */);
/* Cyclic tile sharding of the mode sum (multi-GPU frequency-bin sharding of ONE long, high-mode-count waveform, SURVEY.md
 * section 8e(2)): the f >= 0 bins are cut into tiles of emrifd_tile_bins() bins and this call evaluates tiles tile_first,
 * tile_first + tile_stride, ... (rank r of W passes r, W).  Neighbouring tiles carry similar work, so the interleave balances
 * the ranks without a work histogram, and every rank keeps several waves of tiles.  like_out [B][3] holds the partial sums
 * over the owned tiles (all_reduce(SUM) them); hp/hc, if given, need EMRIFD_MASK_POSITIVE and receive the owned tiles only. */
int emrifd_tile_bins(void);
int emrifd_batch_sum_cyclic(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                            const double *t, const double *coeff, const int32_t *m_arr, const int32_t *n_arr,
                            const double *ylm, const emrifd_branch_t *branches,
                            int64_t N, double val, const double *fpos, int flags, int64_t tile_first, int64_t tile_stride,
                            double *hp, double *hc, double *like_out);
int emrifd_fd_waveform_batch(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                             const double *t, const double *teuk, const double *f_phi, const double *f_r,
                             const double *Phi_phi, const double *Phi_r,
                             const int32_t *m_arr, const int32_t *n_arr, const double *ylm,
                             int64_t N, double val, const double *fpos, int flags,
                             double *coeff, emrifd_branch_t *branches,
                             double *hp, double *hc, double *like_out);
/* status of the last batch on this handle: reads the device error word (first failure of ANY walker). sync. */
int emrifd_batch_status(emrifd_handle_t *h);
/* Per-walker status of the last batch that went through emrifd_batch_segment / emrifd_fd_waveform_batch /
 * emrifd_loglike_batch_host: status_host[B] = 0 or EMRIFD_ERR_KNOT_ORDER / EMRIFD_ERR_BRANCHES.  A failing walker gets
 * h = 0 and ll = <d|h> = <h|h> = NaN; every other walker of the batch is bit-identical to a clean run.  This is the
 * reference's per-walker contract: Eryn maps a NaN likelihood to -1e300 (Eryn/eryn/moves/red_blue.py:282-284) and the sweep
 * skips a failing point (check_mode_by_mode.py:328-330).  sync. */
int emrifd_walker_status(emrifd_handle_t *h, int64_t B, int32_t *status_host);
/* same words copied to a DEVICE buffer on the handle's stream, no sync (pipelined callers read them back with the results) */
int emrifd_walker_status_dev(emrifd_handle_t *h, int64_t B, int32_t *status_dev);
/* 1 (default): the zero-fill of the empty tiles runs on an internal stream underneath mode_sum_kernel (joined before the call's
 * results are used); 0: after it on the handle's stream (profiling: the sum's duration without the store stream beside it).
 * Results are identical either way. */
int emrifd_set_overlap(emrifd_handle_t *h, int enable);
/* EMRIFD_K13_EXACT (default) or EMRIFD_K13_FEW for every later mode sum on this handle. */
int emrifd_set_k13_mode(emrifd_handle_t *h, int mode);

/* ---- A10/A11: PSD-weighted inner product and likelihood ------------------------------------
 * emrifd_set_data replaces lisatools Likelihood.inject_signal's stored state
 * (LISAanalysistools/lisatools/sampling/likelihood.py:213-220): whitened data
 * d*sqrt(df/S) [nch=2][n] complex and noise_factor sqrt(df/S) [2][n], n = (N+1)/2 positive bins.
 * Device pointers; the handle keeps the pointers (caller keeps the memory alive). */
int emrifd_set_data(emrifd_handle_t *h, const double *d_whitened, const double *noise_factor, int64_t n);
/* lisatools.diagnostic.inner_product(sig1, sig2, f_arr=, PSD=array|None) (diagnostic.py:95-110):
 * out[0] = 4 * sum_ch sum_k dx_k Re(conj(a) b)/S_k, out[1] = same with Im (complex=True).
 * a, b: [nch][n] complex; freqs [n]; psd [n] or NULL.  out: device double[2]. */
int emrifd_inner_product(emrifd_handle_t *h, const double *a, const double *b, int64_t nch, int64_t n,
                         const double *freqs, const double *psd, double *out);
/* Likelihood.get_ll on materialised templates (likelihood.py:257-274): templates [B][2][n]
 * complex -> out [B][3] = (ll, 4 sum Re(d~* h w), 4 sum |h w|^2) against emrifd_set_data. */
int emrifd_loglike(emrifd_handle_t *h, const double *templates, int64_t B, double *out);

/* ---- reference-facing call with HOST buffers (the e2e path) --------------------------------
 * One call = what Likelihood.get_ll does per batch of walkers on the reference
 * (likelihood.py:245-274): FD template for every walker + ll against the injected data.
 * All array arguments are HOST pointers (pinned or pageable); H2D of the packed sparse inputs,
 * all kernels and the D2H of ll[B][3] happen inside; returns after the results are on the host.
 * If hp_dev/hc_dev are given the f >= 0 waveforms are also materialised on the device (like the
 * reference's GPU path, where h stays a device array and only the likelihood crosses to the host). */
int emrifd_loglike_batch_host(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B,
                              const double *t, const double *teuk, const double *f_phi, const double *f_r,
                              const double *Phi_phi, const double *Phi_r,
                              const int32_t *m_arr, const int32_t *n_arr, const double *ylm,
                              int64_t N, double val, const double *fpos_dev, int flags,
                              double *hp_dev, double *hc_dev, /* optional DEVICE outputs [B][(N+1)/2] complex (walker out_off), or NULL */
                              double *like_out_host /* [B][3] */);

/* ---- SURVEY section 8f rank 1 ("next"): Ylm, mode selection by power and mode compaction on the device -------
 * The producers that sit immediately before the path in FastSchwarzschildEccentricFlux.__call__ (upstream few/waveform.py;
 * in-repo call sites emri_pe.py:86-105,212): ylm_gen -> mode_selector -> teuk_modes[:, keep].  With them on the device the
 * full-basis amplitudes never cross PCIe and a walker batch needs one small D2H (the kept-mode counts).
 *
 * emrifd_ylm_batch replaces few.utils.ylm.GetYlms(assume_positive_m=True)(l, m, theta, phi) plus the per-mode expansion
 * (Tutorial_FD_construction_single_mode.ipynb:87,597-611): l_arr, m_arr [M] mode basis (m >= 0), neg_src [Mneg] = basis
 * index of each m > 0 mode (the -m copies), theta, phi [B]; ylm_out [B][M + Mneg] complex = Y_{l,m} block then Y_{l,-m}. */
int emrifd_ylm_batch(emrifd_handle_t *h, const int32_t *l_arr, const int32_t *m_arr, int64_t M, const int32_t *neg_src,
                     int64_t Mneg, const double *theta, const double *phi, int64_t B, int lmax, double *ylm_out);
/* emrifd_mode_select replaces few.utils.modeselector.ModeSelector.__call__ (semantics SURVEY.md A.4).  teuk [nsamp][M]
 * complex (all walkers' time samples back to back), samp_walker [nsamp] walker of each sample, ylm [B][M + Mneg] as
 * above, flags [B][M] bytes out (1 = keep; the union over the walker's samples, -m picks folded onto +m). */
int emrifd_mode_select(emrifd_handle_t *h, const double *teuk, int64_t nsamp, int64_t M, const int32_t *samp_walker,
                       const double *ylm, const int32_t *neg_src, int64_t Mneg, int64_t B, double eps, uint8_t *flags);
/* Compaction (the `teuk_modes[:, keep]`, `ylms[keep ++ keep_neg]`, `ls/ms/ns[keep]` indexing of ModeSelector.__call__):
 * step 1 writes each walker's ascending kept indices keep_idx [B][M] (first K valid) and K_out [B]; the caller reads K_out,
 * fills the walker descriptors (K and offsets), then step 2 gathers into the packed layout of emrifd_fd_waveform_batch.
 * teuk_full [sum L][M] complex (walker rows at knot_off), neg_pos [M] = position of a mode in the -m block or -1 (m = 0). */
int emrifd_mode_compact_count(emrifd_handle_t *h, const uint8_t *flags, int64_t B, int64_t M, int32_t *keep_idx, int32_t *K_out);
int emrifd_mode_compact_gather(emrifd_handle_t *h, const emrifd_walker_t *walkers, int64_t B, const double *teuk_full, int64_t M,
                               int64_t Mneg, const int32_t *keep_idx, const int32_t *m_basis, const int32_t *n_basis,
                               const int32_t *neg_pos, const double *ylm_full, double *teuk_out, int32_t *m_out, int32_t *n_out,
                               double *ylm_out);

/* Stand-in amplitude producer on the device (offline substitute for few.amplitude.romannet.RomanAmplitude, whose weights are
 * a Zenodo download; formula in amplitude/synthetic.py): p, e [nsamp] trajectory points, mode basis l/m/n [M], cmode [M]
 * complex per-mode constants; teuk_out [nsamp][M] complex -- the layout emrifd_mode_select / emrifd_mode_compact_gather read. */
int emrifd_synth_amplitude(emrifd_handle_t *h, const double *p, const double *e, int64_t nsamp, const int32_t *l_arr,
                           const int32_t *m_arr, const int32_t *n_arr, const double *cmode, int64_t M, int lmax, int nmax,
                           double *teuk_out);

/* ---- measurement helpers (bench.py roofline denominators) ----------------------------------- */
/* FP64 FMA throughput micro-benchmark: returns achieved GFLOP/s in *gflops. sync. */
int emrifd_bench_fp64_fma(emrifd_handle_t *h, int iters, double *gflops);
/* ---- section 8f rank 2: FD window convolution (FDutils.py:35-47 get_convolution, :66-101 get_fd_windowed) --------------
 * The reference evaluates convolve(hstack((a[1:], a)), b, 'valid') / len(b) with a = conj(fft(window)), i.e. the circular
 * convolution out[k] = (1/N) sum_i a[i] b[(k - i) mod N].  The DFT of the windows the scripts use (scipy.signal.windows hann,
 * blackman, hamming, nuttall, blackmanharris: check_mode_by_mode.py:43,269; emri_pe.py:261) is concentrated in a few taps around
 * i = 0 (mod N), so the product path applies a banded stencil of 2 H + 1 taps and reports the truncation bound
 * (energy of the dropped taps, by Parseval); lengths N need not be powers of two (the 1-yr grid is 3 155 815).
 *
 * emrifd_window_taps: taps[2 (H + 2)] doubles (device) <- W_j = sum_n window[n] e^{-2 pi i j n / N} for j = 0..H as (re, im)
 *   pairs (a REAL time-domain window: W_{-j} = conj W_j), then (sum_n window[n]^2, 0) in slot H + 1.  By Parseval the energy of
 *   the taps outside the band is N sum w^2 - (|W_0|^2 + 2 sum_{j=1..H} |W_j|^2).
 * emrifd_band_energy: for a window given in the frequency domain (window_in_fd=True): energy[H + 2] doubles (device) <-
 *   sum of |a_i|^2 over circular distance min(i, N - i) == h for h = 0..H, and everything farther out in slot H + 1.
 * emrifd_band_convolve: out[nch][out_n] complex <- (1/N) sum_{i = -H..H} taps[i + H] signal[c][(k - i) mod N] for
 *   k = out_lo .. out_lo + out_n - 1; taps [2 H + 1] complex in the order i = -H..H (device).  Signal and output must not alias. */
#define EMRIFD_WINDOW_MAX_TAPS 256
int emrifd_window_taps(emrifd_handle_t *h, const double *window, int64_t N, int H, double *taps);
int emrifd_band_energy(emrifd_handle_t *h, const double *a, int64_t N, int H, double *energy);
int emrifd_band_convolve(emrifd_handle_t *h, const double *taps, int H, const double *signal, int64_t nch, int64_t N,
                         int64_t out_lo, int64_t out_n, double *out);

/* number of kernels this handle has launched since creation (bench.py "gpu_launches") */
int64_t emrifd_launch_count(emrifd_handle_t *h);
/* CUDA-event time (ms) accumulated by the dominant kernel (mode-sum) since the last reset, and launches */
int emrifd_sum_kernel_time(emrifd_handle_t *h, int enable, double *ms, int64_t *launches);
/* the same with the split: ms_pair = tile classification + mode_sum_kernel and the concurrent empty_tile_kernel, up to their
 * join (one bracket), ms_mode_sum = mode_sum_kernel alone (events recorded on the handle's stream while timing is enabled) */
int emrifd_sum_kernel_times(emrifd_handle_t *h, int enable, double *ms_pair, double *ms_mode_sum, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* EMRIFD_H */

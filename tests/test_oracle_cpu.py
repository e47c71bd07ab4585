"""CPU tests: pin the oracle (oracle/emrifd_oracle.c) against the golden vectors.

The goldens come from the reference's own lisatools code, SciPy and mpmath
(tests/golden/make_golden.py); the FD-vs-FFT(TD) test is the first-principles check of the
sign/flip/split conventions since FastEMRIWaveforms cannot run here.
"""
import os

import numpy as np
import pytest

from helpers import CASES, make_item, oracle_waveform, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_spline_matches_scipy_golden(oracle_quad, oracle_f64):
    g = np.load(os.path.join(GOLD, "spline_golden.npz"))
    c = oracle_quad.spline_build(g["t"], g["y"])
    ref = g["coeff"]
    # compare each coefficient by its contribution to the interpolant over its own interval
    # (c3 = tau/h suffers cancellation on short intervals in ANY double implementation, SciPy's included)
    h = np.diff(g["t"])[:, None, None] ** np.arange(4)[None, None, :]
    yscale = np.max(np.abs(g["y"]), axis=1)[None, :, None]
    assert np.max(np.abs(c[:-1] - ref) * h / yscale) < 1e-13
    assert np.array_equal(c[:, :, 0], g["y"].T)
    assert np.array_equal(c, oracle_f64.spline_build(g["t"], g["y"]))      # spline build is always plain double
    ev = oracle_quad.spline_eval(g["t"], c, g["tq"])
    assert np.max(np.abs(ev - g["yq"]) / np.max(np.abs(g["yq"]), axis=1, keepdims=True)) < 1e-12


def test_psd_spline_with_extrapolation(oracle_quad):
    g = np.load(os.path.join(GOLD, "spline_golden.npz"))
    S = np.load(os.path.join(os.path.dirname(GOLD), "..", "emri_frequencydomainwaveforms_b200", "data", "lisa_alloc_sh.npy"))
    c = oracle_quad.spline_build(S[:, 0], S[:, 1][None, :])
    ev = oracle_quad.spline_eval(S[:, 0], c, g["psd_f"])[0]
    assert np.all(ev[:3] > 0)                                  # S(0) > 0: bin 0 stays in the likelihood
    assert np.max(np.abs(ev - g["psd_val"]) / np.abs(g["psd_val"])) < 1e-9


def test_k13_factor_against_mpmath_and_scipy(oracle_quad, oracle_f64):
    g = np.load(os.path.join(GOLD, "k13_golden.npz"))
    X = g["X"]
    conv = np.sqrt(2 * X / np.pi) * np.exp(-1j * np.pi / 4)
    for orc, tol in ((oracle_quad, 4e-16), (oracle_f64, 2e-15)):
        R = np.array([orc.spa_R(x) for x in X])
        assert np.max(np.abs(R - g["Q_mpmath"] * conv) / np.abs(g["Q_mpmath"] * conv)) < tol
    # the notebook's own evaluation (scipy.special.kv) agrees with the same numbers
    assert np.max(np.abs(g["Q_scipy"] - g["Q_mpmath"]) / np.abs(g["Q_mpmath"])) < 1e-14
    # small-X form S = R / X^(1/6) is finite as X -> 0 (turnover regime)
    S0 = oracle_quad.spa_S(1e-12)
    assert np.isfinite(S0.real) and abs(S0) > 0.1


def test_inner_product_and_likelihood_against_lisatools(oracle_quad):
    g = np.load(os.path.join(GOLD, "lisatools_golden.npz"))
    a, b, f, psd = g["a"], g["b"], g["f"], g["psd"]
    assert np.isclose(oracle_quad.inner_product(a, b, f, psd), g["ip_ab"], rtol=1e-13)
    assert np.isclose(oracle_quad.inner_product(a[:1], b[:1], f, psd), g["ip_a0b0"], rtol=1e-13)
    naa, nbb = oracle_quad.inner_product(a, a, f, psd), oracle_quad.inner_product(b, b, f, psd)
    assert np.isclose(oracle_quad.inner_product(a, b, f, psd) / np.sqrt(naa * nbb), g["ip_ab_norm"], rtol=1e-13)
    assert np.isclose(np.sqrt(naa), g["snr_a"], rtol=1e-13)
    # likelihood: whitened data / noise factor exactly as Likelihood.inject_signal stores them
    for k, (idx, amp) in enumerate(g["like_params"]):
        ll = oracle_quad.loglike(g["like_injection"], amp * g["templates"][int(idx)], g["like_noise_factor"])[0]
        assert np.isclose(ll, g["like_ll"][k], rtol=1e-11, atol=1e-14)


def test_waveform_regression_goldens(oracle_quad):
    for name in ("plunge", "ecc_many"):
        g = np.load(os.path.join(GOLD, f"waveform_{name}.npz"))
        N, dt = int(g["N"]), float(g["dt"])
        hp, hc, coeff, br, nbr = oracle_quad.fd_sum(g["t"], g["teuk_modes"], g["ylms"], g["Phi_phi"], g["Phi_r"],
                                                    g["m_arr"], g["n_arr"], g["f_phi"], g["f_r"], N, 1.0 / (N * dt),
                                                    scale=float(g["scale"]))
        idx = g["nz_index"]
        assert np.max(np.abs(hp[idx] - g["hp_nz"])) <= 1e-13 * np.max(np.abs(g["hp_nz"]))
        assert np.max(np.abs(hc[idx] - g["hc_nz"])) <= 1e-13 * np.max(np.abs(g["hc_nz"]))
        for key in ("start", "end", "ja", "jb", "dir"):
            assert np.array_equal(br[key], g["branches"][key])
        nz = np.nonzero((hp != 0) | (hc != 0))[0]
        assert len(nz) == int(g["nnz_total"]) and nz.min() == int(g["support_lo"]) and nz.max() == int(g["support_hi"])
        assert oracle_quad.last_n_eval == int(g["n_eval"])


def test_double_oracle_tracks_quad_oracle(generator, oracle_quad, oracle_f64):
    it = make_item(generator, "plunge", dt=40.0)
    hq = oracle_waveform(oracle_quad, it)
    hd = oracle_waveform(oracle_f64, it)
    assert rel_err(hd[0], hq[0]) < 5e-10 and rel_err(hd[1], hq[1]) < 5e-10     # plain-double phase floor
    assert np.array_equal(hd[3]["start"], hq[3]["start"]) and np.array_equal(hd[3]["end"], hq[3]["end"])


def test_fd_matches_fft_of_time_domain(generator, oracle_f64):
    """First-principles pin of the conventions: S = h+ - i hx must be the DFT (numpy sign) of the
    time-domain mode sum built from the same splines (cf. notebook cell 26)."""
    from scipy.interpolate import CubicSpline
    it = make_item(generator, "plunge", dt=40.0)
    # Harmonics whose frequency turns over (here the m = 0 ones, f = n f_r) are reproduced to ~1 %
    # only: the SPA/K_{1/3} model is zero beyond the turnover frequency where the true spectrum has an
    # Airy tail.  Monotone harmonics are the sharp test of the conventions.
    keep = np.nonzero(it["m_arr"] > 0)[0]
    K0 = len(it["m_arr"])
    it = dict(it, teuk_modes=np.ascontiguousarray(it["teuk_modes"][:, keep]), m_arr=it["m_arr"][keep],
              n_arr=it["n_arr"][keep], ylms=np.concatenate([it["ylms"][keep], it["ylms"][K0 + keep]]))
    N, dt, t = it["N"], it["dt"], it["t"]
    hp, hc, *_ = oracle_waveform(oracle_f64, it)
    num_pts = int(min(t[-1], int(it["T"] * 31558149.763545603 / dt) * dt) / dt) + 1
    tt = np.arange(num_pts) * dt
    A = CubicSpline(t, it["teuk_modes"], axis=0)(tt)
    Pp, Pr = CubicSpline(t, it["Phi_phi"])(tt), CubicSpline(t, it["Phi_r"])(tt)
    K = len(it["m_arr"])
    h = np.zeros(num_pts, dtype=complex)
    for k in range(K):
        ph = it["m_arr"][k] * Pp + it["n_arr"][k] * Pr
        h += it["ylms"][k] * A[:, k] * np.exp(-1j * ph)
        if it["m_arr"][k] > 0:
            h += it["ylms"][K + k] * np.conj(A[:, k]) * np.exp(1j * ph)
    h *= it["scale"] * np.hanning(num_pts)
    hpad = np.zeros(N, dtype=complex)
    hpad[:num_pts] = h
    H = np.fft.fftshift(np.fft.fft(hpad)) * dt
    Hp = np.fft.fftshift(np.fft.fft(hpad.real)) * dt
    Hx = np.fft.fftshift(np.fft.fft(-hpad.imag)) * dt
    # window the FD model the way FDutils.get_fd_windowed does: convolve with the window's DFT
    win = np.zeros(N)
    win[:num_pts] = np.hanning(num_pts)
    conv = lambda x: np.fft.fftshift(np.fft.fft(np.fft.ifft(np.fft.ifftshift(x)) * win))
    ov = lambda x, y: np.real(np.vdot(x, y)) / np.sqrt(np.real(np.vdot(x, x)) * np.real(np.vdot(y, y)))
    assert ov(H, conv(hp - 1j * hc)) > 0.9999
    assert ov(Hp, conv(hp)) > 0.9999 and ov(Hx, conv(hc)) > 0.9999
    # Hermitian symmetry of the two polarisations (real time series)
    assert np.allclose(hp, np.conj(hp[::-1]), rtol=0, atol=1e-30) and np.allclose(hc, np.conj(hc[::-1]), rtol=0, atol=1e-30)


def test_segmentation_properties(generator, oracle_quad):
    it = make_item(generator, "plunge", dt=40.0)
    hp, hc, coeff, br, nbr = oracle_waveform(oracle_quad, it)
    N = it["N"]
    f = np.fft.fftshift(np.fft.fftfreq(N, it["dt"]))
    assert nbr.max() == 2                                      # plunging orbit: some harmonics turn over
    assert np.array_equal(nbr == 0, (it["m_arr"] == 0) & (it["n_arr"] == 0))   # (l,0,0): f = 0 identically, no branch
    for k in range(len(nbr)):
        for q in range(nbr[k]):
            b = br[k, q]
            lo, hi = min(b["Fa"], b["Fb"]), max(b["Fa"], b["Fb"])
            s, e = int(b["start"]), int(b["end"])
            if e >= s:
                assert lo <= f[max(s, 0)] and f[min(e, N - 1)] <= hi
                if s > 0:
                    assert f[s - 1] <= lo
                if e < N - 1:
                    assert f[e + 1] >= hi
            if q > 0:
                assert br[k, q - 1]["dir"] == -b["dir"]        # consecutive branches alternate direction
                assert br[k, q - 1]["Fb"] == b["Fa"]           # and share the turnover point
    # a truncated explicit f_arr reproduces the implicit grid at coincident frequencies (SURVEY section 4 property 4)
    zero = (N - 1) // 2
    sub = f[zero: zero + 2001]
    hp2, hc2, *_ = oracle_waveform(oracle_quad, it, N=4001, fpos=sub)
    assert np.array_equal(hp2[2000:], hp[zero: zero + 2001])


def test_mode_select_restatement_matches_host_selector(generator):
    """oracle.mode_select_ref (frozen summation orders, checker of the device kernel) picks the same modes as the
    product's host ModeSelector (few semantics) and obeys its defining properties."""
    from oracle.oracle import mode_select_ref, ylm_ref
    t, p, e, *_ = generator.inspiral_generator(1e6, 10.0, 0.0, 12.0, 0.35, 1.0, T=0.2, dt=10.0)
    teuk = generator.amplitude_generator(p, e)
    nl = len(generator.unique_l)
    y = generator.ylm_gen(generator.unique_l, generator.unique_m, 1.0, -np.pi / 2)
    ylms = np.concatenate([y[:nl][generator.inverse_lm], y[nl:][generator.inverse_lm][generator.m0mask]])
    prev = None
    for eps in (0.3, 1e-2, 1e-5):
        keep = mode_select_ref(teuk, ylms, generator.m0mask, eps)
        _, _, ls, ms, ns = generator.mode_selector(teuk, ylms, [generator.l_arr, generator.m_arr, generator.n_arr], eps=eps)
        assert np.array_equal(generator.l_arr[keep], ls) and np.array_equal(generator.m_arr[keep], ms) and np.array_equal(generator.n_arr[keep], ns)
        # kept power fraction >= 1 - eps at every time sample; smaller eps keeps a superset
        full = np.concatenate([teuk, np.conj(teuk[:, generator.m0mask])], axis=1) * ylms[None, :]
        pw = np.abs(full) ** 2
        pos = np.cumsum(generator.m0mask) - 1
        sel = np.concatenate([keep, generator.num_teuk_modes + pos[keep][generator.m0mask[keep]]])
        assert np.all(pw[:, sel].sum(axis=1) >= (1 - eps) * pw.sum(axis=1) * (1 - 1e-12))
        if prev is not None:
            assert set(prev) <= set(keep)
        prev = keep
    # Ylm restatement against the product's host GetYlms
    for (l, m) in [(2, 2), (2, -2), (5, 3), (10, -7), (10, 0)]:
        from emri_frequencydomainwaveforms_b200.utils.ylm import spin_weighted_ylm
        assert abs(ylm_ref(l, m, 0.9, 2.2) - spin_weighted_ylm(-2, l, m, 0.9, 2.2)) < 1e-15


# ---- the oracle against the reference's own statement of the per-harmonic construction ------------------------------
def test_cell26_golden_pins_oracle_conventions(oracle_quad):
    """tests/golden/cell26_golden.npz holds W(f) computed by the body of the reference notebook's FD_waveform
    (Tutorial_FD_construction_single_mode.ipynb:548-623, cell 26: SciPy CubicSpline + scipy.special.kv) on three monotone single
    harmonics (f_mn > 0 rising twice, f_mn < 0 falling), generated by tests/golden/make_cell26_golden.py.  The oracle must
    reproduce it after undoing its final flip / split, W(f) = -[h+ - i hx](-f): identical support (bin index sets), the sign,
    flip, conjugation and +-m conventions exactly (overlap real and positive), values to the accuracy of the cell's own
    approximations (t(f) from a spline of the inverse function, fddot from a spline through the fdot knots): measured mismatch
    1.5e-13 / 8.2e-10 / 1.7e-14, asserted <= 1e-8."""
    g = np.load(os.path.join(GOLD, "cell26_golden.npz"))
    assert len(g["names"]) == 3
    for name in g["names"]:
        M, mu, T, dt, N, scale = g[f"{name}.params"]
        N = int(N)
        l, m, n = (int(x) for x in g[f"{name}.lmn"])
        get = lambda k: g[f"{name}.{k}"]
        hp, hc, *_ = oracle_quad.fd_sum(get("t"), get("teuk_modes"), get("ylms"), get("Phi_phi"), get("Phi_r"), np.array([m], dtype=np.int32),
                                        np.array([n], dtype=np.int32), get("f_phi"), get("f_r"), N, 1.0 / (N * dt), scale=scale)
        W_or = -np.flip(hp - 1j * hc)                     # S = h+ - i hx = -flip(W)
        W_ref = np.fft.fftshift(g[f"{name}.W"])           # the cell works on an fftfreq-ordered grid
        assert np.array_equal(W_or != 0, W_ref != 0), name                       # identical bin index sets, both signs of f
        sup = W_ref != 0
        zero = (N - 1) // 2
        if "negative" in name:
            assert np.all(np.where(sup)[0][np.abs(W_ref[sup]) > 0.5 * np.abs(W_ref).max()] < zero)   # the direct term lives at f < 0
        ip = np.vdot(W_ref[sup], W_or[sup])
        nrm = np.sqrt(np.vdot(W_ref[sup], W_ref[sup]).real * np.vdot(W_or[sup], W_or[sup]).real)
        assert abs(ip.imag) / nrm <= 1e-5 and 1.0 - ip.real / nrm <= 1e-8, (name, ip / nrm)
        rel = np.abs(W_or[sup] - W_ref[sup]) / np.abs(W_ref).max()
        assert np.median(rel) <= 1e-7 and rel.max() <= 5e-3, (name, np.median(rel), rel.max())
        # a wrong convention is far outside these bounds: conjugating, flipping or negating W gives overlap <= 0 or ~ 0
        for wrong in (np.conj(W_or), np.flip(W_or), -W_or):
            assert np.vdot(W_ref[sup], wrong[sup]).real / nrm < 0.9


def test_k13_few_mode_differs_only_near_the_seam(oracle_f64):
    """EMRIFD_K13_FEW / Oracle.set_k13_mode("few"): FastEMRIWaveforms' SPAFunc truncations (14-term ascending series for
    |X| <= 7, 9-term asymptotic series above).  Against the exact evaluation the two differ by <= 1e-6 just
    below the seam (9.8e-7 at X = 6.96 from the truncated series, 2.4e-7 just above from the asymptotic one) and agree to
    rounding away from it."""
    Xs = np.concatenate([np.logspace(-3, 0, 20), np.linspace(1, 40, 300), np.logspace(np.log10(40), 6, 30)])
    try:
        oracle_f64.set_k13_mode("exact")
        exact = np.array([oracle_f64.spa_R(x) for x in Xs])
        oracle_f64.set_k13_mode("few")
        few = np.array([oracle_f64.spa_R(x) for x in Xs])
    finally:
        oracle_f64.set_k13_mode("exact")
    d = np.abs(few - exact) / np.abs(exact)
    assert d.max() <= 2e-6 and 4.0 < Xs[d.argmax()] < 9.0
    assert d[(Xs < 3.0) | (Xs > 30.0)].max() <= 1e-10
    assert d[(Xs > 6.0) & (Xs < 8.0)].max() >= 1e-8          # the switch really changes the evaluation at the seam


@pytest.mark.parametrize("name", list(CASES))
def test_fast_cpu_baseline_matches_oracle(name, generator, oracle_quad):
    """oracle/emrifd_cpu_fast.c (the optimised double CPU implementation that bench.py times as `cpu_baseline` and as the
    `--impl reference` arm) against the binary128 oracle: waveform <= 1e-9 of max|h|, identical support, likelihood sums."""
    from oracle.oracle import FastCPU
    fast = FastCPU()
    it = make_item(generator, name)
    N = it["N"]
    n, val = (N + 1) // 2, 1.0 / (N * it["dt"])
    hp_o, hc_o, *_ = oracle_waveform(oracle_quad, it)
    hp_o, hc_o = hp_o[n - 1:], hc_o[n - 1:]
    rng = np.random.default_rng(4)
    w = np.full((2, n), 2.0e19)
    d = np.stack([hp_o, hc_o]) * w + 1e-3 * np.abs(hp_o).max() * w * (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n)))
    hp, hc, like, nev = fast.sum(it, N, val, data_w=d, wfac=w)
    assert rel_err(hp, hp_o) <= 1e-9 and rel_err(hc, hc_o) <= 1e-9
    assert np.array_equal(hp != 0, hp_o != 0)
    ref = oracle_quad.loglike(d, np.stack([hp_o, hc_o]), w)
    assert np.allclose(like, ref, rtol=1e-9, atol=1e-9 * abs(ref[2]))
    _, _, like2, nev2 = fast.sum(it, N, val, data_w=d, wfac=w, want_h=False)            # likelihood only: same sums
    assert np.allclose(like2, like, rtol=1e-13) and nev2 == nev and 0 < nev <= oracle_quad.last_n_eval


def test_walker_cost_estimate_tracks_the_work_list(generator, oracle_f64):
    """engine.walker_cost_estimate (host-side, from the tracks alone: bins swept by every distinct (m, n) harmonic) against the
    exact number of stationary points the oracle's work-list implies (one per (m, n) group and covered bin): within 10 % on
    systems with and without turnovers -- good enough to balance walker shards across ranks (distributed.balanced_walker_assignment)."""
    from emri_frequencydomainwaveforms_b200 import engine
    gen, orc = generator, oracle_f64
    for name in ("cfg1_like", "plunge", "ecc_many"):
        it = make_item(gen, name, dt=10.0)
        N = it["N"]
        val = 1.0 / (N * it["dt"])
        R = np.concatenate([it["teuk_modes"].real.T, it["teuk_modes"].imag.T, [it["f_phi"], it["f_r"], it["Phi_phi"], it["Phi_r"]]])
        coeff = orc.spline_build(it["t"], R)
        br, nbr = orc.segment_build(it["t"], coeff, it["m_arr"], it["n_arr"], N, val)
        cnt = np.where(br["end"] >= br["start"], br["end"] - br["start"] + 1, 0).reshape(len(it["m_arr"]), -1).sum(axis=1)
        seen, exact = set(), 0
        for k, mn in enumerate(zip(it["m_arr"].tolist(), it["n_arr"].tolist())):
            if mn not in seen:
                seen.add(mn)
                exact += int(cnt[k])
        est = engine.walker_cost_estimate(it, val)
        assert abs(est - exact) <= 0.10 * exact, (name, est, exact)

"""GPU parity tests: the CUDA path (through the C-ABI) against the binary128 CPU oracle.

Tolerances (BASELINE.json north_star): per-bin error <= 1e-10 of max|h| in FP64, mismatch <= 1e-12,
bit-exact mode/bin index sets; spline coefficients are compared bit for bit as well.
"""
import numpy as np
import pytest

from helpers import CASES, make_item, oracle_waveform, rel_err

pytestmark = pytest.mark.gpu

TOL_BIN = 1e-10


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _gpu_sum(it, torch, **kw):
    from emri_frequencydomainwaveforms_b200.summation.fdinterp import FDInterpolatedModeSum
    s = FDInterpolatedModeSum(pad_output=True, output_type="fd", odd_len=True)
    out = s(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"],
            it["M"], it["p"], it["e"], T=it["T"], dt=it["dt"], scale=it["scale"], **kw)
    return s, out.cpu().numpy()


@pytest.mark.parametrize("name", list(CASES))
def test_spline_segment_sum_parity(name, generator, oracle_quad, torch_cuda):
    it = make_item(generator, name)
    hp_o, hc_o, coeff_o, br_o, nbr_o = oracle_waveform(oracle_quad, it)
    s, out = _gpu_sum(it, torch_cuda)
    assert out.shape == (2, it["N"])
    # A3: spline coefficients bit-exact
    coeff_g = s.last_batch.coeff_host(0)
    assert np.array_equal(coeff_g, coeff_o), f"max diff {np.max(np.abs(coeff_g - coeff_o))}"
    # A4: work-list bit-exact
    br_g = s.last_batch.branches_host()
    for key in ("mode", "dir", "ja", "jb", "closed_end", "start", "end", "xa", "xb", "Fa", "Fb"):
        assert np.array_equal(br_g[key], br_o[key]), key
    # A5-A7: waveform
    assert rel_err(out[0], hp_o) <= TOL_BIN
    assert rel_err(out[1], hc_o) <= TOL_BIN
    # identical support (bin index set)
    assert np.array_equal(out[0] != 0, hp_o != 0)
    # frequency attribute == fftshift(fftfreq(N, dt))
    assert np.array_equal(s.frequency.cpu().numpy(), np.fft.fftshift(np.fft.fftfreq(it["N"], it["dt"])))


def test_mask_positive_and_f_arr(generator, oracle_quad, torch_cuda):
    it = make_item(generator, "plunge")
    N = it["N"]
    _, full = _gpu_sum(it, torch_cuda)
    _, pos = _gpu_sum(it, torch_cuda, mask_positive=True)
    assert pos.shape == (2, (N + 1) // 2)
    assert np.array_equal(pos, full[:, (N - 1) // 2:])          # SURVEY section 4 property 2
    # down-sampled symmetric f_arr (emri_pe.py:333-349)
    f = np.fft.fftshift(np.fft.fftfreq(N, it["dt"]))
    nz = np.abs(full[0][(N - 1) // 2:]) > 0
    fp = f[(N - 1) // 2:]
    p_freq = np.linspace(0.0, fp[nz].max() * 1.01, num=int(nz.sum() / 20))
    newf = np.hstack((-p_freq[::-1][:-1], p_freq))
    s, ds = _gpu_sum(it, torch_cuda, f_arr=newf)
    hp_o, hc_o, *_ = oracle_waveform(oracle_quad, it, N=len(newf), fpos=p_freq)
    assert rel_err(ds[0], hp_o) <= TOL_BIN and rel_err(ds[1], hc_o) <= TOL_BIN
    assert np.array_equal(s.frequency.cpu().numpy(), newf)


def test_include_minus_m_and_rotation(generator, oracle_quad, torch_cuda):
    it = make_item(generator, "cfg1_like")
    c2, s2 = np.cos(0.7), np.sin(0.7)
    _, out = _gpu_sum(it, torch_cuda, include_minus_m=False, cos2psi=c2, sin2psi=s2)
    hp_o, hc_o, *_ = oracle_waveform(oracle_quad, it, include_minus_m=False, cos2psi=c2, sin2psi=s2)
    assert rel_err(out[0], hp_o) <= TOL_BIN and rel_err(out[1], hc_o) <= TOL_BIN


def test_cubic_spline_interpolant(oracle_quad, torch_cuda):
    from scipy.interpolate import CubicSpline
    from emri_frequencydomainwaveforms_b200.summation.interpolatedmodesum import CubicSplineInterpolant
    rng = np.random.default_rng(5)
    t = np.sort(rng.uniform(0, 3e7, 80)); t[0] = 0.0
    y = rng.normal(size=(9, 80)).cumsum(axis=1)
    sp = CubicSplineInterpolant(t, y)
    coeff = sp.coeff.cpu().numpy()
    assert np.array_equal(coeff, oracle_quad.spline_build(t, y))
    assert sp.interp_array.shape == (4, 80, 9)
    tn = np.linspace(-1e5, 3.01e7, 1001)               # includes extrapolation on both sides
    ev = sp(tn).cpu().numpy()
    ref = CubicSpline(t, y, axis=1)(tn)
    assert ev.shape == (9, 1001)
    assert np.max(np.abs(ev - ref)) <= 1e-11 * np.max(np.abs(ref))
    assert np.array_equal(ev, oracle_quad.spline_eval(t, coeff, tn))
    with pytest.raises(ValueError):
        CubicSplineInterpolant(t[:3], y[:, :3])         # not-a-knot needs >= 4 knots
    with pytest.raises(ValueError):
        CubicSplineInterpolant(t[::-1].copy(), y)       # knots must increase


# ------------------------------------------------------------------------------------------------
# A9-A11: PSD, inner product, likelihood
# ------------------------------------------------------------------------------------------------
import os
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_get_sensitivity_matches_scipy_golden(torch_cuda):
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    g = np.load(os.path.join(GOLD, "spline_golden.npz"))
    ev = get_sensitivity(g["psd_f"])
    assert isinstance(ev, np.ndarray) and np.all(ev[:3] > 0)            # extrapolated S(0) > 0
    assert np.max(np.abs(ev - g["psd_val"]) / np.abs(g["psd_val"])) < 1e-9


def test_inner_product_matches_lisatools_golden(torch_cuda):
    from emri_frequencydomainwaveforms_b200.lisatools.diagnostic import inner_product, snr
    g = np.load(os.path.join(GOLD, "lisatools_golden.npz"))
    a, b, f, psd = [g["a"][0], g["a"][1]], [g["b"][0], g["b"][1]], g["f"], g["psd"]
    assert np.isclose(inner_product(a, b, f_arr=f, PSD=psd), g["ip_ab"], rtol=1e-12)
    assert np.isclose(inner_product(a, b, f_arr=f, PSD=psd, normalize=True), g["ip_ab_norm"], rtol=1e-12)
    assert np.isclose(inner_product(a, b, f_arr=f, PSD=psd, complex=True), g["ip_ab_complex"], rtol=1e-12)
    assert np.isclose(inner_product(a, b, f_arr=f, PSD=psd, normalize="sig1"), g["ip_ab_sig1"], rtol=1e-12)
    assert np.isclose(inner_product(a[0], b[0], f_arr=f, PSD=psd), g["ip_a0b0"], rtol=1e-12)
    assert np.isclose(inner_product(a, b, df=f[1] - f[0], PSD=psd), g["ip_df"], rtol=1e-12)
    assert np.isclose(snr(a, f_arr=f, PSD=psd), g["snr_a"], rtol=1e-12)
    assert np.isclose(inner_product(a, a, f_arr=f, PSD=psd, normalize=True), 1.0, rtol=1e-14)
    ta = [torch_cuda.as_tensor(x).cuda() for x in a]                      # device tensors in -> same number
    assert np.isclose(inner_product(ta, ta, f_arr=torch_cuda.as_tensor(f).cuda(), PSD=torch_cuda.as_tensor(psd).cuda()),
                      inner_product(a, a, f_arr=f, PSD=psd), rtol=1e-15)
    with pytest.raises(ValueError):
        inner_product(a, b)
    with pytest.raises(ValueError):
        inner_product(a, b[:1], f_arr=f, PSD=psd)


def test_inner_product_time_domain_branch_matches_lisatools_golden(torch_cuda):
    """inner_product(..., dt=) (diagnostic.py:49-67: rfft * dt, DC dropped, zero padding) against the reference's own output."""
    import warnings
    from emri_frequencydomainwaveforms_b200.lisatools.diagnostic import inner_product, snr
    g = np.load(os.path.join(GOLD, "lisatools_td_golden.npz"))
    x, y, dt, psd = [g["x"][0], g["x"][1]], [g["y"][0], g["y"][1]], float(g["dt"]), g["psd"]
    assert np.isclose(inner_product(x, y, dt=dt, PSD=psd), g["ip_xy"], rtol=1e-11)
    assert np.isclose(inner_product(x, y, dt=dt, PSD=psd, normalize=True), g["ip_xy_norm"], rtol=1e-11)
    assert np.isclose(snr(x, dt=dt, PSD=psd), g["snr_x"], rtol=1e-11)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        v = inner_product(x, [c[:4000] for c in y], dt=dt, PSD=psd)
    assert np.isclose(v, g["ip_xy_short"], rtol=1e-11) and any("Zero padding" in str(w.message) for w in wlist)


def test_likelihood_mirror_matches_lisatools_golden(torch_cuda):
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import Likelihood
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    g = np.load(os.path.join(GOLD, "lisatools_golden.npz"))
    templates, f = g["templates"], g["f"]

    def model(idx, amp, **kw):
        return [amp * templates[int(idx)][0], amp * templates[int(idx)][1]]

    like = Likelihood(model, 2, f_arr=f, vectorized=False, transpose_params=False, subset=2)
    like.inject_signal(data_stream=[g["a"][0], g["a"][1]], noise_fn=[get_sensitivity, get_sensitivity], noise_kwargs=[{}, {}])
    assert np.allclose(like.noise_factor, g["like_noise_factor"], rtol=1e-9)
    ll = like(g["like_params"])
    assert ll.shape == (5,) and abs(ll[0]) < 1e-20                        # likelihood at the injection is 0
    assert np.allclose(ll[1:], g["like_ll"][1:], rtol=1e-8)
    with pytest.raises(ValueError):
        like(list(g["like_params"]))


def _emri_params(n, rng=None):
    base = np.array([1e6, 50.0, 0.0, 9.0, 0.3, 1.0, 1.0, 0.8, 0.4, 1.1, 2.0, 0.3, 0.0, 1.1])
    out = np.tile(base, (n, 1))
    if rng is not None and n > 1:
        out[1:, 0] *= 1 + 1e-5 * rng.normal(size=n - 1)
        out[1:, 3] += 1e-4 * rng.normal(size=n - 1)
        out[1:, 4] += 1e-4 * rng.normal(size=n - 1)
        out[1:, 11] += 1e-2 * rng.normal(size=n - 1)
    return out


def test_likelihood_skips_dc_bin_when_psd_is_nan_there(torch_cuda):
    """likelihood.py:268: start_ind = 1 when noise_factor[0, 0] is NaN (a PSD that is undefined at f = 0)."""
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import Likelihood
    rng = np.random.default_rng(4)
    n = 513
    f = np.linspace(0.0, 1e-2, n)
    data = [rng.normal(size=n) + 1j * rng.normal(size=n) for _ in range(2)]
    tmpl = [[c * (1.0 + 0.1 * k) for c in data] for k in range(3)]

    def psd_nan(freqs, **kw):
        out = 1e-2 + np.asarray(freqs) ** 2
        out = out.copy()
        out[0] = np.nan
        return out

    like = Likelihood(lambda k, **kw: tmpl[int(k)], 2, f_arr=f)
    like.inject_signal(data_stream=data, noise_fn=[psd_nan, psd_nan], noise_kwargs=[{}, {}])
    ll = like(np.arange(3.0)[:, None])
    nf = like.noise_factor
    ref = [-2.0 * np.sum(np.abs((np.asarray(data) - np.asarray(tmpl[k])) * nf)[:, 1:] ** 2) for k in range(3)]
    assert np.all(np.isfinite(ll)) and np.allclose(ll, ref, rtol=1e-12, atol=1e-18)
    with pytest.raises(NotImplementedError):
        Likelihood(lambda k, **kw: tmpl[int(k)], 2, f_arr=f, separate_d_h=True).get_ll(np.arange(3.0)[:, None])


def test_generate_emri_waveform_surface(oracle_quad, torch_cuda):
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    kw = dict(T=0.1, dt=20.0, eps=1e-2)
    sum_kwargs = dict(pad_output=True, output_type="fd", odd_len=True)
    gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=sum_kwargs, use_gpu=True, return_list=False)
    gen_list = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=sum_kwargs, use_gpu=True, return_list=True)
    p = _emri_params(1)[0]
    h = gen(*p, **kw)
    hl = gen_list(*p, **kw)
    freq = gen.waveform_generator.create_waveform.frequency                 # emri_pe.py:238
    N = h.shape[0]
    assert N % 2 == 1 and freq.shape[0] == N and int((freq == 0).sum()) == 1
    assert torch_cuda.equal(h, hl[0] - 1j * hl[1])                          # check_mode_by_mode.py:247 "check 1 =="
    pos = gen_list(*p, mask_positive=True, **kw)                            # emri_pe.py:241-243
    mask = (freq >= 0.0)
    assert torch_cuda.equal(pos[0], hl[0][mask]) and torch_cuda.equal(pos[1], hl[1][mask])
    # against the oracle, including distance scaling and the SSB polarisation rotation
    base = gen.waveform_generator
    theta, phi, c2, s2 = gen._transform(*p[7:11])
    it = base.prepare(p[0], p[1], p[3], p[4], theta, phi, dist=p[6], Phi_phi0=p[11], Phi_r0=p[13], **kw)
    hp_o, hc_o, *_ = oracle_quad.fd_sum(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"],
                                        it["n_arr"], it["f_phi"], it["f_r"], N, 1.0 / (N * kw["dt"]), scale=it["scale"],
                                        cos2psi=c2, sin2psi=s2)
    assert rel_err(hl[0].cpu().numpy(), hp_o) <= TOL_BIN and rel_err(hl[1].cpu().numpy(), hc_o) <= TOL_BIN
    with pytest.raises(ValueError):
        gen(*np.r_[p[:4], 0.9, p[5:]], **kw)                                # e0 out of range -> ValueError like FEW


def test_fused_batched_likelihood_against_oracle(oracle_quad, torch_cuda):
    """emri_pe.py-style: inject the FD signal, evaluate a ragged walker batch through the plugin
    (fill_data_noise=True path of likelihood.py:330-331) -- one fused launch sequence, no h(f) in HBM."""
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel, Likelihood
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    kw = dict(T=0.1, dt=20.0, eps=1e-2)
    gen_list = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True),
                                    use_gpu=True, return_list=True)
    rng = np.random.default_rng(3)
    params = _emri_params(5, rng)
    model = FDTemplateModel(gen_list)
    sig = model(*params[0], **kw)
    n = sig[0].shape[0]
    N = 2 * n - 1
    f_arr = np.arange(n) / (N * kw["dt"])
    like = Likelihood(model, 2, f_arr=f_arr, fill_data_noise=True, subset=3)
    like.inject_signal(data_stream=[sig[0], sig[1]], noise_fn=[get_sensitivity, get_sensitivity], noise_kwargs=[{}, {}])
    ll = like(params, **kw)
    assert ll.shape == (5,)
    # oracle: per-walker waveform on f >= 0 and the lisatools likelihood algebra
    items, ok = model.prepare_batch(params, **kw)
    assert ok.all() and len({it["teuk_modes"].shape for it in items}) >= 1
    dd = oracle_quad.loglike(like.injection_channels, np.zeros_like(like.injection_channels), like.noise_factor)[0]
    for i, it in enumerate(items):
        hp_o, hc_o, *_ = oracle_quad.fd_sum(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"],
                                            it["n_arr"], it["f_phi"], it["f_r"], N, 1.0 / (N * kw["dt"]), scale=it["scale"],
                                            cos2psi=it["cos2psi"], sin2psi=it["sin2psi"], out_lo=n - 1, out_n=n)
        ref = oracle_quad.loglike(like.injection_channels, np.stack([hp_o, hc_o]), like.noise_factor)[0]
        assert abs(ll[i] - ref) <= 1e-10 * abs(dd), (i, ll[i], ref)
    assert abs(ll[0]) <= 1e-10 * abs(dd) and np.all(ll[1:] < 0)             # emri_pe.py:451: ll(injection) = 0
    # materialised route (per-walker templates + emrifd_loglike) agrees with the fused route
    like2 = Likelihood(lambda *p, **k: model(*p, **k), 2, f_arr=f_arr)
    like2.inject_signal(data_stream=[sig[0], sig[1]], noise_fn=[get_sensitivity, get_sensitivity], noise_kwargs=[{}, {}])
    ll2 = like2(params[:3], **kw)
    assert np.allclose(ll2, ll[:3], rtol=1e-9, atol=1e-12 * abs(dd))
    # an out-of-domain walker yields NaN (Eryn maps it to -1e300), the others are unaffected
    bad = params.copy()
    bad[2, 4] = 0.95
    ll3 = like(bad, **kw)
    assert np.isnan(ll3[2]) and np.array_equal(ll3[[0, 1, 3, 4]], ll[[0, 1, 3, 4]])


def test_ragged_batch_equals_single_calls(generator, torch_cuda):
    from emri_frequencydomainwaveforms_b200 import engine, _lib
    items = [make_item(generator, n, dt=40.0) for n in ("plunge", "cfg1_like", "ecc_many")]
    N = max(it["N"] for it in items)
    val = 1.0 / (N * 40.0)
    h = _lib.get_handle()
    db = engine.DeviceBatch(engine.PackedBatch(items), h)
    hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)
    h.status()
    for i, it in enumerate(items):
        d1 = engine.DeviceBatch(engine.PackedBatch([it]), h)
        a, b, _ = engine.run_waveform(d1, N, val, mask_positive=True)
        assert torch_cuda.equal(a[0], hp[i]) and torch_cuda.equal(b[0], hc[i])   # bit-identical: no atomics, fixed order


def test_error_codes(generator, torch_cuda):
    from emri_frequencydomainwaveforms_b200 import engine, _lib
    from emri_frequencydomainwaveforms_b200.summation.fdinterp import FDInterpolatedModeSum
    it = make_item(generator, "cfg1_like")
    s = FDInterpolatedModeSum(pad_output=True, output_type="fd", odd_len=False)
    args = (it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"], it["M"], it["p"], it["e"])
    with pytest.raises(ValueError):
        s.sum(*args, f_arr=np.array([]))                                   # "Input f_arr has zero length."
    with pytest.raises(ValueError):
        s.sum(*args, f_arr=np.linspace(-1, 1, 10))                         # even length / no single zero
    with pytest.raises(ValueError):
        FDInterpolatedModeSum(output_type="td")
    h = _lib.get_handle()
    short = dict(it, t=it["t"][:3], teuk_modes=it["teuk_modes"][:3], Phi_phi=it["Phi_phi"][:3], Phi_r=it["Phi_r"][:3],
                 f_phi=it["f_phi"][:3], f_r=it["f_r"][:3])
    with pytest.raises(ValueError):
        engine.run_waveform(engine.DeviceBatch(engine.PackedBatch([short]), h), 1001, 1e-5)    # < 4 knots
    with pytest.raises(ValueError):
        engine.run_loglike(engine.DeviceBatch(engine.PackedBatch([it]), _lib.Handle()), it["N"], 1.0 / (it["N"] * 10.0))  # no data set


def test_all_modes_high_mode_count(generator, oracle_quad, torch_cuda):
    """config-4 flavour: every (l, m, n) of the amplitude basis (K = 3843 -> 15372 work-list slots, 61 compaction
    chunks per CTA), high eccentricity, negative-frequency harmonics (m f_phi + n f_r < 0) on the -f side."""
    it = generator.prepare(3e5, 30.0, 9.5, 0.68, 0.9, -np.pi / 2, dist=1.0, Phi_phi0=0.2, Phi_r0=2.0, T=0.03, dt=50.0,
                           mode_selection="all")
    from helpers import grid_size
    it["T"], it["dt"] = 0.03, 50.0
    it["N"] = grid_size(it["t"], 0.03, 50.0)
    assert it["teuk_modes"].shape[1] == 3843
    hp_o, hc_o, coeff_o, br_o, nbr_o = oracle_waveform(oracle_quad, it)
    s, out = _gpu_sum(it, torch_cuda)
    br_g = s.last_batch.branches_host()
    assert np.array_equal(s.last_batch.coeff_host(0), coeff_o)
    for key in ("dir", "ja", "jb", "start", "end", "xa", "xb", "Fa", "Fb"):
        assert np.array_equal(br_g[key], br_o[key]), key
    zero = (it["N"] - 1) // 2
    assert (br_o["end"][br_o["end"] >= br_o["start"]] < zero).any()          # some harmonics live at negative frequency
    assert rel_err(out[0], hp_o) <= TOL_BIN and rel_err(out[1], hc_o) <= TOL_BIN
    assert np.array_equal(out[0] != 0, hp_o != 0)


def test_full_size_one_year_grid(generator, oracle_quad, torch_cuda):
    """BASELINE size: T = 1 yr, dt = 10 s, N = 3 155 815, a plunging eps = 1e-2 system (~1e7 evaluations).
    Bins are independent, so the binary128 oracle is run on every 64th bin (an explicit, still uniform and
    symmetric f_arr built from the same doubles) and compared with the full-grid GPU result at the coincident
    frequencies; size-independent properties cover the rest."""
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200 import engine, _lib
    from helpers import grid_size
    M, mu, e0, T, dt = 8e5, 20.0, 0.4, 1.0, 10.0
    p0 = get_p_at_t(generator.inspiral_generator, 0.99 * T, [M, mu, 0.0, e0, 1.0], xtol=1e-9)
    it = generator.prepare(M, mu, p0, e0, 1.1, -np.pi / 2, dist=1.0, Phi_phi0=1.0, Phi_r0=2.0, T=T, dt=dt, eps=1e-2)
    it["T"], it["dt"] = T, dt
    N = it["N"] = grid_size(it["t"], T, dt)
    assert N == 3155815
    s, out = _gpu_sum(it, torch_cuda)                       # full two-sided output [2, N]
    zero, n = (N - 1) // 2, (N + 1) // 2
    # (1) against binary128 on the strided grid
    step = 64
    fpos = (np.arange(n, dtype=np.float64) * (1.0 / (N * dt)))[::step]
    Ns = 2 * len(fpos) - 1
    hp_o, hc_o, *_ = oracle_waveform(oracle_quad, it, N=Ns, fpos=fpos)
    sel = zero + step * np.arange(-(len(fpos) - 1), len(fpos))
    scale = np.max(np.abs(out[0]))
    assert np.max(np.abs(out[0][sel] - hp_o)) <= TOL_BIN * scale and np.max(np.abs(out[1][sel] - hc_o)) <= TOL_BIN * scale
    assert np.array_equal(out[0][sel] != 0, hp_o != 0)
    # (2) Hermitian symmetry of both polarisations, exactly (the mirror is written by the owning thread)
    assert np.array_equal(out[0][:zero], np.conj(out[0][zero + 1:][::-1])) and np.array_equal(out[1][:zero], np.conj(out[1][zero + 1:][::-1]))
    # (3) mask_positive == upper half; (4) linearity in the mode set (two disjoint halves add up to the whole)
    _, pos = _gpu_sum(it, torch_cuda, mask_positive=True)
    assert np.array_equal(pos, out[:, zero:])
    K = len(it["m_arr"])
    halves = []
    for idx in (np.arange(0, K, 2), np.arange(1, K, 2)):
        sub = dict(it, teuk_modes=np.ascontiguousarray(it["teuk_modes"][:, idx]), m_arr=it["m_arr"][idx], n_arr=it["n_arr"][idx],
                   ylms=np.concatenate([it["ylms"][idx], it["ylms"][K + idx]]))
        halves.append(_gpu_sum(sub, torch_cuda, mask_positive=True)[1])
    assert np.max(np.abs(halves[0] + halves[1] - pos)) <= 1e-14 * scale
    # (5) fused likelihood at full size: ll(injection) = 0 and <h|h> equals the materialised inner product
    h = _lib.get_handle()
    w = torch_cuda.full((2, n), 1.0e19, dtype=torch_cuda.float64, device=h.torch_device)
    dw = (torch_cuda.as_tensor(pos).to(h.torch_device) * w).contiguous()
    h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), w.data_ptr(), n))
    like = engine.run_loglike(engine.DeviceBatch(engine.PackedBatch([it]), h), N, 1.0 / (N * dt)).cpu().numpy()[0]
    hh = 4.0 * float((dw.abs() ** 2).sum().item())
    assert abs(like[0]) <= 1e-12 * hh and abs(like[2] - hh) <= 1e-12 * hh and abs(like[1] - hh) <= 1e-12 * hh


def test_fd_window_convolution(torch_cuda):
    """FDutils.get_convolution / get_fd_windowed (the step after the path when window_flag=1): the exact evaluation (rtol=0)
    == the reference's direct 'valid' convolution to rounding, the default (banded kernel, certified bound 1e-7) to 1e-7;
    == DFT(window * IDFT(signal)) (FDutils.py:83-85).  The banded kernel itself: tests/test_gpu_round2.py."""
    from scipy.signal import convolve
    from emri_frequencydomainwaveforms_b200.fdutils import get_convolution, get_fd_windowed, get_fd_waveform_fromFD
    rng = np.random.default_rng(8)
    n = 257
    a = rng.normal(size=n) + 1j * rng.normal(size=n)
    b = rng.normal(size=n) + 1j * rng.normal(size=n)
    ref = convolve(np.hstack((a[1:], a)), b, mode="valid") / len(b)       # FDutils.py:47 verbatim
    got = get_convolution(a, b).cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    window = np.hanning(n)
    sig = [np.fft.fftshift(a), np.fft.fftshift(b)]
    out = get_fd_windowed(sig, window, rtol=0)
    ref0 = convolve(np.hstack((np.conj(np.fft.fft(window))[1:], np.conj(np.fft.fft(window)))), sig[0], mode="valid") / n
    assert np.max(np.abs(out[0].cpu().numpy() - ref0)) <= 1e-12 * np.max(np.abs(ref0))
    out7 = get_fd_windowed(sig, window)
    assert np.linalg.norm(out7[0].cpu().numpy() - ref0) <= 1.5e-7 * np.linalg.norm(np.fft.fft(window)) * np.linalg.norm(sig[0]) / n
    assert get_fd_windowed(sig, None)[1] is sig[1]

    class Gen:                                                             # a generator returning [h+, hx] on the two-sided grid
        def __call__(self, *args, **kw):
            return [torch_cuda.as_tensor(sig[0]).cuda(), torch_cuda.as_tensor(sig[1]).cuda()]

    freq = np.fft.fftshift(np.fft.fftfreq(n, 10.0))
    ad = get_fd_waveform_fromFD(Gen(), freq >= 0.0, 10.0, window=window)
    ch = ad()
    assert ch[0].shape[0] == (n + 1) // 2 and np.max(np.abs(ch[0].cpu().numpy() - ref0[freq >= 0.0])) <= 1e-6 * np.max(np.abs(ref0))


def test_cyclic_tile_sharding_sums_to_full_likelihood(generator, torch_cuda):
    """emrifd_batch_sum_cyclic (multi-GPU frequency-bin sharding): the partial sums over the tile sets
    {r, r + W, ...}, r < W, add up to the full single-call likelihood; owned tiles of hp/hc equal the full waveform."""
    torch = torch_cuda
    from emri_frequencydomainwaveforms_b200 import _lib, engine
    h = _lib.get_handle()
    items = [make_item(generator, "plunge", dt=20.0), make_item(generator, "ecc_many", dt=20.0)]
    N = max(it["N"] for it in items)
    n = (N + 1) // 2
    val = 1.0 / (N * 20.0)
    db = engine.DeviceBatch(engine.PackedBatch(items), h)
    hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)
    w = torch.full((2, n), 2.0e19, dtype=torch.float64, device=h.torch_device)
    dw = (torch.stack([hp[0], hc[0]]) * w).contiguous()
    h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), w.data_ptr(), n))
    full = engine.run_loglike(db, N, val).cpu().numpy()
    tile = h.lib.emrifd_tile_bins()
    pb = db.pb
    flags = _lib.INCLUDE_MINUS_M | _lib.MASK_POSITIVE
    for W in (1, 3, 8):
        tot = np.zeros_like(full)
        for r in range(W):
            out = torch.zeros((pb.B, 3), dtype=torch.float64, device=h.torch_device)
            hp2 = torch.full((pb.B, n), complex(7.0, 7.0), dtype=torch.complex128, device=h.torch_device)
            hc2 = hp2.clone()
            pb.walkers["out_off"] = np.arange(pb.B, dtype=np.int64) * n
            h.check(h.lib.emrifd_batch_sum_cyclic(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(),
                                                  db.n.data_ptr(), db.ylm.data_ptr(), db.branches.data_ptr(), N, val, None, flags, r, W,
                                                  hp2.data_ptr(), hc2.data_ptr(), out.data_ptr()))
            tot += out.cpu().numpy()
            owned = ((torch.arange(n, device=h.torch_device) // tile) % W) == r
            # (not bit-for-bit: the full call may run the wide 6-bins-per-thread variant, the cyclic call always the base one;
            #  their Newton chains start from different bins, which moves results at the rounding level)
            for full_ch, own_ch in ((hp, hp2), (hc, hc2)):
                tol_ = 1e-12 * float(full_ch.abs().max())
                assert float((own_ch[:, owned] - full_ch[:, owned]).abs().max()) <= tol_
            assert torch.all(hp2[:, ~owned] == complex(7.0, 7.0))          # tiles of other ranks are not touched
        scale = np.abs(full[:, 2:3])
        assert np.all(np.abs(tot - full) <= 1e-12 * scale), (W, tot, full)
    # a rank beyond the number of tiles owns nothing
    out = torch.ones((pb.B, 3), dtype=torch.float64, device=h.torch_device)
    ntiles = (n + tile - 1) // tile
    h.check(h.lib.emrifd_batch_sum_cyclic(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(),
                                          db.n.data_ptr(), db.ylm.data_ptr(), db.branches.data_ptr(), N, val, None, flags, ntiles + 2, ntiles + 5,
                                          None, None, out.data_ptr()))
    assert torch.all(out == 0)


def test_queue_and_direct_launch_modes_agree_bit_for_bit(generator, torch_cuda):
    """The mode sum launches a direct (tile, walker) grid for small launches and persistent CTAs fed by a tile queue for
    large ones (> 4 waves).  Results must not depend on which: a 40-walker batch (queue) equals its single-walker calls
    (direct) bit for bit -- waveforms and likelihood sums -- and is reproducible run to run (queue order is not)."""
    torch = torch_cuda
    from emri_frequencydomainwaveforms_b200 import _lib, engine
    h = _lib.get_handle()
    base_items = [make_item(generator, "plunge", dt=20.0), make_item(generator, "ecc_many", dt=20.0), make_item(generator, "cfg1_like", dt=20.0)]
    items = [dict(base_items[i % 3], Phi_phi=base_items[i % 3]["Phi_phi"] + 0.1 * i) for i in range(40)]
    N = max(it["N"] for it in items)
    n = (N + 1) // 2
    val = 1.0 / (N * 20.0)
    tile = h.lib.emrifd_tile_bins()
    assert 40 * ((n + 2 * tile - 1) // (2 * tile)) > 4 * 2 * torch.cuda.get_device_properties(0).multi_processor_count   # -> queue path (any tile size)
    db0 = engine.DeviceBatch(engine.PackedBatch([items[0]]), h)
    hp0, hc0, _ = engine.run_waveform(db0, N, val, mask_positive=True)
    w = torch.full((2, n), 2.0e19, dtype=torch.float64, device=h.torch_device)
    dw = (torch.stack([hp0[0], hc0[0]]) * w).contiguous()
    h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), w.data_ptr(), n))
    dbB = engine.DeviceBatch(engine.PackedBatch(items), h)
    hpB, hcB, likeB = engine.run_waveform(dbB, N, val, mask_positive=True, like=True)
    hpB, hcB, likeB = hpB.clone(), hcB.clone(), likeB.clone()
    hpB2, hcB2, likeB2 = engine.run_waveform(dbB, N, val, mask_positive=True, like=True)
    assert torch.equal(hpB, hpB2) and torch.equal(hcB, hcB2) and torch.equal(likeB, likeB2)          # run-to-run
    only = engine.run_loglike(dbB, N, val)                                                             # no h written
    assert torch.equal(only, likeB)
    for i in (0, 1, 2, 13, 39):
        db1 = engine.DeviceBatch(engine.PackedBatch([items[i]]), h)
        hp1, hc1, like1 = engine.run_waveform(db1, N, val, mask_positive=True, like=True)
        assert torch.equal(hp1[0], hpB[i]) and torch.equal(hc1[0], hcB[i]) and torch.equal(like1[0], likeB[i]), i
    h.status()


def test_minimal_shapes_and_empty_support(generator, oracle_quad, torch_cuda):
    """Edge cases: one mode and four knots (the not-a-knot minimum); a grid whose band misses the signal entirely
    (h = 0 everywhere, ll = -2 sum|d~|^2 from the per-tile table); a non-uniform explicit f_arr."""
    torch = torch_cuda
    from emri_frequencydomainwaveforms_b200 import _lib, engine
    h = _lib.get_handle()
    it = make_item(generator, "cfg1_like")
    K = len(it["m_arr"])
    k = int(np.argmax(np.abs(it["teuk_modes"][0] * it["ylms"][:K])))
    knots = np.array([0, len(it["t"]) // 3, 2 * len(it["t"]) // 3, len(it["t"]) - 1])
    one = dict(it, t=it["t"][knots], p=it["p"][knots], e=it["e"][knots], Phi_phi=it["Phi_phi"][knots], Phi_r=it["Phi_r"][knots],
               f_phi=it["f_phi"][knots], f_r=it["f_r"][knots], teuk_modes=np.ascontiguousarray(it["teuk_modes"][knots][:, [k]]),
               m_arr=it["m_arr"][[k]], n_arr=it["n_arr"][[k]], ylms=it["ylms"][[k, K + k]])
    hp_o, hc_o, coeff_o, br_o, _ = oracle_waveform(oracle_quad, one)
    s, out = _gpu_sum(one, torch)
    assert np.array_equal(s.last_batch.coeff_host(0), coeff_o)
    assert rel_err(out[0], hp_o) <= TOL_BIN and rel_err(out[1], hc_o) <= TOL_BIN and np.array_equal(out[0] != 0, hp_o != 0)
    # a band below the signal: f_max of the grid is under the lowest harmonic frequency of the kept modes
    live = (it["m_arr"] != 0) | (it["n_arr"] != 0)                      # (l, 0, 0) modes have f_mn = 0 and no stationary point
    f_lo = np.min(np.abs(np.outer(it["f_phi"], it["m_arr"][live]) + np.outer(it["f_r"], it["n_arr"][live])))
    assert f_lo > 0
    n = 5001
    fpos = np.linspace(0.0, 0.5 * f_lo, n)
    newf = np.hstack((-fpos[::-1][:-1], fpos))
    s2, out2 = _gpu_sum(it, torch, f_arr=newf)
    assert out2.shape == (2, 2 * n - 1) and not out2.any()
    rng = np.random.default_rng(1)
    d = rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))
    w = np.full((2, n), 3.0)
    d_dev, w_dev = torch.from_numpy((d * w).view(np.float64)).cuda(), torch.from_numpy(w).cuda()
    h.check(h.lib.emrifd_set_data(h.h, d_dev.data_ptr(), w_dev.data_ptr(), n))
    fpos_dev = torch.from_numpy(fpos).cuda()
    like = engine.run_loglike(engine.DeviceBatch(engine.PackedBatch([it]), h), 2 * n - 1, 0.0, fpos_dev).cpu().numpy()
    ref = -2.0 * np.sum(np.abs(d * w) ** 2)
    assert abs(like[0, 0] - ref) <= 1e-12 * abs(ref) and like[0, 1] == 0 and like[0, 2] == 0
    # non-uniform (geometric) positive grid, as a user-supplied f_arr may be
    hi = 1.02 * np.max(np.abs(np.outer(it["f_phi"], it["m_arr"]) + np.outer(it["f_r"], it["n_arr"])))
    fpos3 = np.concatenate([[0.0], np.geomspace(0.2 * f_lo, hi, 4000)])
    newf3 = np.hstack((-fpos3[::-1][:-1], fpos3))
    s3, out3 = _gpu_sum(it, torch, f_arr=newf3)
    hp3, hc3, *_ = oracle_waveform(oracle_quad, it, N=len(newf3), fpos=fpos3)
    assert rel_err(out3[0], hp3) <= TOL_BIN and rel_err(out3[1], hc3) <= TOL_BIN and np.array_equal(out3[0] != 0, hp3 != 0)


def test_td_to_fd_utilities(torch_cuda):
    """FDutils.get_fft_td_windowed / get_fd_waveform_fromTD (FDutils.py:49-64,142-178) against numpy's FFT."""
    from emri_frequencydomainwaveforms_b200.fdutils import get_fft_td_windowed, get_fd_waveform_fromTD
    rng = np.random.default_rng(3)
    n, dt = 1001, 10.0
    hp, hc = rng.normal(size=n), rng.normal(size=n)
    window = np.hanning(n)
    got = get_fft_td_windowed([hp, hc], window, dt)
    ref = [np.fft.fftshift(np.fft.fft(x * window)) * dt for x in (hp, hc)]            # FDutils.py:62-63 verbatim
    for g, r in zip(got, ref):
        assert np.max(np.abs(g.cpu().numpy() - r)) <= 1e-12 * np.max(np.abs(r))
    freq = np.fft.fftshift(np.fft.fftfreq(n, dt))
    pos = freq >= 0.0
    nz = np.zeros(int(pos.sum()), dtype=bool)
    nz[5:200] = True
    ad = get_fd_waveform_fromTD(lambda *a, **k: hp - 1j * hc, pos, dt, non_zero_mask=nz, window=window)
    ch = ad()
    exp0 = np.where(nz, ref[0][pos], 0.0)
    assert ch[0].shape[0] == (n + 1) // 2 and np.max(np.abs(ch[0].cpu().numpy() - exp0)) <= 1e-12 * np.max(np.abs(exp0))
    assert np.all(ch[1].cpu().numpy()[~nz] == 0)


def test_long_trajectory_paths(generator, oracle_quad, torch_cuda):
    """L = 700 knots: the spline kernel's non-tiled variant (forward-sweep intermediates parked in the output), fewer
    modes per CTA in the segmentation kernel and a 118 KB track staging area in the mode-sum kernel."""
    from scipy.interpolate import CubicSpline
    from emri_frequencydomainwaveforms_b200.utils.utility import fundamental_frequencies_hz
    it = make_item(generator, "plunge", dt=40.0)
    t0 = it["t"]
    # refine the knot vector (geometric mix keeps the clustering near the end), re-evaluate everything on it
    u = np.linspace(0.0, 1.0, 700)
    t = np.interp(u, np.linspace(0.0, 1.0, len(t0)), t0)
    p, e = CubicSpline(t0, it["p"])(t), np.clip(CubicSpline(t0, it["e"])(t), 0.0, None)
    K = len(it["m_arr"])
    amp = generator.amplitude_generator
    idx = [int(np.where((amp.l_arr == l) & (amp.m_arr == m) & (amp.n_arr == n))[0][0]) for l, m, n in zip(it["l_arr"], it["m_arr"], it["n_arr"])]
    f_phi, f_r = fundamental_frequencies_hz(p, e, it["M"])
    long_it = dict(it, t=t, p=p, e=e, teuk_modes=np.ascontiguousarray(amp(p, e)[:, idx]), Phi_phi=CubicSpline(t0, it["Phi_phi"])(t),
                   Phi_r=CubicSpline(t0, it["Phi_r"])(t), f_phi=f_phi, f_r=f_r)
    assert long_it["teuk_modes"].shape == (700, K)
    hp_o, hc_o, coeff_o, br_o, nbr_o = oracle_waveform(oracle_quad, long_it)
    s, out = _gpu_sum(long_it, torch_cuda)
    assert np.array_equal(s.last_batch.coeff_host(0), coeff_o)
    br_g = s.last_batch.branches_host()
    for key in ("dir", "ja", "jb", "start", "end", "xa", "xb", "Fa", "Fb"):
        assert np.array_equal(br_g[key], br_o[key]), key
    assert rel_err(out[0], hp_o) <= TOL_BIN and rel_err(out[1], hc_o) <= TOL_BIN
    # beyond EMRIFD_MAX_KNOTS the C-ABI refuses with an error code (-> ValueError), it does not crash
    big = dict(long_it, t=np.linspace(0, 1e7, 1100), p=np.linspace(10, 9, 1100), e=np.full(1100, 0.2), teuk_modes=np.ones((1100, K), dtype=complex),
               Phi_phi=np.linspace(0, 1e4, 1100), Phi_r=np.linspace(0, 7e3, 1100))
    with pytest.raises(ValueError):
        _gpu_sum(big, torch_cuda)


def test_bench_line_contract(torch_cuda):
    """`python bench.py` prints ONE JSON line with the keys the driver reads (metric, value, e2e, roofline, clocks, gpu_launches)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "2", "--warmup", "3", "--batch", "8", "--no-cpu-baseline"], capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0 and d["dtype"] == "f64" and d["gpu_launches"] >= 2 * 5
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "fp64") and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    assert d["e2e_from_parameters"]["value"] > 0 and d["e2e_from_parameters"]["finite"]
    assert d["cfg1"]["walkers_per_s"] > 0 and 0 < d["cfg1"]["hbm_frac"] < 1.3
    c4 = d["cfg4_binsharded"]
    assert c4["modes"] == 3843 and c4["solves"] < c4["mode_evals"] / 4 and c4["ms_per_likelihood"] > 0 and abs(c4["ll_1rank"]) <= 1e-10 * c4["hh"]
    assert r["executed_flops_per_solve"] > 100 and r["solves_per_launch"] > 0

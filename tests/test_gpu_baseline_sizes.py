"""GPU parity at the sizes BASELINE.json's configs name (through the C-ABI, against the binary128 oracle).

configs[0]: M=1e6, mu=10, p0=12, e0=0.35, T=1 yr, dt=10 s, eps=1e-2 -> N = 3 155 815, every 16th bin against binary128.
configs[2]: T=2 yr, same system, the down-sampled f_arr of emri_pe.py:333-349 (downsample=100), 16 walkers through the fused
            likelihood plug-in, every bin against binary128.
configs[3]: T=4 yr, e0=0.7, all 3843 (l, m, n) modes, N = 12 623 261, every 4096th bin against binary128.
Bins are independent, so a strided explicit grid built from the same doubles is an exact sub-sample of the full grid.
Tolerances (north_star): per-bin error <= 1e-10 of max|h|, mismatch <= 1e-12, identical support.
"""
import numpy as np
import pytest

from helpers import grid_size, oracle_waveform

pytestmark = pytest.mark.gpu

TOL_BIN = 1e-10
TOL_MISMATCH = 1e-12


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _psd(f):
    """LISA_Alloc_Sh.txt spline (FDutils.py:4-5,21-33) on the host."""
    import os
    from scipy.interpolate import CubicSpline
    from emri_frequencydomainwaveforms_b200 import fdutils
    tab = np.load(os.path.join(os.path.dirname(fdutils.__file__), "data", "lisa_alloc_sh.npy"))
    return CubicSpline(tab[:, 0], tab[:, 1])(f)


def mismatch(a, b, f):
    """1 - <a|b>/sqrt(<a|a><b|b>) with the lisatools inner product (diagnostic.py:95-139) summed over the channels of
    a, b [nch, n] (check_mode_by_mode.py:299-306)."""
    df = np.empty_like(f)
    df[1:] = np.diff(f)
    df[0] = df[1]
    w = (df / _psd(f))[None, :]
    ip = lambda x, y: 4.0 * np.sum(w * (np.conj(x) * y).real)
    return 1.0 - ip(a, b) / np.sqrt(ip(a, a) * ip(b, b))


def _strided_check(it, out_pos, N, dt, step, orc):
    """out_pos: GPU h+, hx on f >= 0 of the full implicit grid; compares every `step`-th bin with binary128."""
    n = (N + 1) // 2
    fpos = (np.arange(n, dtype=np.float64) * (1.0 / (N * dt)))[::step]
    Ns = 2 * len(fpos) - 1
    hp_o, hc_o, *_ = oracle_waveform(orc, it, N=Ns, fpos=fpos)
    hp_o, hc_o = hp_o[len(fpos) - 1:], hc_o[len(fpos) - 1:]
    g0, g1 = out_pos[0][::step], out_pos[1][::step]
    scale = max(np.max(np.abs(out_pos[0])), np.max(np.abs(out_pos[1])))
    err = max(np.max(np.abs(g0 - hp_o)), np.max(np.abs(g1 - hc_o))) / scale
    assert err <= TOL_BIN, err
    assert np.array_equal(g0 != 0, hp_o != 0) and np.array_equal(g1 != 0, hc_o != 0)      # identical support
    sup = hp_o != 0
    assert sup.sum() > 50
    mm = mismatch(np.stack([g0, g1]), np.stack([hp_o, hc_o]), fpos)
    assert abs(mm) <= TOL_MISMATCH, mm
    return err, mm, int(sup.sum())


def test_config0_one_year_exact_parameters(generator, oracle_quad, torch_cuda):
    """BASELINE.json configs[0] (the reference's own CPU-runnable case): sparse support, HBM-bound regime."""
    from emri_frequencydomainwaveforms_b200.summation.fdinterp import FDInterpolatedModeSum
    M, mu, p0, e0, T, dt = 1e6, 10.0, 12.0, 0.35, 1.0, 10.0
    it = generator.prepare(M, mu, p0, e0, np.pi / 3, -np.pi / 2, dist=1.0, Phi_phi0=0.7, Phi_r0=2.1, T=T, dt=dt, eps=1e-2)
    N = grid_size(it["t"], T, dt)
    assert N == 3155815
    s = FDInterpolatedModeSum(pad_output=True, output_type="fd", odd_len=True)
    out = s(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"], M, it["p"], it["e"],
            T=T, dt=dt, scale=it["scale"]).cpu().numpy()
    assert out.shape == (2, N)
    zero = (N - 1) // 2
    err, mm, nsup = _strided_check(it, out[:, zero:], N, dt, 16, oracle_quad)
    # Hermitian mirror written by the owning thread, exactly
    assert np.array_equal(out[0][:zero], np.conj(out[0][zero + 1:][::-1])) and np.array_equal(out[1][:zero], np.conj(out[1][zero + 1:][::-1]))
    print(f"configs[0]: per-bin err {err:.2e}, mismatch {mm:.2e}, {nsup} support bins checked")


def test_config2_two_year_downsampled_likelihood(generator, oracle_quad, torch_cuda):
    """BASELINE.json configs[2]: emri_pe.py-style batched likelihood, T = 2 yr, downsample = 100 (emri_pe.py:333-361), 16 walkers."""
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel, Likelihood
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    from emri_frequencydomainwaveforms_b200.waveform import ssb_transform_batch
    from emri_frequencydomainwaveforms_b200.utils.constants import MRSUN_SI, Gpc
    T, dt, eps = 2.0, 10.0, 1e-2
    gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True),
                               return_list=True)
    inj = np.array([1e6, 10.0, 0.0, 12.0, 0.35, 1.0, 1.0, np.pi / 3, np.pi / 3, np.pi / 3, np.pi / 3, np.pi / 3, 0.0, np.pi / 3])
    kw = dict(T=T, dt=dt, eps=eps)
    # full grid first: frequency attribute, non-zero mask (emri_pe.py:237-245,333-339)
    full = gen(*inj, mask_positive=True, **kw)
    N = gen.waveform_generator.create_waveform.frequency.shape[0]
    assert N == 6311631
    fixed_freq = gen.waveform_generator.create_waveform.frequency.cpu().numpy()[(N - 1) // 2:]
    non_zero = (full[0].abs() > 0).cpu().numpy()
    end_f = fixed_freq[non_zero].max()
    num = int(non_zero.sum() / 100)
    p_freq = np.linspace(0.0, end_f * 1.01, num=num)
    newfreq = np.hstack((-p_freq[::-1][:-1], p_freq))
    assert 1000 < num < 20000
    # (1) the down-sampled waveform against binary128 on every bin
    base = gen.waveform_generator
    theta, phi, c2, s2 = (float(x[0]) for x in ssb_transform_batch([inj[7]], [inj[8]], [inj[9]], [inj[10]]))
    it = base.prepare(inj[0], inj[1], inj[3], inj[4], theta, phi, dist=inj[6], Phi_phi0=inj[11], Phi_r0=inj[13], T=T, dt=dt, eps=eps)
    # (host producers feed the GPU sum and the oracle the very same doubles; the device producers are compared below)
    ds = base.create_waveform(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"], inj[0],
                              it["p"], it["e"], T=T, dt=dt, f_arr=newfreq, mask_positive=True, scale=it["scale"], cos2psi=c2, sin2psi=s2)
    got = np.stack([ds[0].cpu().numpy(), ds[1].cpu().numpy()])
    hp_o, hc_o, *_ = oracle_waveform(oracle_quad, it, N=len(newfreq), fpos=p_freq, cos2psi=c2, sin2psi=s2)
    ref = np.stack([hp_o[num - 1:], hc_o[num - 1:]])
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(got - ref)) <= TOL_BIN * scale
    assert np.array_equal(got != 0, ref != 0)
    assert abs(mismatch(got, ref, p_freq)) <= TOL_MISMATCH
    ds_dev = gen(*inj, f_arr=newfreq, mask_positive=True, **kw)     # public call: device producers (Ylm, mode selection)
    got_dev = np.stack([ds_dev[0].cpu().numpy(), ds_dev[1].cpu().numpy()])
    assert np.max(np.abs(got_dev - ref)) <= 1e-9 * scale and abs(mismatch(got_dev, ref, p_freq)) <= 1e-11
    # (2) 16 walkers through Likelihood(..., fill_data_noise=True) -> FDTemplateModel.get_ll against the oracle
    model = FDTemplateModel(gen, f_arr=newfreq, producers="host")
    like = Likelihood(model, 2, f_arr=p_freq, fill_data_noise=True)
    like.inject_signal(data_stream=[ds[0], ds[1]], noise_fn=get_sensitivity, noise_kwargs={})
    rng = np.random.default_rng(2601996)     # emri_pe.py:65-66
    params = np.tile(inj, (16, 1))
    params[:, 0] *= 1.0 + 1e-6 * rng.normal(size=16)
    params[:, 3] += 1e-5 * rng.normal(size=16)
    params[:, 4] += 1e-5 * rng.normal(size=16)
    params[:, 11] += 1e-2 * rng.normal(size=16)
    params[:, 13] += 1e-2 * rng.normal(size=16)
    params[0] = inj
    ll = like(params, **kw)
    assert ll.shape == (16,) and np.all(np.isfinite(ll))
    dd = 4.0 * float(np.sum(np.abs(like.injection_channels) ** 2))
    assert abs(ll[0]) <= 1e-12 * dd
    for i in (1, 7, 15):
        p = params[i]
        iti = base.prepare(p[0], p[1], p[3], p[4], theta, phi, dist=p[6], Phi_phi0=p[11], Phi_r0=p[13], T=T, dt=dt, eps=eps)
        hp_i, hc_i, *_ = oracle_waveform(oracle_quad, iti, N=len(newfreq), fpos=p_freq, cos2psi=c2, sin2psi=s2)
        ll_o = oracle_quad.loglike(like.injection_channels, np.stack([hp_i[num - 1:], hc_i[num - 1:]]), like.noise_factor)[0]
        assert abs(ll[i] - ll_o) <= 1e-10 * dd, (i, ll[i], ll_o)
    # the default (device-producer) plug-in agrees with the host-producer one
    like_dev = Likelihood(FDTemplateModel(gen, f_arr=newfreq), 2, f_arr=p_freq, fill_data_noise=True)
    like_dev.inject_signal(data_stream=[ds[0], ds[1]], noise_fn=get_sensitivity, noise_kwargs={})
    ll_dev = like_dev(params, **kw)
    assert np.max(np.abs(ll_dev - ll)) <= 1e-8 * dd
    print(f"configs[2]: {num} frequencies, ll = {ll[:4]}")


def test_config3_four_year_all_modes(generator, oracle_quad, torch_cuda):
    """BASELINE.json configs[3]: 4 yr, e0 = 0.7, every (l, m, n) of the basis (3843 modes -> 671 (m, n) groups), 6.3e6 bins."""
    from emri_frequencydomainwaveforms_b200 import engine, _lib
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    M, mu, e0, T, dt = 1e6, 10.0, 0.7, 4.0, 10.0
    p0 = get_p_at_t(generator.inspiral_generator, T * 0.99, [M, mu, 0.0, e0, 1.0], xtol=1e-9, bounds=[7.2 + 2 * e0 + 0.05, 16.0 + 2 * e0])
    it = generator.prepare(M, mu, p0, e0, 1.0, -np.pi / 2, dist=1.0, Phi_phi0=0.3, Phi_r0=1.3, T=T, dt=dt, mode_selection="all")
    assert it["teuk_modes"].shape[1] == 3843
    N = grid_size(it["t"], T, dt)
    assert N == 12623261
    h = _lib.get_handle()
    db = engine.DeviceBatch(engine.PackedBatch([it]), h)
    hp, hc, _ = engine.run_waveform(db, N, 1.0 / (N * dt), mask_positive=True)
    h.status()
    out = np.stack([hp[0].cpu().numpy(), hc[0].cpu().numpy()])
    gev, ev = int(engine.group_evaluations(db)[0]), int(np.where(db.branches_host()["end"] >= db.branches_host()["start"],
                                                                  db.branches_host()["end"] - db.branches_host()["start"] + 1, 0).sum())
    assert ev > 1e9 and gev < ev / 4          # one stationary point per (m, n) group: > 4x fewer solves than per (l, m, n)
    err, mm, nsup = _strided_check(it, out, N, dt, 4096, oracle_quad)
    print(f"configs[3]: {ev:.3e} per-mode evaluations, {gev:.3e} solved; per-bin err {err:.2e}, mismatch {mm:.2e}, {nsup} support bins checked")

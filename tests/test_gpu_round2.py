"""GPU tests of the round-2 additions: per-walker error isolation, the FEW-compatible K_{1/3} switch, bin slices whose last
tile is cut short (with data that are non-zero everywhere), two models sharing the per-device handle, the un-windowed
``get_fd_waveform_fromFD`` branch, and the cell-26 golden through the CUDA path."""
import os

import numpy as np
import pytest

from helpers import make_item, oracle_waveform, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _noise_data(torch, h, n, seed=5):
    """Whitened 'data' that are non-zero in every bin (noise-like), and a smooth noise factor."""
    rng = np.random.default_rng(seed)
    w = np.stack([1.0e19 * (1.0 + 0.3 * np.sin(np.arange(n) * 1e-3)), 1.3e19 * (1.0 + 0.2 * np.cos(np.arange(n) * 7e-4))])
    d = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))) * 3.0
    dd = torch.from_numpy(np.ascontiguousarray(d).view(np.float64)).to(h.torch_device)
    ww = torch.from_numpy(np.ascontiguousarray(w)).to(h.torch_device)
    h.check(h.lib.emrifd_set_data(h.h, dd.data_ptr(), ww.data_ptr(), n))
    return d, w, (dd, ww)


def test_per_walker_error_isolation(generator, torch_cuda):
    """One walker with non-monotone knots and one whose harmonics have more than EMRIFD_MAX_BRANCHES monotone branches inside a
    16-walker batch: those two come back as h = 0 / ll = NaN with their status codes, every other walker is bit-identical to a
    clean run, and nothing raises (the reference contract is per walker: Eryn/eryn/moves/red_blue.py:282-284,
    check_mode_by_mode.py:328-330)."""
    torch = torch_cuda
    from emri_frequencydomainwaveforms_b200 import _lib, engine
    h = _lib.get_handle()
    base = [make_item(generator, "plunge", dt=40.0), make_item(generator, "cfg1_like", dt=40.0), make_item(generator, "ecc_many", dt=40.0)]
    items = [dict(base[i % 3], Phi_phi=base[i % 3]["Phi_phi"] + 0.05 * i) for i in range(16)]
    N = max(it["N"] for it in items)
    n = (N + 1) // 2
    val = 1.0 / (N * 40.0)
    keep = _noise_data(torch, h, n)      # (the handle keeps only the pointers: hold the tensors)
    clean = engine.DeviceBatch(engine.PackedBatch(items), h)
    hp_c, hc_c, ll_c = engine.run_waveform(clean, N, val, mask_positive=True, like=True)
    hp_c, hc_c, ll_c = hp_c.clone(), hc_c.clone(), ll_c.clone()
    assert not h.walker_status(16).any()
    bad = [dict(it) for it in items]
    t_bad = bad[3]["t"].copy()
    t_bad[5], t_bad[6] = t_bad[6], t_bad[5]                                  # knots not strictly increasing
    bad[3]["t"] = t_bad
    L = len(bad[7]["t"])
    wig = 1.0 + 2e-3 * np.sin(np.linspace(0.0, 9.0 * np.pi, L))             # f_phi, f_r oscillate: > 4 monotone branches per harmonic
    bad[7]["f_phi"], bad[7]["f_r"] = bad[7]["f_phi"] * wig, bad[7]["f_r"] * wig
    db = engine.DeviceBatch(engine.PackedBatch(bad), h)
    hp, hc, ll = engine.run_waveform(db, N, val, mask_positive=True, like=True)      # does not raise
    st = h.walker_status(16)
    assert st[3] == -3 and st[7] == -4 and not st[[i for i in range(16) if i not in (3, 7)]].any()
    good = [i for i in range(16) if i not in (3, 7)]
    assert torch.equal(hp[good], hp_c[good]) and torch.equal(hc[good], hc_c[good]) and torch.equal(ll[good], ll_c[good])
    assert torch.all(hp[[3, 7]] == 0) and torch.all(hc[[3, 7]] == 0) and torch.isnan(ll[[3, 7]]).all()
    # likelihood-only launch (no h written) and the host-buffer e2e call: same contract, return code 0
    only = engine.run_loglike(db, N, val)
    assert torch.equal(only[good], ll_c[good]) and torch.isnan(only[[3, 7]]).all()
    host = engine.run_loglike_host(engine.PackedBatch(bad), h, N, val)
    assert np.array_equal(host[good], ll_c.cpu().numpy()[good]) and np.isnan(host[[3, 7]]).all()
    assert list(h.walker_status(16)[[3, 7]]) == [-3, -4]
    # the single-waveform surface keeps FEW's behaviour: a bad trajectory raises ValueError
    with pytest.raises(ValueError):
        one = engine.DeviceBatch(engine.PackedBatch([bad[3]]), h)
        engine.run_waveform(one, N, val, mask_positive=True)
        h.status()
    # and the next clean batch on the same handle is unaffected
    hp2, hc2, ll2 = engine.run_waveform(clean, N, val, mask_positive=True, like=True)
    assert torch.equal(hp2, hp_c) and torch.equal(ll2, ll_c) and not h.walker_status(16).any()


def test_k13_few_mode_matches_oracle_few_mode(generator, oracle_quad, torch_cuda):
    """emrifd_set_k13_mode(EMRIFD_K13_FEW): the kernel's FEW-compatible evaluation (14-term series below |X| = 7, 9-term
    asymptotic above) against the oracle in the same mode, <= 1e-10; against the exact mode the waveform moves by a small but
    non-zero amount, only where X = 2 pi fdot^3/(3 fddot^2) is O(10) (turnovers, late inspiral)."""
    from emri_frequencydomainwaveforms_b200 import _lib
    from test_gpu_parity import _gpu_sum
    it = make_item(generator, "plunge")
    h = _lib.get_handle()
    _, exact = _gpu_sum(it, torch_cuda)
    try:
        h.set_k13_mode("few")
        oracle_quad.set_k13_mode("few")
        _, few = _gpu_sum(it, torch_cuda)
        hp_o, hc_o, *_ = oracle_waveform(oracle_quad, it)
    finally:
        h.set_k13_mode("exact")
        oracle_quad.set_k13_mode("exact")
    assert rel_err(few[0], hp_o) <= 1e-10 and rel_err(few[1], hc_o) <= 1e-10
    d = np.abs(few[0] - exact[0]) / np.max(np.abs(exact[0]))
    assert 1e-12 < d.max() <= 1e-5
    assert np.count_nonzero(d > 1e-11) < 0.5 * np.count_nonzero(exact[0])
    _, again = _gpu_sum(it, torch_cuda)
    assert np.array_equal(again, exact)                               # the switch is per handle and was restored
    with pytest.raises(ValueError):
        h.set_k13_mode("fast")


def test_bin_slices_with_truncated_last_tile(generator, torch_cuda):
    """Contiguous bin slices whose ends are not tile-aligned (distributed.balanced_bin_slices aligns STARTS to 1024 bins, while a
    slice starting on a 1536-bin boundary runs 1536-bin tiles): with data that are non-zero in EVERY bin the partial sums of the
    slices must add up to the full likelihood -- a truncated last tile may not take the whole-tile sum |d~|^2 from the
    precomputed table."""
    torch = torch_cuda
    from emri_frequencydomainwaveforms_b200 import _lib, engine
    h = _lib.get_handle()
    items = [make_item(generator, "plunge", dt=20.0), make_item(generator, "cfg1_like", dt=20.0)]
    N = max(it["N"] for it in items)
    n = (N + 1) // 2
    val = 1.0 / (N * 20.0)
    d, w, keep = _noise_data(torch, h, n, seed=11)
    db = engine.DeviceBatch(engine.PackedBatch(items), h)
    full = engine.run_loglike(db, N, val).cpu().numpy()
    dd = 4.0 * np.sum(np.abs(d) ** 2)
    assert np.all(full[:, 0] < -0.4 * dd)                              # the data dominate: every bin matters
    pb = db.pb
    flags = _lib.INCLUDE_MINUS_M | _lib.MASK_POSITIVE
    assert n > 12 * 1536
    for edges in ([0, 1024 * 5, n], [0, 1536 * 2 + 1024, 1536 * 6 + 512 + 1024 * 3, n], [0, 1024 * 7 + 3, 1024 * 20, n], [0, 777, n]):
        tot = np.zeros_like(full)
        for lo, hi in zip(edges[:-1], edges[1:]):
            out = torch.zeros((pb.B, 3), dtype=torch.float64, device=h.torch_device)
            h.check(h.lib.emrifd_batch_sum(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(),
                                           db.n.data_ptr(), db.ylm.data_ptr(), db.branches.data_ptr(), N, val, None, flags, lo, hi - lo,
                                           None, None, out.data_ptr()))
            tot += out.cpu().numpy()
        assert np.all(np.abs(tot - full) <= 1e-12 * dd), (edges, tot, full)


def test_two_models_share_the_device_handle(torch_cuda):
    """The injected data live on the per-device handle.  Two FDTemplateModels with different data (and a like-here Likelihood in
    between) must each evaluate against THEIR data, whatever ran last on the handle."""
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel, Likelihood
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    kw = dict(T=0.05, dt=20.0, eps=1e-2)
    gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True), return_list=True)
    pA = np.array([1e6, 10.0, 0.0, 12.0, 0.35, 1.0, 1.0, 0.8, 0.3, 1.1, 0.5, 0.4, 0.0, 1.0])
    pB = pA.copy()
    pB[11] += 0.3
    mA, mB = FDTemplateModel(gen), FDTemplateModel(gen)
    sA, sB = mA(*pA, **kw), mB(*pB, **kw)
    n = sA[0].shape[0]
    f_arr = np.arange(n) / ((2 * n - 1) * kw["dt"])
    likes = []
    for m_, s_ in ((mA, sA), (mB, sB)):
        lk = Likelihood(m_, 2, f_arr=f_arr, fill_data_noise=True)
        lk.inject_signal(data_stream=[s_[0], s_[1]], noise_fn=get_sensitivity, noise_kwargs={})
        likes.append(lk)
    P = np.stack([pA, pB])
    a1, b1 = likes[0](P, **kw), likes[1](P, **kw)
    dd = 4.0 * float(np.sum(np.abs(likes[0].injection_channels) ** 2))
    assert abs(a1[0]) <= 1e-10 * dd and a1[1] < -1e-6 * dd and abs(b1[1]) <= 1e-10 * dd and b1[0] < -1e-6 * dd
    # interleave, including a likelihood that is evaluated "here" (emrifd_loglike on materialised templates)
    here = Likelihood(lambda *p, **k: mA(*p, **k), 2, f_arr=f_arr)
    here.inject_signal(data_stream=[sB[0], sB[1]], noise_fn=get_sensitivity, noise_kwargs={})
    for _ in range(2):
        assert np.array_equal(likes[0](P, **kw), a1)
        assert abs(here(P[:1], **kw)[0] - b1[0]) <= 1e-9 * dd
        assert np.array_equal(likes[1](P, **kw), b1)
        assert np.array_equal(likes[0](P, **kw), a1)
    with pytest.raises(ValueError):       # complex h+ - i hx cannot be split back into two complex channels
        FDTemplateModel(GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True),
                                             return_list=False))(*pA, **kw)


def test_fd_waveform_fromFD_without_window(torch_cuda):
    """get_fd_waveform_fromFD with window=None and a non_zero_mask -- the branch emri_pe.py:266 uses (FDutils.py:131-139):
    f >= 0 kept, zero outside the mask; for a mask that is not the upper half the boolean gather is used."""
    torch = torch_cuda
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    from emri_frequencydomainwaveforms_b200.fdutils import get_fd_waveform_fromFD
    kw = dict(T=0.05, dt=20.0, eps=1e-2)
    gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True), return_list=True)
    p = [1e6, 10.0, 0.0, 12.0, 0.35, 1.0, 1.0, 0.8, 0.3, 1.1, 0.5, 0.4, 0.0, 1.0]
    full = gen(*p, **kw)
    freq = gen.waveform_generator.create_waveform.frequency
    pos = freq >= 0.0                                                     # emri_pe.py:239
    hpos = [full[0][pos], full[1][pos]]
    nz = (hpos[0].abs() > 0)
    nz[: int(nz.nonzero()[0].item()) + 40] = False                       # a mask that drops part of the support (emri_pe.py:244-245 keeps it all)
    out = get_fd_waveform_fromFD(gen, pos, kw["dt"], non_zero_mask=nz)(*p, **kw)
    assert out[0].shape == hpos[0].shape
    assert torch.equal(out[0][nz], hpos[0][nz]) and torch.equal(out[1][nz], hpos[1][nz])
    assert torch.all(out[0][~nz] == 0) and torch.all(out[1][~nz] == 0) and (hpos[0][~nz].abs() > 0).any()
    out2 = get_fd_waveform_fromFD(gen, pos, kw["dt"])(*p, **kw)           # no mask at all: passthrough of the f >= 0 half
    assert torch.equal(out2[0], hpos[0]) and torch.equal(out2[1], hpos[1])
    odd = pos.clone()
    odd[-5:] = False                                                      # not the upper half: boolean gather path
    out3 = get_fd_waveform_fromFD(gen, odd, kw["dt"])(*p, **kw)
    assert torch.equal(out3[0], full[0][odd]) and torch.equal(out3[1], full[1][odd])


def test_cell26_golden_through_the_cuda_path(torch_cuda):
    """The reference notebook's cell 26 executed verbatim (tests/golden/make_cell26_golden.py) against the CUDA path: identical
    support and conventions, mismatch <= 1e-8 (the cell's own approximations; measured 1.5e-13 / 8.2e-10 / 1.7e-14)."""
    from emri_frequencydomainwaveforms_b200.summation.fdinterp import FDInterpolatedModeSum
    g = np.load(os.path.join(GOLD, "cell26_golden.npz"))
    s = FDInterpolatedModeSum(pad_output=True, output_type="fd", odd_len=True)
    for name in g["names"]:
        M, mu, T, dt, N, scale = g[f"{name}.params"]
        l, m, n = (int(x) for x in g[f"{name}.lmn"])
        get = lambda k: g[f"{name}.{k}"]
        f_arr = np.fft.fftshift(np.fft.fftfreq(int(N), dt))
        out = s.sum(get("t"), get("teuk_modes"), get("ylms"), get("Phi_phi"), get("Phi_r"), np.array([m]), np.array([n]), M, get("p"), get("e"),
                    dt=dt, f_arr=f_arr, scale=scale).cpu().numpy()
        W_gpu = -np.flip(out[0] - 1j * out[1])
        W_ref = np.fft.fftshift(g[f"{name}.W"])
        assert np.array_equal(W_gpu != 0, W_ref != 0), name
        sup = W_ref != 0
        ip = np.vdot(W_ref[sup], W_gpu[sup])
        nrm = np.sqrt(np.vdot(W_ref[sup], W_ref[sup]).real * np.vdot(W_gpu[sup], W_gpu[sup]).real)
        assert 1.0 - ip.real / nrm <= 1e-8 and abs(ip.imag) / nrm <= 1e-5, (name, ip / nrm)


def test_fd_window_banded_kernel(torch_cuda):
    """SURVEY 8f rank 2 as a hand-written kernel (FDutils.py:35-47,66-101): the conjugated DFT of a Hann-type window is a narrow
    band, so get_fd_windowed applies it as a banded stencil (emrifd_window_taps: direct DFT of the central taps + Parseval bound;
    emrifd_band_convolve).  (1) small N, forced band: against the reference's verbatim convolve(hstack((a[1:], a)), b, 'valid')/len(b)
    within the certified bound; (2) the 1-yr grid length N = 3 155 815 with the scripts' windows at the default bound: against the
    exact FFT evaluation, the measured error below the certified one; (3) window_in_fd=True; (4) f >= 0 output range."""
    torch = torch_cuda
    from scipy.signal import convolve
    from scipy.signal.windows import hann, blackman, blackmanharris, hamming, nuttall
    from emri_frequencydomainwaveforms_b200 import fdutils
    rng = np.random.default_rng(12)
    # (1) verbatim reference at a size the O(N^2) direct sum finishes
    n = 1025
    window = hann(n)
    sig = [rng.normal(size=n) + 1j * rng.normal(size=n) for _ in range(2)]
    band = fdutils.window_band(window, rtol=5e-3)
    assert band is not None and band[1] <= 16 and band[2] <= 5e-3
    out = fdutils.get_fd_windowed(sig, window, rtol=5e-3)
    cw = np.conj(np.fft.fft(window))
    for c in range(2):
        ref = convolve(np.hstack((cw[1:], cw)), sig[c], mode="valid") / n                 # FDutils.py:47 verbatim
        scale = np.sqrt(np.sum(np.abs(cw) ** 2)) * np.sqrt(np.sum(np.abs(sig[c]) ** 2)) / n
        assert np.max(np.abs(out[c].cpu().numpy() - ref)) <= 1.01 * band[2] * scale
    # the band is symmetric and its centre tap is sum(window)
    taps = band[0].cpu().numpy()
    assert abs(taps[band[1]] - window.sum()) <= 1e-12 * window.sum() and np.allclose(taps, np.conj(taps[::-1]), rtol=0, atol=1e-9)
    # (2) full 1-yr grid length (odd, not a power of two)
    N = 3155815
    s = torch.complex(torch.randn(2, N, dtype=torch.float64, device="cuda"), torch.randn(2, N, dtype=torch.float64, device="cuda"))
    for ww in (hann, blackman, blackmanharris, hamming, nuttall):
        w = torch.as_tensor(ww(N)).cuda()
        band = fdutils.window_band(w)
        assert band is not None and band[1] <= 16, (ww.__name__, band)
        got = fdutils.get_fd_windowed([s[0], s[1]], w)
        cwd = torch.conj(torch.fft.fft(w.to(torch.complex128)))
        for c in range(2):
            exact = fdutils._fft_convolution(cwd, s[c])
            rel = (torch.linalg.vector_norm(got[c] - exact) / torch.linalg.vector_norm(exact)).item()
            assert rel <= max(band[2], 3e-8) * 1.5 and rel <= 2e-7, (ww.__name__, rel, band[1:])
    # (3) the same window handed over in the frequency domain
    w = torch.as_tensor(hann(N)).cuda()
    fw = torch.fft.fft(w.to(torch.complex128))
    g_td = fdutils.get_fd_windowed([s[0], s[1]], w)
    g_fd = fdutils.get_fd_windowed([s[0], s[1]], fw, window_in_fd=True)
    assert (torch.linalg.vector_norm(g_td[0] - g_fd[0]) / torch.linalg.vector_norm(g_td[0])).item() <= 1e-7
    # (4) f >= 0 half only == slice of the full result; circular wrap at both ends of the array
    lo, cnt = (N - 1) // 2, (N + 1) // 2
    half = fdutils.get_fd_windowed([s[0], s[1]], w, out_lo=lo, out_n=cnt)
    assert torch.equal(half[1], g_td[1][lo:]) and half[0].shape[0] == cnt
    exact = fdutils._fft_convolution(torch.conj(fw), s[0])
    for k in (0, 1, N - 1):
        assert abs((g_td[0][k] - exact[k]).item()) <= 1e-6 * abs(exact[k].item()) + 1e-9
    # get_convolution routes a band-limited first argument through the same kernel, anything else through the exact path
    a_bl = torch.conj(fw)
    assert (torch.linalg.vector_norm(fdutils.get_convolution(a_bl, s[0]) - exact) / torch.linalg.vector_norm(exact)).item() <= 2e-7
    a_gen = torch.complex(torch.randn(4097, dtype=torch.float64, device="cuda"), torch.randn(4097, dtype=torch.float64, device="cuda"))
    b_gen = torch.complex(torch.randn(4097, dtype=torch.float64, device="cuda"), torch.randn(4097, dtype=torch.float64, device="cuda"))
    ref = fdutils._fft_convolution(a_gen, b_gen)
    assert torch.allclose(fdutils.get_convolution(a_gen, b_gen), ref, rtol=1e-12, atol=1e-12)

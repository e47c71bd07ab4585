"""GPU tests of the device producers that sit immediately before the path:
Ylm, mode selection by power, compaction, and the batched end-to-end FDTemplateModel.get_ll built on them.
Bar: kept-mode index sets bit-exact against the numpy restatement (oracle.mode_select_ref) on the same inputs;
Ylm within 1e-13 (floating point); likelihoods identical to the host-packed path on the same packed inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def test_ylm_device_matches_wigner_sum(generator, torch_cuda):
    from emri_frequencydomainwaveforms_b200 import _lib
    from emri_frequencydomainwaveforms_b200.utils.ylm import ylm_batch_device
    from oracle.oracle import ylm_ref
    h = _lib.get_handle()
    basis = generator._device_basis(h)
    theta = np.array([0.0, np.pi, np.pi / 3, 1.0, 2.9, 1e-9])
    phi = np.array([0.0, -np.pi / 2, 1.3, 3 * np.pi / 2, 0.2, 6.0])
    out = ylm_batch_device(basis["l"], basis["m"], basis["neg_src"], theta, phi, h).cpu().numpy()
    M = generator.num_teuk_modes
    src = np.concatenate([np.arange(M), np.where(generator.m0mask)[0]])
    sign = np.concatenate([np.ones(M, dtype=int), -np.ones(out.shape[1] - M, dtype=int)])
    for w in range(len(theta)):
        cache = {}
        ref = np.empty(out.shape[1], dtype=np.complex128)
        for i in range(out.shape[1]):
            key = (int(generator.l_arr[src[i]]), int(sign[i] * generator.m_arr[src[i]]))
            if key not in cache:
                cache[key] = ylm_ref(key[0], key[1], theta[w], phi[w])
            ref[i] = cache[key]
        assert np.max(np.abs(out[w] - ref)) <= 1e-13, (w, np.max(np.abs(out[w] - ref)))
    # face-on: only m = 2 survives at theta = 0, only m = -2 at theta = pi (tests/test_host_cpu.py does the host twin)
    m_all = sign * generator.m_arr[src]
    assert np.all(out[0][m_all != 2] == 0) and np.all(np.abs(out[0][m_all == 2]) > 0)


def test_device_amplitude_matches_host_standin(generator, torch_cuda):
    """emrifd_synth_amplitude (device stand-in for the ROMAN amplitudes) against the NumPy version of the same formula:
    floating point, 1e-13 of each mode's own magnitude (pow / exp / sincos implementations differ by rounding)."""
    amp = generator.amplitude_generator
    rng = np.random.default_rng(3)
    p = rng.uniform(7.3, 17.0, 57)
    e = rng.uniform(0.0, 0.74, 57)
    e[:3] = [0.0, 0.74, 0.35]
    host = amp(p, e)
    dev = amp.device_call(p, e, torch_cuda.device("cuda", torch_cuda.cuda.current_device())).cpu().numpy()
    assert dev.shape == host.shape == (57, amp.num_teuk_modes)
    big = np.abs(host) > 1e-290
    assert np.max(np.abs(dev - host)[big] / np.abs(host)[big]) <= 1e-12
    assert np.max(np.abs(dev - host)[~big], initial=0.0) <= 1e-290


@pytest.mark.parametrize("eps", [1e-2, 1e-5, 0.3])
def test_mode_select_device_bit_exact(generator, torch_cuda, eps):
    from emri_frequencydomainwaveforms_b200 import _lib
    from oracle.oracle import mode_select_ref
    torch = torch_cuda
    h = _lib.get_handle()
    rng = np.random.default_rng(11)
    walkers, teuks, ylms, sw = [(1e6, 10.0, 12.0, 0.35, np.pi / 3), (5e5, 20.0, 10.5, 0.6, 0.7), (1e6, 50.0, 9.0, 0.3, np.pi)], [], [], []
    for w, (M, mu, p0, e0, th) in enumerate(walkers):
        t, p, e, *_ = generator.inspiral_generator(M, mu, 0.0, p0, e0, 1.0, T=0.1, dt=10.0)
        sub = np.sort(rng.choice(len(p), size=min(len(p), 9), replace=False))
        teuks.append(generator.amplitude_generator(p[sub], e[sub]))
        nl = len(generator.unique_l)
        y = generator.ylm_gen(generator.unique_l, generator.unique_m, th, -np.pi / 2)
        ylms.append(np.concatenate([y[:nl][generator.inverse_lm], y[nl:][generator.inverse_lm][generator.m0mask]]))
        sw += [w] * len(sub)
    teuk_dev = torch.from_numpy(np.concatenate(teuks)).cuda()
    flags = generator.mode_selector.select_device(teuk_dev, np.asarray(sw, dtype=np.int32), np.stack(ylms), eps=eps, handle=h)
    flags = flags.cpu().numpy()
    for w in range(len(walkers)):
        ref = mode_select_ref(teuks[w], ylms[w], generator.m0mask, eps)
        assert np.array_equal(np.where(flags[w])[0], ref), (w, flags[w].sum(), len(ref))


def _params(nb, rng):
    P = np.zeros((nb, 14))
    P[:, 0] = np.exp(rng.uniform(np.log(5e5), np.log(3e6), nb))
    P[:, 1] = P[:, 0] * np.exp(rng.uniform(np.log(1e-5), np.log(5e-5), nb))
    P[:, 4] = rng.uniform(0.05, 0.6, nb)
    P[:, 3] = rng.uniform(9.5, 12.0, nb) + 2 * P[:, 4] * 0.5
    P[:, 5], P[:, 6] = 1.0, rng.uniform(0.5, 2.0, nb)
    P[:, 7:11] = rng.uniform(0.2, 2.8, (nb, 4))
    P[:, 11], P[:, 13] = rng.uniform(0, 2 * np.pi, nb), rng.uniform(0, 2 * np.pi, nb)
    return P


def test_device_producers_end_to_end(torch_cuda):
    """prepare_batch_device -> fused likelihood, against (a) the numpy selection restatement on the same device inputs,
    (b) the host-packed C-ABI path on the downloaded packed arrays (must be identical), (c) the host-producer path
    (amplitudes differ by rounding only)."""
    torch = torch_cuda
    from emri_frequencydomainwaveforms_b200 import _lib, engine
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    from oracle.oracle import mode_select_ref
    T, dt, eps = 0.05, 10.0, 1e-3
    gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True),
                               return_list=True)
    base = gen.waveform_generator
    rng = np.random.default_rng(5)
    P = _params(6, rng)
    P[2, 4] = 0.9          # e0 > 0.75: out of domain -> NaN, like the reference's skipped draws
    inj = gen(*P[0], T=T, dt=dt, eps=eps, mask_positive=True)
    n = inj[0].shape[0]
    N = 2 * n - 1
    wfac = np.full((2, n), 2.0e19)
    data = np.stack([c.cpu().numpy() for c in inj]) * wfac
    tm_dev = FDTemplateModel(gen, producers="device")
    tm_host = FDTemplateModel(gen, producers="host")
    ll_dev = tm_dev.get_ll(P, data, wfac, T=T, dt=dt, eps=eps, N=N)
    ll_host = tm_host.get_ll(P, data, wfac, T=T, dt=dt, eps=eps, N=N)
    assert np.isnan(ll_dev[2]) and np.isnan(ll_host[2])
    ok = ~np.isnan(ll_host)
    dd = 4.0 * np.sum(np.abs(data) ** 2)
    assert np.all(np.abs(ll_dev[ok] - ll_host[ok]) <= 1e-9 * dd), (ll_dev, ll_host)
    assert abs(ll_dev[0]) <= 1e-9 * dd          # the injection itself
    # (a) + (b): same inputs, index-exact selection and identical likelihood through the host-packed entry point
    h = _lib.get_handle()
    ang = np.array([gen._transform(*row[7:11]) for row in P])
    db, okd = base.prepare_batch_device(P[:, 0], P[:, 1], P[:, 3], P[:, 4], ang[:, 0], ang[:, 1], dist=P[:, 6], Phi_phi0=P[:, 11],
                                        Phi_r0=P[:, 13], T=T, dt=dt, eps=eps, cos2psi=ang[:, 2], sin2psi=ang[:, 3], handle=h,
                                        keep_full=True)
    assert np.array_equal(okd, ok)
    W = db.pb.walkers
    teuk_full, ylm_full = db.teuk_full.cpu().numpy(), db.ylm_full.cpu().numpy()
    teuk, ylm = db.teuk.cpu().numpy(), db.ylm.cpu().numpy()
    m_k, n_k = db.m.cpu().numpy(), db.n.cpu().numpy()
    tr = {k: getattr(db, k).cpu().numpy() for k in ("t", "f_phi", "f_r", "Phi_phi", "Phi_r")}
    items = []
    for w in range(len(W)):
        L, K, ko, to, mo = (int(W[w][k]) for k in ("L", "K", "knot_off", "teuk_off", "mode_off"))
        ref = mode_select_ref(teuk_full[ko:ko + L], ylm_full[w], base.m0mask, eps)
        assert K == len(ref)
        assert np.array_equal(m_k[mo:mo + K], base.m_arr[ref]) and np.array_equal(n_k[mo:mo + K], base.n_arr[ref])
        blk = teuk[to:to + L * K].reshape(L, K)
        assert np.array_equal(blk, teuk_full[ko:ko + L][:, ref])
        pos = np.cumsum(base.m0mask) - 1
        neg = np.where(base.m0mask[ref], base.num_teuk_modes + pos[ref], ref)
        assert np.array_equal(ylm[2 * mo:2 * mo + 2 * K], np.concatenate([ylm_full[w][ref], ylm_full[w][neg]]))
        items.append(dict(t=tr["t"][ko:ko + L], f_phi=tr["f_phi"][ko:ko + L], f_r=tr["f_r"][ko:ko + L], Phi_phi=tr["Phi_phi"][ko:ko + L],
                          Phi_r=tr["Phi_r"][ko:ko + L], teuk_modes=blk, m_arr=m_k[mo:mo + K], n_arr=n_k[mo:mo + K],
                          ylms=ylm[2 * mo:2 * mo + 2 * K], scale=W[w]["scale"], cos2psi=W[w]["cos2psi"], sin2psi=W[w]["sin2psi"]))
    out_host = engine.run_loglike_host(engine.PackedBatch(items), h, N, 1.0 / (N * dt))
    out_dev = engine.run_loglike(db, N, 1.0 / (N * dt)).cpu().numpy()
    assert np.array_equal(out_host, out_dev)

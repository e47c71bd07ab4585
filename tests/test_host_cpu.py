"""CPU tests of the host logic and of the C-ABI library's export table (no compute without a GPU)."""
import ctypes
import os
import sys
import re

import numpy as np
import pytest

from helpers import CASES, grid_size, make_item

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    from emri_frequencydomainwaveforms_b200 import _lib
    from emri_frequencydomainwaveforms_b200.csrc import build
    build.build()
    header = open(os.path.join(ROOT, "include", "emrifd.h")).read()
    declared = set(re.findall(r"\b(emrifd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.emrifd_version() == 200
    assert lib.emrifd_sizeof_branch() == _lib.BRANCH_DTYPE.itemsize == 72
    assert lib.emrifd_sizeof_walker() == _lib.WALKER_DTYPE.itemsize


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from emri_frequencydomainwaveforms_b200 import _lib
    from emri_frequencydomainwaveforms_b200.summation.interpolatedmodesum import CubicSplineInterpolant
    with pytest.raises(_lib.EmrifdError):
        CubicSplineInterpolant(np.arange(5.0), np.arange(5.0))
    # the C-ABI itself refuses to create a handle without a device (no abort, an error code)
    lib = _lib.load()
    hp = ctypes.c_void_p()
    assert lib.emrifd_create(0, None, ctypes.byref(hp)) == -6 and not hp.value


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "emri_frequencydomainwaveforms_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|liboracle|oracle[/.]_build|#include.*oracle|orc_[a-z_]+\()", src), f


def test_fundamental_frequencies_against_quadrature():
    from emri_frequencydomainwaveforms_b200.utils.utility import get_fundamental_frequencies, get_separatrix
    p = np.array([12.0, 8.0, 7.5, 10.0, 6.3, 20.0, 9.1])
    e = np.array([0.35, 0.5, 0.7, 1e-6, 0.1, 0.0, 0.62])
    om_phi, om_th, om_r = get_fundamental_frequencies(0.0, p, e, np.ones_like(p))
    chi = 2 * np.pi * np.arange(8192) / 8192
    for i in range(len(p)):
        c = np.cos(chi)
        dt = p[i] ** 2 / ((p[i] - 2 - 2 * e[i] * c) * (1 + e[i] * c) ** 2) * np.sqrt(((p[i] - 2) ** 2 - 4 * e[i] ** 2) / (p[i] - 6 - 2 * e[i] * c))
        dphi = np.sqrt(p[i] / (p[i] - 6 - 2 * e[i] * c))
        Tr = dt.mean() * 2 * np.pi
        assert abs(om_r[i] * Tr / (2 * np.pi) - 1) < 5e-14
        assert abs(om_phi[i] * Tr / (dphi.mean() * 2 * np.pi) - 1) < 5e-14
    assert np.allclose(om_phi[5], 20.0 ** -1.5) and np.allclose(om_r[5], 20.0 ** -1.5 * np.sqrt(1 - 6 / 20.0))
    assert np.array_equal(get_separatrix(0.0, e, 1.0), 6 + 2 * e)
    with pytest.raises(ValueError):
        get_fundamental_frequencies(0.5, p, e, np.ones_like(p))


def test_ylm_closed_forms_and_poles():
    from emri_frequencydomainwaveforms_b200.utils.ylm import GetYlms, spin_weighted_ylm
    th, ph = 0.7, 0.3
    assert np.isclose(spin_weighted_ylm(-2, 2, 2, th, ph), np.sqrt(5 / (64 * np.pi)) * (1 + np.cos(th)) ** 2 * np.exp(2j * ph))
    assert np.isclose(spin_weighted_ylm(-2, 2, -2, th, ph), np.sqrt(5 / (64 * np.pi)) * (1 - np.cos(th)) ** 2 * np.exp(-2j * ph))
    assert np.isclose(spin_weighted_ylm(-2, 2, 0, th, ph), np.sqrt(15 / (32 * np.pi)) * np.sin(th) ** 2)
    assert np.isclose(spin_weighted_ylm(-2, 3, 3, th, ph), -np.sqrt(21 / (2 * np.pi)) * np.cos(th / 2) ** 5 * np.sin(th / 2) * np.exp(3j * ph))
    g = GetYlms(assume_positive_m=True)
    y = g(np.array([2, 2, 3]), np.array([2, 0, 2]), np.pi, -np.pi / 2)      # the scripts' geometry: theta = pi
    assert len(y) == 6 and abs(y[0]) < 1e-30 and abs(y[3]) > 0.1            # only Y_{l,-2} survives at the south pole
    # orthonormality of l = 2..4, m = 1 on a quadrature grid
    x, w = np.polynomial.legendre.leggauss(40)
    for l1 in (2, 3, 4):
        for l2 in (2, 3, 4):
            v = sum(wi * spin_weighted_ylm(-2, l1, 1, np.arccos(xi), 0.0) * np.conj(spin_weighted_ylm(-2, l2, 1, np.arccos(xi), 0.0))
                    for xi, wi in zip(x, w)) * 2 * np.pi
            assert np.isclose(v, 1.0 if l1 == l2 else 0.0, atol=1e-12)
    with pytest.raises(ValueError):
        g(np.array([2]), np.array([-2]), 1.0, 0.0)


def test_mode_selector_semantics(generator):
    it2 = make_item(generator, "plunge")
    it5 = make_item(generator, "plunge", eps=1e-5)
    s2 = set(zip(it2["l_arr"].tolist(), it2["m_arr"].tolist(), it2["n_arr"].tolist()))
    s5 = set(zip(it5["l_arr"].tolist(), it5["m_arr"].tolist(), it5["n_arr"].tolist()))
    assert s2 < s5 and len(s5) > 3 * len(s2)                    # smaller eps keeps a superset
    assert np.all(it2["m_arr"] >= 0) and len(it2["ylms"]) == 2 * len(it2["m_arr"])
    K = len(it2["m_arr"])
    m0 = it2["m_arr"] == 0
    assert np.array_equal(it2["ylms"][:K][m0], it2["ylms"][K:][m0])   # m = 0: the -m slot repeats the +m ylm
    it3 = generator.prepare(1e6, 10.0, 12.0, 0.35, 1.0, -np.pi / 2, T=0.05, mode_selection=[(2, 2, 0), (2, -2, 1), (3, 0, 1)])
    assert sorted(zip(it3["l_arr"], it3["m_arr"], it3["n_arr"])) == [(2, 2, -1), (2, 2, 0), (3, 0, 1)]
    with pytest.raises(ValueError):
        generator.prepare(1e6, 10.0, 12.0, 0.35, 1.0, 0.0, T=0.05, mode_selection=[])
    with pytest.raises(ValueError):
        generator.prepare(1e6, 10.0, 12.0, 0.9, 1.0, 0.0, T=0.05)         # sanity_check_init: e0 > 0.75


def test_trajectory_and_sizing(generator):
    from emri_frequencydomainwaveforms_b200.trajectory.inspiral import EMRIInspiral
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200.utils.constants import YRSID_SI
    traj = EMRIInspiral(func="SchwarzEccFlux")
    t, p, e, x, Pp, Pt, Pr = traj(1e6, 10.0, 0.0, 12.0, 0.35, 1.0, T=1.0)
    assert 20 <= len(t) <= 200 and t[0] == 0.0 and np.all(np.diff(t) > 0)
    assert np.isclose(t[-1], YRSID_SI) and np.all(np.diff(p) < 0) and np.all(np.diff(Pp) > 0)
    assert grid_size(t, 1.0, 10.0) == 3155815                     # N at T = 1 yr, dt = 10 s
    assert grid_size(t, 2.0, 10.0) % 2 == 1
    p0 = get_p_at_t(traj, 0.2, [1e6, 50.0, 0.0, 0.3, 1.0])
    t2, p2, e2, *_ = traj(1e6, 50.0, 0.0, p0, 0.3, 1.0, T=1.0)
    assert abs(t2[-1] / YRSID_SI - 0.2) < 1e-6 and abs(p2[-1] - (6 + 2 * e2[-1] + 0.1)) < 1e-8
    with pytest.raises(ValueError):
        traj(1e6, 10.0, 0.0, 6.5, 0.3, 1.0)


def test_grid_validation_and_packing(generator):
    from emri_frequencydomainwaveforms_b200 import engine
    N, fpos = engine.grid_from_frequency(np.fft.fftshift(np.fft.fftfreq(101, 10.0)))
    assert N == 101 and len(fpos) == 51 and fpos[0] == 0.0
    p = np.linspace(0.0, 0.01, 50)
    N, fpos = engine.grid_from_frequency(np.hstack((-p[::-1][:-1], p)))       # emri_pe.py:339-342
    assert N == 99 and np.array_equal(fpos, p)
    for bad in (np.array([]), np.arange(-2.0, 2.0), np.array([-1.0, 0.0, 2.0]), np.array([-1.0, 0.5, 1.0])):
        with pytest.raises(ValueError):
            engine.grid_from_frequency(bad)
    items = [make_item(generator, "cfg1_like"), make_item(generator, "ecc_many")]
    pb = engine.PackedBatch(items)
    w = pb.walkers
    assert pb.B == 2 and w["knot_off"][1] == w["L"][0] and w["mode_off"][1] == w["K"][0]
    assert w["coeff_off"][1] == w["L"][0] * (2 * w["K"][0] + 4) * 4 and len(pb.ylm) == 2 * pb.n_modes
    assert pb.h2d_bytes() > 0
    bad = dict(items[0], m_arr=items[0]["m_arr"][:-1])
    with pytest.raises(ValueError):
        engine.PackedBatch([bad])


def test_transform_container_contract():
    """The 6 -> 14 parameter fill + (logM, log eta) -> (M, mu) of emri_pe.py:161-206 (eryn transform.py:181-226)."""
    g = np.load(os.path.join(GOLD, "lisatools_golden.npz"))
    p6, p14 = g["tc_in"], g["tc_out"]
    from emri_frequencydomainwaveforms_b200.utils.transform import fill_and_transform
    out = fill_and_transform(p6, fill_inds=np.array([2, 5, 6, 7, 8, 9, 10, 12]),
                             fill_values=np.array([0.0, 1.0, 2.45, np.pi / 3, np.pi / 3, np.pi / 3, np.pi / 3, 0.0]))
    assert np.allclose(out, p14, rtol=1e-15)


def test_native_host_producers_match_python_twins():
    """csrc/emrihost.c (native trajectory ODE + Schwarzschild frequencies) against the SciPy implementations."""
    from scipy.interpolate import CubicSpline
    from emri_frequencydomainwaveforms_b200 import _hostlib
    from emri_frequencydomainwaveforms_b200.csrc import build
    from emri_frequencydomainwaveforms_b200.trajectory.inspiral import EMRIInspiral
    from emri_frequencydomainwaveforms_b200.utils.utility import schwarzschild_frequencies
    build.build_host()
    assert _hostlib.load() is not None
    rng = np.random.default_rng(0)
    e = rng.uniform(0, 0.75, 500)
    p = 6 + 2 * e + 0.1 + rng.uniform(0, 12, 500)
    a, b = schwarzschild_frequencies(p, e, native=True)
    ra, rb = schwarzschild_frequencies(p, e, native=False)
    assert np.max(np.abs(a / ra - 1)) < 5e-15 and np.max(np.abs(b / rb - 1)) < 5e-15
    nat, py = EMRIInspiral(use_native=True), EMRIInspiral(use_native=False)
    for (M, mu, p0, e0, T) in [(1e6, 10.0, 12.0, 0.35, 1.0), (1e6, 50.0, 9.0, 0.3, 0.25), (1e5, 1.0, 15.0, 0.6, 0.5)]:
        tn, pn, en, _, Ppn, _, Prn = nat(M, mu, 0.0, p0, e0, 1.0, Phi_phi0=0.3, Phi_r0=1.1, T=T)
        tp, pp, ep, _, Ppp, _, Prp = py(M, mu, 0.0, p0, e0, 1.0, Phi_phi0=0.3, Phi_r0=1.1, T=T)
        assert abs(len(tn) - len(tp)) <= 2 and tn[0] == 0.0 and np.all(np.diff(tn) > 0)
        assert abs(tn[-1] - tp[-1]) <= 1e-9 * tp[-1]                       # same end (T or the separatrix buffer)
        assert abs(Ppn[-1] - Ppp[-1]) <= 1e-9 * abs(Ppp[-1]) and abs(Prn[-1] - Prp[-1]) <= 1e-9 * abs(Prp[-1])
        tt = np.linspace(0, min(tn[-1], tp[-1]), 500)
        assert np.max(np.abs(CubicSpline(tn, pn)(tt) - CubicSpline(tp, pp)(tt))) < 1e-7      # different knots, same orbit
    with pytest.raises(ValueError):
        nat(1e6, 10.0, 0.0, 6.5, 0.3, 1.0)


def test_batched_host_helpers_match_scalar_twins(generator):
    """The vectorised domain checks and SSB transform used by the batched device-producer path make the same decisions /
    return the same doubles as the per-walker versions the single-waveform call uses."""
    import warnings
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform, ssb_transform_batch
    g = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
    rng = np.random.default_rng(0)
    A = rng.uniform(0.0, np.pi, (500, 4))
    A[:3] = np.pi / 3                                   # the scripts' geometry: theta = pi, sin(phiS - phiK) = 0
    A[3] = [1.0, 2.0, np.pi - 1.0, 2.0 + np.pi]         # R = -S: theta = 0
    ref = np.array([g._transform(*row) for row in A])
    got = np.stack(ssb_transform_batch(A[:, 0], A[:, 1], A[:, 2], A[:, 3]), axis=1)
    assert np.array_equal(ref, got)
    src = np.stack(ssb_transform_batch(A[:, 0], A[:, 1], A[:, 2], A[:, 3], detector_frame=False), axis=1)
    assert np.array_equal(src[:, 0], ref[:, 0]) and np.all(src[:, 2] == 1.0) and np.all(src[:, 3] == 0.0)
    # trajectories: the threaded batch equals the single calls bit for bit
    from emri_frequencydomainwaveforms_b200 import _hostlib
    if _hostlib.load() is not None:
        M, mu = np.array([1e6, 5e5, 2e6]), np.array([10.0, 20.0, 50.0])
        p0, e0 = np.array([12.0, 10.5, 9.0]), np.array([0.35, 0.6, 0.3])
        ig = generator.inspiral_generator
        out1, l1 = _hostlib.trajectory_batch(M, mu, p0, e0, np.zeros(3), np.ones(3), 0.1, ig.rtol, ig.atol, ig.max_init_len, nthreads=1)
        out4, l4 = _hostlib.trajectory_batch(M, mu, p0, e0, np.zeros(3), np.ones(3), 0.1, ig.rtol, ig.atol, ig.max_init_len, nthreads=4)
        assert np.array_equal(l1, l4)
        for a, b in zip(out1, out4):
            for w in range(3):
                assert np.array_equal(a[w, :l1[w]], b[w, :l4[w]])
        t, p, e, x, Pp, Pt, Pr = ig(M[1], mu[1], 0.0, p0[1], e0[1], 1.0, Phi_phi0=0.0, Phi_r0=1.0, T=0.1, dt=10.0)
        assert np.array_equal(t, out1[0][1, :l1[1]]) and np.array_equal(Pr, out1[4][1, :l1[1]])


def test_bench_reference_arm_prints_one_json_line():
    """Driver contract: `bench.py --impl reference` runs on the host cores alone and prints exactly ONE JSON line on stdout
    with the keys the driver reads (library chatter is diverted to stderr by bench.StdoutGuard)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fd_waveform_likelihoods_per_s" and d["unit"] == "walkers/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "walkers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_likelihood_adopts_the_plugin():
    """The reference's OWN Likelihood (LISAanalysistools/lisatools/sampling/likelihood.py) hands the whole batch to a template
    model that has ``get_ll`` (:70-72) and, with fill_data_noise=True, appends (injection_channels, noise_factor) (:330-331).
    A stub with FDTemplateModel's ``get_ll`` signature records what it is called with: parameters already filled / transformed
    by eryn's TransformContainer, subset chunking, the whitened data and noise factor of inject_signal, the waveform kwargs."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("/root/reference is not on this machine")
    import inspect
    for sub in ("LISAanalysistools", "Eryn"):
        if os.path.join(ref, sub) not in sys.path:
            sys.path.insert(0, os.path.join(ref, sub))
    from lisatools.sampling.likelihood import Likelihood as RefLikelihood
    from eryn.utils import TransformContainer
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel

    calls = []

    class Stub:
        def __call__(self, *params, **kw):          # used once by inject_signal(params=...): [h+, hx] on f >= 0
            n = 64
            return [np.full(n, params[0] * 1e-30 + 0j), np.full(n, params[1] * 1e-30 + 0j)]

        def get_ll(self, params, data=None, noise_factor=None, T=1.0, dt=10.0, eps=1e-5, **kwargs):
            calls.append(dict(params=np.array(params), data=data, noise_factor=noise_factor, T=T, dt=dt, eps=eps, extra=kwargs))
            return -np.arange(len(params), dtype=float)

    # the plug-in's signature is what the reference will call: same leading parameters as the stub
    sig = list(inspect.signature(FDTemplateModel.get_ll).parameters)
    assert sig[:4] == ["self", "params", "data", "noise_factor"]
    fill = {"ndim_full": 14, "fill_inds": np.array([2, 5, 6, 7, 8, 9, 10, 12]), "fill_values": np.array([0.0, 1.0, 1.0, 0.3, 0.4, 0.5, 0.6, 0.0])}
    tc = TransformContainer(parameter_transforms={(0, 1): lambda lnM, lneta: (np.exp(lnM), np.exp(lnM) * np.exp(lneta))}, fill_dict=fill)   # emri_pe.py:161-206
    f_arr = np.linspace(0.0, 1e-2, 64)
    like = RefLikelihood(Stub(), 2, f_arr=f_arr, parameter_transforms={"emri": tc}, fill_data_noise=True, vectorized=False,
                         transpose_params=False, subset=3, use_gpu=False)
    assert like.like_here is False                                                  # likelihood.py:70-72: the plug-in took over
    inj6 = np.array([np.log(1e6), np.log(1e-5), 12.0, 0.35, 1.0, 2.0])
    like.inject_signal(params=inj6.copy(), waveform_kwargs=dict(T=1.0), noise_fn=lambda f, **kw: np.full(len(f), 4.0), noise_kwargs={}, add_noise=False)
    params = np.tile(inj6, (7, 1)) + 1e-3 * np.arange(7)[:, None]
    out = like(params, T=0.5, dt=15.0, eps=1e-2)
    assert [len(c["params"]) for c in calls] == [3, 3, 1]                           # subset chunking (likelihood.py:313-319)
    assert np.array_equal(out, np.concatenate([-np.arange(3.0), -np.arange(3.0), -np.arange(1.0)]))
    c = calls[0]
    assert c["params"].shape == (3, 14) and np.allclose(c["params"][:, 0], np.exp(params[:3, 0]))      # (ln M, ln eta) -> (M, mu), 8 filled
    assert np.allclose(c["params"][:, 1], np.exp(params[:3, 0] + params[:3, 1])) and np.allclose(c["params"][:, 6], 1.0)
    assert c["data"] is like.injection_channels and c["noise_factor"] is like.noise_factor            # likelihood.py:330-331
    assert (c["T"], c["dt"], c["eps"]) == (0.5, 15.0, 1e-2)
    df = f_arr[1] - f_arr[0]
    assert np.allclose(np.asarray(like.noise_factor), np.sqrt(df / 4.0))                                # likelihood.py:218-220


def test_window_band_host_logic():
    """fdutils._choose_band (host side of the banded FD-window kernel): the chosen half-width meets the requested bound, the
    Cauchy-Schwarz error bound holds for the truncated circular convolution (numpy restatement of FDutils.py:35-47 as the DFT
    identity), and bounds below the certifiable floor are refused (-> exact FFT evaluation)."""
    from scipy.signal.windows import hann
    from emri_frequencydomainwaveforms_b200.fdutils import _choose_band
    n = 4001
    rng = np.random.default_rng(3)
    w = hann(n)
    W = np.fft.fft(w)
    p2 = np.abs(W) ** 2
    Hm = 128
    e_band = p2[0] + 2.0 * np.concatenate([[0.0], np.cumsum(p2[1:Hm + 1])])
    e_tot = n * np.sum(w * w)                                   # Parseval, as the device path computes it
    assert abs(e_tot - p2.sum()) <= 1e-12 * e_tot
    H, bound = _choose_band(e_band, e_tot, 1e-3)
    assert 1 <= H <= 8 and bound <= 1e-3
    assert _choose_band(e_band, e_tot, 1e-9) is None and _choose_band(e_band, e_tot, 0.0) is None
    a = np.conj(W)
    b = rng.normal(size=n) + 1j * rng.normal(size=n)
    exact = np.fft.ifft(np.fft.fft(a) * np.fft.fft(b)) / n      # == convolve(hstack((a[1:], a)), b, 'valid') / n
    at = np.zeros_like(a)
    idx = np.arange(-H, H + 1) % n
    at[idx] = a[idx]
    trunc = np.fft.ifft(np.fft.fft(at) * np.fft.fft(b)) / n
    assert np.max(np.abs(trunc - exact)) <= bound * np.sqrt(e_tot) * np.linalg.norm(b) / n

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

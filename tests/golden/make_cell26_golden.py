"""Golden vectors from the reference's own statement of the per-harmonic FD construction.

`cell26_fd_waveform` below is the body of ``FD_waveform`` of the reference notebook
(/root/reference/Tutorial_FD_construction_single_mode.ipynb:548-623, cell 26) kept statement for statement: SciPy
``CubicSpline`` for t(f), fdot, fddot, ``scipy.special.kv`` for K_{1/3}, the +f term with A Y_lm, the mirrored term at -f with
conj(A) Y_{l,-m}, the distance scaling of its last line.  Only the three FEW producers the cell calls -- ``traj(...)``,
``amp(p, e, specific_modes=...)`` and ``ylm_gen(...)`` -- are replaced by the synthetic trajectory / amplitude / Ylm arrays
of this repo's stand-in producers (FEW's flux and amplitude data files are not available offline), and
``few.summation.interpolatedmodesum.CubicSplineInterpolant`` by SciPy's not-a-knot ``CubicSpline`` (the notebook itself uses the
two interchangeably, cells 8/11/20).

The output is the raw accumulation W(f) of the cell on an fftfreq grid.  The oracle / CUDA path produce h+, hx after the
flip S(f) = -W(-f) and the Hermitian split; tests/test_oracle_cpu.py::test_cell26_* undo that (W(f) = -[h+ - i hx](-f)) and
compare: the sign, flip, conjugation and +-m conventions must agree exactly; the values agree to the accuracy of the cell's
own approximations (t(f) from a spline of the inverse function, fddot from a spline through the fdot knots).

Run in the build container:  python tests/golden/make_cell26_golden.py   -> tests/golden/cell26_golden.npz
"""
import os
import sys

import numpy as np
from scipy import special
from scipy.interpolate import CubicSpline

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from emri_frequencydomainwaveforms_b200.utils.constants import MRSUN_SI, Gpc  # noqa: E402


def cell26_fd_waveform(freq, t, Phi_phi, Phi_r, f_phi, f_r, true_teuk, ylms, m_sel, n_sel, dist, mu):
    # Phi_mn (t)
    phase_evolution = m_sel*Phi_phi + n_sel*Phi_r
    phase_spline = CubicSpline(t, phase_evolution)

    # t(f)   (the cell: theo_f = (m_sel*OmegaPhi + n_sel*OmegaR)/(2*np.pi*M*MTSUN_SI); f_phi, f_r are those tracks in Hz)
    theo_f = m_sel*f_phi + n_sel*f_r

    if theo_f[0] < theo_f[-1]:
        time_f_spline_0 = CubicSpline(theo_f, t)
    else:   # CubicSpline wants increasing abscissae; same interpolant of the same points
        time_f_spline_0 = CubicSpline(theo_f[::-1], t[::-1])

    # frequency
    index_positive_f = (freq>np.min(theo_f))*(freq<np.max(theo_f))
    index_negative_f = (freq>np.min(-theo_f))*(freq<np.max(-theo_f))
    f_0 = freq[index_positive_f]
    f_1 = freq[index_negative_f]

    # time evluated quantities
    t_f_0 = time_f_spline_0(f_0)
    t_f_1 = np.flip(t_f_0)

    # Fdot
    fdot_spline_0 = CubicSpline(t,theo_f).derivative()
    fdot_spline_1 = CubicSpline(t,-theo_f).derivative()

    # Fddot
    fdd_0 = CubicSpline(t,fdot_spline_0(t)).derivative()
    fdd_1 = CubicSpline(t,fdot_spline_1(t)).derivative()

    # amplitude
    interp_teuk = np.zeros((2,true_teuk.shape[0]))

    interp_teuk[0,:] = true_teuk.T.real
    interp_teuk[1,:] = true_teuk.T.imag

    H_spline = CubicSpline(t, interp_teuk, axis=1)

    arg_0 = -2*np.pi*1j* fdot_spline_0(t_f_0)**3 / (3*fdd_0(t_f_0)**2)
    K_1over3_0 = special.kv(1./3.,arg_0)*np.exp(arg_0)

    arg_1 = -2*np.pi*1j* fdot_spline_1(t_f_1)**3 / (3*fdd_1(t_f_1)**2)
    K_1over3_1 = special.kv(1./3.,arg_1)*np.exp(arg_1)

    Amp0 = (H_spline(t_f_0)[0] + 1j* H_spline(t_f_0)[1])*ylms[0] \
            *1j* fdot_spline_0(t_f_0)/np.abs(fdd_0(t_f_0)) \
            * K_1over3_0 * 2/np.sqrt(3)

    Amp1 = (H_spline(t_f_1)[0] - 1j* H_spline(t_f_1)[1])*ylms[1]\
            *1j* fdot_spline_1(t_f_1)/np.abs(fdd_1(t_f_1)) \
            * K_1over3_1 * 2/np.sqrt(3)

    Exp0 = np.exp(1j*(2*np.pi*f_0* t_f_0  - phase_spline(t_f_0)) )
    Exp1 = np.exp(1j*(2*np.pi*f_1* t_f_1  + phase_spline(t_f_1)) )

    # final waveform
    h = np.zeros_like(freq,dtype=complex)
    h[index_positive_f] = Amp0*Exp0
    h[index_negative_f] = Amp1*Exp1

    return h / ((dist * Gpc) / (mu * MRSUN_SI))


# (name, M, mu, p0, e0, theta, phi, T [yr], dt, (l, m, n)) -- single monotone harmonics
CASES = [
    ("l2m2n0_rising", 1e6, 50.0, 10.0, 0.4, np.pi / 4, np.pi / 3, 0.1, 20.0, (2, 2, 0)),       # the cell's own parameters (shorter T)
    ("l2m1nm4_negative_f", 1e6, 50.0, 10.0, 0.4, 1.0, -np.pi / 2, 0.25, 20.0, (2, 1, -4)),     # f_mn < 0 and falling: direct term lands at f < 0
    ("l3m2n1_rising", 5e5, 20.0, 11.0, 0.5, 2.2, 0.7, 0.05, 20.0, (3, 2, 1)),
]


def main():
    from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
    from emri_frequencydomainwaveforms_b200.utils.constants import YRSID_SI
    gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True), producers="host")
    out = {"names": np.array([c[0] for c in CASES])}
    for name, M, mu, p0, e0, theta, phi, T, dt, (l, m, n) in CASES:
        it = gen.prepare(M, mu, p0, e0, theta, phi, dist=1.0, Phi_phi0=0.4, Phi_r0=2.0, T=T, dt=dt, mode_selection=[(l, m, n)])
        theo_f = m * it["f_phi"] + n * it["f_r"]
        assert np.all(np.diff(theo_f) > 0) or np.all(np.diff(theo_f) < 0), f"{name}: harmonic is not monotone"
        N = int(T * YRSID_SI / dt) + 1
        N += (N % 2 == 0)
        freq = np.fft.fftfreq(N, dt)
        ylms = it["ylms"]     # [Y_lm, Y_l-m] of the single selected mode
        h = cell26_fd_waveform(freq, it["t"], it["Phi_phi"], it["Phi_r"], it["f_phi"], it["f_r"], it["teuk_modes"][:, 0], ylms, m, n, 1.0, mu)
        assert np.all(np.isfinite(h.view(float))) and np.count_nonzero(h) > 100, name
        for k in ("t", "p", "e", "Phi_phi", "Phi_r", "f_phi", "f_r", "teuk_modes", "ylms"):
            out[f"{name}.{k}"] = it[k]
        out[f"{name}.lmn"] = np.array([l, m, n])
        out[f"{name}.params"] = np.array([M, mu, T, dt, N, it["scale"]])
        out[f"{name}.W"] = h
        print(name, "N", N, "support", np.count_nonzero(h), "max|W|", np.abs(h).max(), "f range", theo_f.min(), theo_f.max())
    np.savez_compressed(os.path.join(HERE, "cell26_golden.npz"), **out)


if __name__ == "__main__":
    main()

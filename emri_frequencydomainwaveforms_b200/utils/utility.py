"""Host-side helpers that sit on the FD path.

``get_fundamental_frequencies`` mirrors ``few.utils.utility.get_fundamental_frequencies`` as
it is called inside ``FDInterpolatedModeSum.sum`` and by the reference notebook
(Tutorial_FD_construction_single_mode.ipynb:280, cell 11): Schwarzschild (a = 0) only,
dimensionless Omega_phi, Omega_theta, Omega_r.
"""
import numpy as np
from scipy.special import ellipk, ellipe, elliprf, elliprj
from scipy.optimize import brentq

from .. import _hostlib
from .constants import MTSUN_SI, YRSID_SI


def _ellip_pi(n, m):
    """Complete elliptic integral of the third kind Pi(n | m) through Carlson's forms."""
    return elliprf(0.0, 1.0 - m, 1.0) + n / 3.0 * elliprj(0.0, 1.0 - m, 1.0, 1.0 - n)


def fundamental_frequencies_hz(p, e, M, native=True):
    """f_phi, f_r [Hz] at the sparse trajectory points: Omega / (2 pi M MTSUN_SI), evaluated as Omega / ((2 pi) * (M MTSUN_SI)) --
    the association csrc/emrihost.c uses, so every producer path hands the kernels bit-identical tracks.  (It matters: the
    waveform is ill-conditioned in the knot frequencies near turnovers of f_mn(t) -- a 1-ulp change moves individual bins
    by ~3e-9 of max|h| -- so parity is only meaningful on identical inputs.)"""
    om_phi, om_r = schwarzschild_frequencies(p, e, native=native)
    Msec = M * MTSUN_SI
    return om_phi / (2.0 * np.pi * Msec), om_r / (2.0 * np.pi * Msec)


def schwarzschild_frequencies(p, e, native=True):
    """Omega_phi, Omega_r (dimensionless, units of 1/M) for a bound Schwarzschild geodesic.

    Closed form in complete elliptic integrals K, E, Pi with parameter 4e/(p-6+2e); checked in
    tests against the chi-quadrature of dt/dchi and dphi/dchi (Darwin parametrisation).
    """
    p = np.asarray(p, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    if native and p.ndim >= 1 and _hostlib.load() is not None:
        return _hostlib.frequencies(p, e)          # csrc/emrihost.c: Carlson duplication, agrees to 2e-15
    m = 4.0 * e / (p - 6.0 + 2.0 * e)
    K = ellipk(m)
    E = ellipe(m)
    P1 = _ellip_pi(16.0 * e / (12.0 + 8.0 * e - 4.0 * e * e - 8.0 * p + p * p), m)
    P2 = _ellip_pi(2.0 * e * (p - 4.0) / ((1.0 + e) * (p - 6.0 + 2.0 * e)), m)
    p2 = p * p
    B = (
        (-2.0 * P2 * (6.0 + 2.0 * e - p) * (3.0 + e * e - p) * p2) / ((-1.0 + e) * (1.0 + e) ** 2)
        - (E * (-4.0 + p) * p2 * (-6.0 + 2.0 * e + p)) / (-1.0 + e * e)
        + (K * p2 * (28.0 + 4.0 * e * e - 12.0 * p + p2)) / (-1.0 + e * e)
        + (4.0 * (-4.0 + p) * p * (2.0 * (1.0 + e) * K + P2 * (-6.0 - 2.0 * e + p))) / (1.0 + e)
        + 2.0 * (-4.0 + p) ** 2 * (K * (-4.0 + p) + (P1 * p * (-6.0 - 2.0 * e + p)) / (2.0 + 2.0 * e - p))
    )
    D = (p - 2.0) ** 2 - 4.0 * e * e
    om_phi = 2.0 * p ** 1.5 / (np.sqrt(D) * (8.0 + B / (K * (p - 4.0) ** 2)))
    om_r = np.pi * p * np.sqrt((p - 6.0 + 2.0 * e) / D) / (8.0 * K + B / (p - 4.0) ** 2)
    return om_phi, om_r


def get_fundamental_frequencies(a, p, e, x):
    """(Omega_phi, Omega_theta, Omega_r), dimensionless.  Only a == 0 is supported (the
    FastSchwarzschildEccentricFlux model ignores spin: emri_pe.py:598,602)."""
    if np.any(np.asarray(a) != 0.0):
        raise ValueError("Only Schwarzschild (a = 0) frequencies are implemented on this path.")
    om_phi, om_r = schwarzschild_frequencies(p, e)
    return om_phi, om_phi.copy() if isinstance(om_phi, np.ndarray) else om_phi, om_r


def get_separatrix(a, e, x):
    """Schwarzschild separatrix p_s = 6 + 2e."""
    return 6.0 + 2.0 * np.asarray(e)


def get_p_at_t(traj_module, t_out, traj_args, index_of_p=3, index_of_a=2, index_of_e=4,
               index_of_x=5, traj_kwargs={}, xtol=2e-12, rtol=8.881784197001252e-16, bounds=None):
    """Find p0 such that the inspiral plunges at ``t_out`` years (call pattern of emri_pe.py:620-635,
    check_mode_by_mode.py:200-212).  ``traj_args`` = [M, mu, a, e0, x0] (p omitted)."""
    args = list(traj_args)
    e0 = args[index_of_e - 1]
    if bounds is None:
        bounds = [6.0 + 2.0 * e0 + 0.2, 60.0]

    def root_fn(p0):
        full = args[: index_of_p] + [p0] + args[index_of_p:]
        out = traj_module(*full, T=t_out * 100.0, **traj_kwargs)
        return out[0][-1] / YRSID_SI - t_out

    lo, hi = bounds
    return brentq(root_fn, lo, hi, xtol=xtol, rtol=max(rtol, 4 * np.finfo(float).eps))

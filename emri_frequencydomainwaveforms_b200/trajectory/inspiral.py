"""Host-side sparse inspiral trajectory (stand-in producer for the FD hot path).

The reference obtains (t, p, e, x, Phi_phi, Phi_theta, Phi_r) from
``few.trajectory.inspiral.EMRIInspiral(func="SchwarzEccFlux")`` (emri_pe.py:620,
Tutorial_FD_construction_single_mode.ipynb cell 5).  FEW's flux grid is a Zenodo download that is
not available offline, so this module integrates the same ODE *structure* -- (p, e, Phi_phi, Phi_r)
advanced with an adaptive Runge-Kutta stepper whose accepted steps ARE the sparse
trajectory, stopping 0.1 outside the separatrix -- with leading-order (Peters) fluxes and the exact
Schwarzschild geodesic frequencies.  Per BASELINE.json's north_star the trajectory stays on the host;
it is an input producer, not part of the accelerated path.  Any object returning the same 7-tuple
(e.g. the real FEW ``EMRIInspiral``) can be passed to the waveform classes instead.
"""
import numpy as np
from scipy.integrate import solve_ivp

from .. import _hostlib
from ..utils.constants import MTSUN_SI, YRSID_SI
from ..utils.utility import schwarzschild_frequencies

DIST_TO_SEPARATRIX = 0.1


def _rhs(_, y, q):
    p, e = y[0], y[1]
    e = max(e, 0.0)
    ome2 = 1.0 - e * e
    s = ome2 * np.sqrt(ome2)
    pdot = -(64.0 / 5.0) * q * s * (1.0 + 7.0 / 8.0 * e * e) / (p * p * p)
    edot = -(304.0 / 15.0) * q * e * s * (1.0 + 121.0 / 304.0 * e * e) / (p ** 4)
    om_phi, om_r = schwarzschild_frequencies(p, e)
    return [pdot, edot, float(om_phi), float(om_r)]


class EMRIInspiral:
    """Callable with the call surface of ``few.trajectory.inspiral.EMRIInspiral``.

    ``traj(M, mu, a, p0, e0, x0, Phi_phi0=0, Phi_theta0=0, Phi_r0=0, T=1.0, dt=10.0)`` returns
    ``(t, p, e, x, Phi_phi, Phi_theta, Phi_r)`` with ``t`` in seconds starting at 0.
    """

    def __init__(self, func="SchwarzEccFlux", rtol=1e-12, atol=1e-14, max_init_len=1000, use_native=True, **kwargs):
        if func != "SchwarzEccFlux":
            raise ValueError("Only func='SchwarzEccFlux' is available on this path.")
        self.rtol, self.atol, self.max_init_len = rtol, atol, max_init_len
        self.use_native = use_native   # csrc/emrihost.c (same ODE, same step control); False = SciPy twin

    def __call__(self, M, mu, a, p0, e0, x0, *args, Phi_phi0=0.0, Phi_theta0=0.0, Phi_r0=0.0,
                 T=1.0, dt=10.0, **kwargs):
        if len(args) >= 1:
            Phi_phi0 = args[0]
        if len(args) >= 3:
            Phi_r0 = args[2]
        if not (M > 0 and mu > 0):
            raise ValueError("Masses must be positive.")
        if e0 < 0.0 or e0 >= 1.0:
            raise ValueError("e0 must be in [0, 1).")
        if p0 < 6.0 + 2.0 * e0 + DIST_TO_SEPARATRIX:
            raise ValueError("p0 is inside the separatrix buffer (p0 < 6 + 2 e0 + 0.1).")
        if self.use_native and _hostlib.load() is not None:
            out, lens = _hostlib.trajectory_batch(M, mu, p0, e0, Phi_phi0, Phi_r0, T, self.rtol, self.atol, self.max_init_len)
            n = int(lens[0])
            if n == -2:
                raise ValueError("trajectory longer than max_init_len")
            if n < 2:
                raise ValueError("trajectory integration failed")
            t, p, e, Pp, Pr = (out[i][0, :n].copy() for i in range(5))
            return (t, p, e, np.ones_like(t), Pp, Pp.copy(), Pr)
        q = mu / M
        Msec = M * MTSUN_SI
        t_end = T * YRSID_SI / Msec

        def plunge(_, y, q):
            return y[0] - (6.0 + 2.0 * y[1] + DIST_TO_SEPARATRIX)

        plunge.terminal = True
        plunge.direction = -1
        sol = solve_ivp(_rhs, (0.0, t_end), [p0, e0, Phi_phi0, Phi_r0], method="RK45", args=(q,),
                        rtol=self.rtol, atol=self.atol, events=plunge)
        if not sol.success:
            raise ValueError("trajectory integration failed: " + str(sol.message))
        t = sol.t * Msec
        p, e, Phi_phi, Phi_r = sol.y
        if len(t) > self.max_init_len:
            raise ValueError("trajectory longer than max_init_len")
        x = np.ones_like(t)
        return (t, p.copy(), np.maximum(e, 0.0), x, Phi_phi.copy(), Phi_phi.copy(), Phi_r.copy())

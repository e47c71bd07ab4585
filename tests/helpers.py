"""Shared test helpers: synthetic walkers and the output-sizing rule (few SummationBase)."""
import numpy as np

from emri_frequencydomainwaveforms_b200.utils.constants import YRSID_SI

# (M, mu, p0, e0, theta, T[yr], eps) -- small enough for the binary128 oracle to finish in seconds
CASES = {
    "cfg1_like": (1e6, 10.0, 12.0, 0.35, 1.0, 0.05, 1e-2),      # slowly evolving, all modes monotone
    "plunge": (1e6, 50.0, 9.0, 0.3, 1.0, 0.25, 1e-2),            # reaches the separatrix buffer: turnover modes
    "face_on": (1e6, 10.0, 12.0, 0.35, np.pi, 0.05, 1e-2),       # the scripts' angles: theta = pi, only m = 2 mirrored terms
    "ecc_many": (5e5, 20.0, 10.5, 0.6, 0.7, 0.08, 1e-4),         # eps = 1e-4: ~100+ modes, negative-frequency harmonics
}


def grid_size(t, T, dt, pad_output=True, odd_len=True):
    n_pts = int(T * YRSID_SI / dt)
    Ts = n_pts * dt
    if Ts < t[-1]:
        num_pts, pad = int((Ts - t[0]) / dt) + 1, 0
    else:
        num_pts = int((t[-1] - t[0]) / dt) + 1
        pad = int((Ts - t[0]) / dt) + 1 - num_pts if pad_output else 0
    if odd_len and (num_pts + pad) % 2 == 0:
        pad += 1
    return num_pts + pad


def make_item(gen, name, dt=10.0, dist=1.0, **over):
    M, mu, p0, e0, theta, T, eps = CASES[name]
    it = gen.prepare(M, mu, p0, e0, theta, -np.pi / 2, dist=dist, Phi_phi0=over.get("Phi_phi0", 0.3),
                     Phi_r0=over.get("Phi_r0", 1.1), T=T, dt=dt, eps=over.get("eps", eps))
    it["T"], it["dt"] = T, dt
    it["N"] = grid_size(it["t"], T, dt)
    return it


def oracle_waveform(orc, it, N=None, val=None, fpos=None, **kw):
    N = it["N"] if N is None else N
    if fpos is None and val is None:
        val = 1.0 / (N * it["dt"])
    return orc.fd_sum(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"],
                      it["f_phi"], it["f_r"], N, val or 0.0, fpos, scale=kw.pop("scale", it["scale"]), **kw)


def rel_err(a, b):
    """per-bin error normalised by max|b| (the metric of SURVEY.md section 8d)."""
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

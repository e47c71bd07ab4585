"""Waveform generators with the call surface the reference scripts use.

* ``GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True,
  output_type="fd", odd_len=True), use_gpu=, return_list=)`` -- emri_pe.py:86-105,
  check_mode_by_mode.py:69-99; called as ``gen(M, mu, a, p0, e0, x0, dist, qS, phiS, qK, phiK,
  Phi_phi0, Phi_theta0, Phi_r0, T=, dt=, eps= | mode_selection=, f_arr=, mask_positive=,
  include_minus_m=)`` (emri_pe.py:140-155,212,241-242,349).
* ``gen.waveform_generator.create_waveform.frequency`` (emri_pe.py:238).

The accelerated part is ``create_waveform`` (FDInterpolatedModeSum: spline build, segmentation, SPA
mode sum, h+/hx split, distance scaling and SSB rotation fused into one CUDA pipeline).  Trajectory,
amplitudes, Ylm and mode selection are host-side producers; FEW's data-driven versions can be
plugged in through ``inspiral_generator`` / ``amplitude_generator``.
"""
import numpy as np

from .amplitude.synthetic import SyntheticAmplitude
from .summation.fdinterp import FDInterpolatedModeSum
from .trajectory.inspiral import EMRIInspiral
from .utils.constants import MRSUN_SI, MTSUN_SI, Gpc, YRSID_SI
from .utils.modeselector import ModeSelector
from .utils.utility import schwarzschild_frequencies
from .utils.ylm import GetYlms


class FastSchwarzschildEccentricFlux:
    """Schwarzschild eccentric FD waveform in the source frame (mirror of few.waveform's class)."""

    descriptor = "eccentric"
    background = "Schwarzschild"
    frame = "source"

    def __init__(self, inspiral_kwargs={}, amplitude_kwargs={}, Ylm_kwargs={}, sum_kwargs={}, use_gpu=True,
                 inspiral_generator=None, amplitude_generator=None, *args, **kwargs):
        sk = dict(sum_kwargs)
        sk.pop("use_gpu", None)
        if sk.get("output_type", "td") != "fd":
            raise ValueError("This package implements the frequency-domain path only: pass "
                             "sum_kwargs=dict(output_type='fd', ...).")
        self.inspiral_generator = inspiral_generator or EMRIInspiral(func="SchwarzEccFlux")
        self.amplitude_generator = amplitude_generator or SyntheticAmplitude()
        amp = self.amplitude_generator
        self.l_arr, self.m_arr, self.n_arr = amp.l_arr, amp.m_arr, amp.n_arr
        self.unique_l, self.unique_m, self.inverse_lm = amp.unique_l, amp.unique_m, amp.inverse_lm
        self.m0mask = self.m_arr != 0
        self.num_teuk_modes = len(self.l_arr)
        self.ylm_gen = GetYlms(assume_positive_m=True)
        self.mode_selector = ModeSelector(self.m0mask)
        self.create_waveform = FDInterpolatedModeSum(**sk)
        self.inspiral_kwargs = dict(inspiral_kwargs)
        for k in ("DENSE_STEPPING", "max_init_len", "use_rk4"):
            self.inspiral_kwargs.pop(k, None)

    # FEW's sanity checks raise ValueError (check_mode_by_mode.py:218-219 relies on that)
    @staticmethod
    def sanity_check_init(M, mu, p0, e0):
        if e0 > 0.75:
            raise ValueError(f"Initial eccentricity above 0.75 not allowed. (e0={e0})")
        if e0 < 0.0:
            raise ValueError(f"Initial eccentricity below 0.0 not physical. (e0={e0})")
        if mu / M > 1e-4 * (1 + 1e-9):
            import warnings
            warnings.warn(f"Mass ratio is outside of generally accepted range for an extreme mass ratio (1e-4). (q={mu / M})")
        if p0 < 10.0 and not (p0 >= 7.2 + 2 * e0):
            raise ValueError(f"This p0 ({p0}) and e0 ({e0}) combination is outside of our domain of validity.")
        if p0 > 16.0 + 2 * e0:
            raise ValueError(f"Initial p0 is too large (p0={p0}). Must be 10 <= p0 <= 16 + 2 * e.")

    @staticmethod
    def sanity_check_viewing_angles(theta, phi):
        if theta < 0.0 or theta > np.pi:
            raise ValueError("theta must be between 0 and pi.")
        return theta, phi % (2 * np.pi) if phi < 0 else phi

    def prepare(self, M, mu, p0, e0, theta, phi, dist=None, Phi_phi0=0.0, Phi_r0=0.0, T=1.0, dt=10.0,
                eps=1e-5, mode_selection=None):
        """Host-side producers of one walker: trajectory -> amplitudes -> Ylm -> mode selection.
        Returns the dict FDInterpolatedModeSum / engine.PackedBatch consume."""
        self.sanity_check_init(M, mu, p0, e0)
        t, p, e, x, Phi_phi, Phi_theta, Phi_r = self.inspiral_generator(
            M, mu, 0.0, p0, e0, 1.0, Phi_phi0=Phi_phi0, Phi_theta0=0.0, Phi_r0=Phi_r0, T=T, dt=dt,
            **self.inspiral_kwargs)
        teuk = self.amplitude_generator(p, e)
        nl = len(self.unique_l)
        y = self.ylm_gen(self.unique_l, self.unique_m, theta, phi)
        ylms = np.concatenate([y[:nl][self.inverse_lm], y[nl:][self.inverse_lm][self.m0mask]])
        if mode_selection is None or isinstance(mode_selection, str):
            if isinstance(mode_selection, str) and mode_selection != "all":
                raise ValueError("If mode selection is a string, must be `all`.")
            if mode_selection == "all":
                keep = np.arange(self.num_teuk_modes)
                pos = np.cumsum(self.m0mask) - 1
                neg = np.where(self.m0mask, self.num_teuk_modes + pos, keep)
                tm, yk = teuk, np.concatenate([ylms[keep], ylms[neg]])
                ls, ms, ns = self.l_arr, self.m_arr, self.n_arr
            else:
                tm, yk, ls, ms, ns = self.mode_selector(teuk, ylms, [self.l_arr, self.m_arr, self.n_arr], eps=eps)
        else:
            if len(mode_selection) == 0:
                raise ValueError("If mode selection is a list, cannot be empty.")
            keep, yp, ym = [], [], []
            for (l, m, n) in mode_selection:
                if m < 0:
                    l, m, n = l, -m, -n    # the stored partner carries the -m term
                idx = np.where((self.l_arr == l) & (self.m_arr == m) & (self.n_arr == n))[0]
                if len(idx) == 0:
                    raise ValueError(f"mode {(l, m, n)} is not in the amplitude basis")
                keep.append(idx[0])
            keep = np.unique(np.asarray(keep))
            pos = np.cumsum(self.m0mask) - 1
            neg = np.where(self.m0mask[keep], self.num_teuk_modes + pos[keep], keep)
            tm, yk = teuk[:, keep], np.concatenate([ylms[keep], ylms[neg]])
            ls, ms, ns = self.l_arr[keep], self.m_arr[keep], self.n_arr[keep]
        om_phi, om_r = schwarzschild_frequencies(p, e)
        scale = 1.0 if dist is None else (mu * MRSUN_SI) / (dist * Gpc)
        self.ls, self.ms, self.ns = ls, ms, ns
        self.num_modes_kept = len(ls)
        return dict(t=t, p=p, e=e, teuk_modes=np.ascontiguousarray(tm), ylms=yk, Phi_phi=Phi_phi, Phi_r=Phi_r,
                    m_arr=ms.astype(np.int32), n_arr=ns.astype(np.int32), l_arr=ls.astype(np.int32),
                    f_phi=om_phi / (2 * np.pi * M * MTSUN_SI), f_r=om_r / (2 * np.pi * M * MTSUN_SI),
                    scale=scale, M=M, mu=mu)

    def __call__(self, M, mu, p0, e0, theta, phi, *args, dist=None, Phi_phi0=0.0, Phi_r0=0.0, dt=10.0, T=1.0,
                 eps=1e-5, show_progress=False, batch_size=-1, mode_selection=None, include_minus_m=True,
                 f_arr=None, mask_positive=False, cos2psi=1.0, sin2psi=0.0, **kwargs):
        theta, phi = self.sanity_check_viewing_angles(theta, phi)
        it = self.prepare(M, mu, p0, e0, theta, phi, dist=dist, Phi_phi0=Phi_phi0, Phi_r0=Phi_r0, T=T, dt=dt,
                          eps=eps, mode_selection=mode_selection)
        return self.create_waveform(
            it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"], M,
            it["p"], it["e"], T=T, dt=dt, include_minus_m=include_minus_m, f_arr=f_arr,
            mask_positive=mask_positive, scale=it["scale"], cos2psi=cos2psi, sin2psi=sin2psi)


def viewing_angles(qS, phiS, qK, phiK):
    """Source-frame viewing angles (theta, phi) from SSB sky/spin angles; mirror of
    GenerateEMRIWaveform._get_viewing_angles: theta = arccos(-R.S), phi = -pi/2."""
    R = np.array([np.sin(qS) * np.cos(phiS), np.sin(qS) * np.sin(phiS), np.cos(qS)])
    S = np.array([np.sin(qK) * np.cos(phiK), np.sin(qK) * np.sin(phiK), np.cos(qK)])
    theta = np.arccos(np.clip(-np.dot(R, S), -1.0, 1.0))
    return theta, -np.pi / 2.0


def polarization_angle(qS, phiS, qK, phiK):
    """psi = -atan2(cos qS sin qK cos(phiS - phiK) - cos qK sin qS, sin qK sin(phiS - phiK))."""
    up = np.cos(qS) * np.sin(qK) * np.cos(phiS - phiK) - np.cos(qK) * np.sin(qS)
    dw = np.sin(qK) * np.sin(phiS - phiK)
    return -np.arctan2(up, dw) if dw != 0.0 else 0.5 * np.pi


class GenerateEMRIWaveform:
    """Generic SSB-frame wrapper (mirror of few.waveform.GenerateEMRIWaveform) for the FD model."""

    def __init__(self, waveform_class, *args, frame="detector", return_list=False, use_gpu=True, **kwargs):
        if isinstance(waveform_class, str):
            if waveform_class != "FastSchwarzschildEccentricFlux":
                raise ValueError("Only 'FastSchwarzschildEccentricFlux' is available on this path.")
            waveform_class = FastSchwarzschildEccentricFlux
        self.waveform_generator = waveform_class(*args, use_gpu=use_gpu, **kwargs)
        self.frame, self.return_list = frame, return_list
        self.phases_needed = {"Phi_phi0": 11, "Phi_r0": 13}

    def _transform(self, qS, phiS, qK, phiK):
        theta, phi = viewing_angles(qS, phiS, qK, phiK)
        if self.frame == "detector":
            psi = polarization_angle(qS, phiS, qK, phiK)
            return theta, phi, np.cos(2.0 * psi), np.sin(2.0 * psi)
        return theta, phi, 1.0, 0.0

    def __call__(self, M, mu, a, p0, e0, x0, dist, qS, phiS, qK, phiK, Phi_phi0, Phi_theta0, Phi_r0, *args,
                 **kwargs):
        # the Schwarzschild model ignores a, x0, Phi_theta0 (emri_pe.py:598,602)
        theta, phi, c2, s2 = self._transform(qS, phiS, qK, phiK)
        h = self.waveform_generator(M, mu, p0, e0, theta, phi, *args, dist=dist, Phi_phi0=Phi_phi0,
                                    Phi_r0=Phi_r0, cos2psi=c2, sin2psi=s2, **kwargs)
        hp, hc = h[0], h[1]
        if self.return_list:
            return [hp, hc]
        return hp - 1j * hc   # check_mode_by_mode.py:247: h = h+ - i hx

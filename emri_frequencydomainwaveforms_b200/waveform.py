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
from .utils.utility import fundamental_frequencies_hz
from .utils.ylm import GetYlms


_warned_standins = False


def _warn_standins(traj, amp):
    """One warning per process: the class keeps FEW's name, but its DEFAULT producers are offline stand-ins."""
    global _warned_standins
    if _warned_standins:
        return
    _warned_standins = True
    import warnings
    what = " and ".join(w for w, on in (("trajectory (leading-order Peters fluxes + exact geodesic frequencies)", traj),
                                        ("amplitudes (SyntheticAmplitude, FEW's 3843-mode layout)", amp)) if on)
    warnings.warn("FastSchwarzschildEccentricFlux is using this package's stand-in " + what + ": FastEMRIWaveforms' flux grid and "
                  "ROMAN amplitude weights are Zenodo downloads that are not available here, so waveforms differ physically from "
                  "few's.  Pass inspiral_generator= / amplitude_generator= (any objects with few's call signatures) to use the "
                  "real producers; the frequency-domain summation and likelihood below them are unchanged.", UserWarning, stacklevel=3)


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
        if inspiral_generator is None or amplitude_generator is None:
            _warn_standins(inspiral_generator is None, amplitude_generator is None)
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
        producers = kwargs.pop("producers", "auto")
        if producers == "auto":
            from . import _hostlib
            producers = ("device" if hasattr(amp, "device_call") and getattr(self.inspiral_generator, "use_native", False)
                         and _hostlib.load() is not None else "host")
        self.producers = producers   # "device": batched CUDA producers for Ylm / mode selection; "host": NumPy
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
        f_phi, f_r = fundamental_frequencies_hz(p, e, M)
        scale = 1.0 if dist is None else (mu * MRSUN_SI) / (dist * Gpc)
        self.ls, self.ms, self.ns = ls, ms, ns
        self.num_modes_kept = len(ls)
        return dict(t=t, p=p, e=e, teuk_modes=np.ascontiguousarray(tm), ylms=yk, Phi_phi=Phi_phi, Phi_r=Phi_r,
                    m_arr=ms.astype(np.int32), n_arr=ns.astype(np.int32), l_arr=ls.astype(np.int32),
                    f_phi=f_phi, f_r=f_r,
                    scale=scale, M=M, mu=mu)

    # ---- batched producers with Ylm, mode selection and compaction on the device ----
    def _device_basis(self, handle):
        import torch
        dev = handle.torch_device
        if getattr(self, "_basis_dev", None) is None or self._basis_dev["l"].device != dev:
            up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
            neg_src = np.where(self.m0mask)[0]
            neg_pos = np.where(self.m0mask, np.cumsum(self.m0mask) - 1, -1)
            self._basis_dev = dict(l=up(self.l_arr), m=up(self.m_arr), n=up(self.n_arr), neg_src=up(neg_src), neg_pos=up(neg_pos))
        return self._basis_dev

    def prepare_batch_device(self, M, mu, p0, e0, theta, phi, dist=None, Phi_phi0=0.0, Phi_r0=0.0, T=1.0, dt=10.0,
                             eps=1e-5, cos2psi=1.0, sin2psi=0.0, handle=None, nthreads=None, keep_full=False):
        """Producers of a whole walker batch: trajectories on the host (native, threaded; north_star keeps the ODE
        there), amplitudes / Ylm / mode selection / compaction on the device.  Arguments are arrays [nb] (scalars
        broadcast).  Returns ``(DeviceBatch or None, ok[nb])``; walkers whose parameters are outside the domain of
        validity or whose trajectory fails have ok = False (the single-walker path raises ValueError for those)."""
        import os
        import torch
        from . import _hostlib, _lib, engine
        from .utils.ylm import ylm_batch_device
        h = handle or _lib.get_handle()
        dev = h.torch_device
        amp = self.amplitude_generator
        if not hasattr(amp, "device_call"):
            raise ValueError("prepare_batch_device needs an amplitude generator with device_call(p, e, device)")
        if _hostlib.load() is None or not getattr(self.inspiral_generator, "use_native", False):
            raise ValueError("prepare_batch_device needs the native trajectory library (csrc/libemrihost.so)")
        bc = lambda x: np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), np.shape(M)).ravel())
        M = np.atleast_1d(np.asarray(M, dtype=np.float64))
        mu, p0, e0, theta, phi, Phi_phi0, Phi_r0, cos2psi, sin2psi = map(bc, (mu, p0, e0, theta, phi, Phi_phi0, Phi_r0, cos2psi, sin2psi))
        nb = len(M)
        # the domain checks of sanity_check_init / sanity_check_viewing_angles, vectorised (same conditions)
        ok = ~((e0 > 0.75) | (e0 < 0.0) | ((p0 < 10.0) & ~(p0 >= 7.2 + 2.0 * e0)) | (p0 > 16.0 + 2.0 * e0)
               | (theta < 0.0) | (theta > np.pi) | ~(M > 0.0) | ~(mu > 0.0))
        ok &= np.isfinite(M) & np.isfinite(mu) & np.isfinite(p0) & np.isfinite(e0) & np.isfinite(theta) & np.isfinite(phi)
        ig = self.inspiral_generator
        if not nthreads:    # this process's share of the host cores (one process per GPU: torchrun exports LOCAL_WORLD_SIZE)
            nthreads = max(1, min(len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))), 16))
        sel = np.where(ok)[0]
        if len(sel) == 0:
            return None, ok
        import time as _time
        _t0 = _time.perf_counter()
        out, lens = _hostlib.trajectory_batch(M[sel], mu[sel], p0[sel], e0[sel], Phi_phi0[sel], Phi_r0[sel], T, ig.rtol, ig.atol,
                                              ig.max_init_len, nthreads=nthreads)
        _stage = (nthreads, round(1e3 * (_time.perf_counter() - _t0), 2))     # (threads, host trajectory ms): diagnostics
        good = lens >= 4
        ok[sel[~good]] = False
        sel, out, lens = sel[good], [o[good] for o in out], lens[good].astype(np.int64)
        B = len(sel)
        if B == 0:
            return None, ok
        # ragged pack of the tracks: row mask of the [B, max_len] arrays
        mask = np.arange(out[0].shape[1])[None, :] < lens[:, None]
        tr = np.stack([o[mask] for o in out])             # [7, sum L]: t, p, e, Phi_phi, Phi_r, f_phi, f_r
        nk = tr.shape[1]
        tr_dev = torch.from_numpy(tr).to(dev)
        samp_walker = torch.from_numpy(np.repeat(np.arange(B, dtype=np.int32), lens)).to(dev)
        basis = self._device_basis(h)
        Mb, Mneg = self.num_teuk_modes, int(self.m0mask.sum())
        teuk_full = amp.device_call(tr_dev[1], tr_dev[2], dev, handle=h)             # [sum L, Mb] complex128
        ylm_full = ylm_batch_device(basis["l"], basis["m"], basis["neg_src"], theta[sel], phi[sel], h, lmax=int(self.l_arr.max()))
        flags = torch.empty((B, Mb), dtype=torch.uint8, device=dev)
        h.check(h.lib.emrifd_mode_select(h.h, teuk_full.data_ptr(), nk, Mb, samp_walker.data_ptr(), ylm_full.data_ptr(),
                                         basis["neg_src"].data_ptr(), Mneg, B, float(eps), flags.data_ptr()))
        keep_idx = torch.empty((B, Mb), dtype=torch.int32, device=dev)
        K_dev = torch.empty(B, dtype=torch.int32, device=dev)
        h.check(h.lib.emrifd_mode_compact_count(h.h, flags.data_ptr(), B, Mb, keep_idx.data_ptr(), K_dev.data_ptr()))
        K = K_dev.cpu().numpy().astype(np.int64)          # the one D2H of the producer stage: B ints
        w = np.zeros(B, dtype=_lib.WALKER_DTYPE)
        w["L"], w["K"] = lens, K
        w["knot_off"] = np.concatenate([[0], np.cumsum(lens)[:-1]])
        w["teuk_off"] = np.concatenate([[0], np.cumsum(lens * K)[:-1]])
        w["mode_off"] = np.concatenate([[0], np.cumsum(K)[:-1]])
        w["coeff_off"] = np.concatenate([[0], np.cumsum(lens * (2 * K + 4) * 4)[:-1]])
        w["scale"] = 1.0 if dist is None else (mu[sel] * MRSUN_SI) / (bc(dist)[sel] * Gpc)
        w["cos2psi"], w["sin2psi"] = cos2psi[sel], sin2psi[sel]
        teuk = torch.empty(int((lens * K).sum()), dtype=torch.complex128, device=dev)
        m_out = torch.empty(int(K.sum()), dtype=torch.int32, device=dev)
        n_out = torch.empty_like(m_out)
        ylm = torch.empty(2 * int(K.sum()), dtype=torch.complex128, device=dev)
        h.check(h.lib.emrifd_mode_compact_gather(h.h, w.ctypes.data, B, teuk_full.data_ptr(), Mb, Mneg, keep_idx.data_ptr(),
                                                 basis["m"].data_ptr(), basis["n"].data_ptr(), basis["neg_pos"].data_ptr(),
                                                 ylm_full.data_ptr(), teuk.data_ptr(), m_out.data_ptr(), n_out.data_ptr(),
                                                 ylm.data_ptr()))
        tracks = dict(t=tr_dev[0], Phi_phi=tr_dev[3], Phi_r=tr_dev[4], f_phi=tr_dev[5], f_r=tr_dev[6])
        db = engine.DeviceBatch.from_device_parts(h, w, tracks, teuk, m_out, n_out, ylm)
        db.keep_idx, db.h2d_bytes = keep_idx, int(tr.nbytes + samp_walker.numel() * 4 + 2 * 8 * B + w.nbytes)
        db.stage_ms = _stage
        db.p_e_host = (tr[1], tr[2])
        ends = np.cumsum(lens) - 1
        db.t_first, db.t_last = tr[0][ends - lens + 1], tr[0][ends]      # per walker: output sizing (few SummationBase)
        if keep_full:   # tests compare the selection against the numpy restatement on the very same inputs
            db.teuk_full, db.ylm_full, db.flags = teuk_full, ylm_full, flags
        return db, ok

    def __call__(self, M, mu, p0, e0, theta, phi, *args, dist=None, Phi_phi0=0.0, Phi_r0=0.0, dt=10.0, T=1.0,
                 eps=1e-5, show_progress=False, batch_size=-1, mode_selection=None, include_minus_m=True,
                 f_arr=None, mask_positive=False, cos2psi=1.0, sin2psi=0.0, **kwargs):
        theta, phi = self.sanity_check_viewing_angles(theta, phi)
        if mode_selection is None and self.producers == "device":
            # Ylm / mode selection / compaction on the device (the NumPy producers cost ~30 ms per waveform)
            db, ok = self.prepare_batch_device(M, mu, p0, e0, theta, phi, dist=dist, Phi_phi0=Phi_phi0, Phi_r0=Phi_r0, T=T, dt=dt,
                                               eps=eps, cos2psi=cos2psi, sin2psi=sin2psi, handle=self.create_waveform.handle)
            if db is not None:
                out = self.create_waveform.sum_device_batch(db, db.t_first[0], db.t_last[0], T=T, dt=dt, include_minus_m=include_minus_m,
                                                            f_arr=f_arr, mask_positive=mask_positive)
                self.num_modes_kept = int(db.pb.walkers["K"][0])
                self._last_device_batch = db
                return out[0]
            # invalid parameters: fall through so that the host producers raise FEW's ValueError
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
    # explicit sum (not np.dot): the batched twin ssb_transform_batch evaluates the same expression, bit for bit --
    # near theta = pi (the scripts' geometry) arccos turns a 1-ulp difference of R.S into 1.5e-8 rad
    dot = (R[0] * S[0] + R[1] * S[1]) + R[2] * S[2]
    theta = np.arccos(np.clip(-dot, -1.0, 1.0))
    return theta, -np.pi / 2.0


def polarization_angle(qS, phiS, qK, phiK):
    """psi = -atan2(cos qS sin qK cos(phiS - phiK) - cos qK sin qS, sin qK sin(phiS - phiK))."""
    up = np.cos(qS) * np.sin(qK) * np.cos(phiS - phiK) - np.cos(qK) * np.sin(qS)
    dw = np.sin(qK) * np.sin(phiS - phiK)
    return -np.arctan2(up, dw) if dw != 0.0 else 0.5 * np.pi


def ssb_transform_batch(qS, phiS, qK, phiK, detector_frame=True):
    """Vectorised ``GenerateEMRIWaveform._transform``: arrays [nb] -> (theta, phi, cos2psi, sin2psi) arrays."""
    qS, phiS, qK, phiK = (np.asarray(a, dtype=np.float64) for a in (qS, phiS, qK, phiK))
    dot = (((np.sin(qS) * np.cos(phiS)) * (np.sin(qK) * np.cos(phiK)) + (np.sin(qS) * np.sin(phiS)) * (np.sin(qK) * np.sin(phiK)))
           + np.cos(qS) * np.cos(qK))
    theta = np.arccos(np.clip(-dot, -1.0, 1.0))
    phi = np.full_like(theta, -np.pi / 2.0)
    if not detector_frame:
        return theta, phi, np.ones_like(theta), np.zeros_like(theta)
    up = np.cos(qS) * np.sin(qK) * np.cos(phiS - phiK) - np.cos(qK) * np.sin(qS)
    dw = np.sin(qK) * np.sin(phiS - phiK)
    psi = np.where(dw != 0.0, -np.arctan2(up, dw), 0.5 * np.pi)
    return theta, phi, np.cos(2.0 * psi), np.sin(2.0 * psi)


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

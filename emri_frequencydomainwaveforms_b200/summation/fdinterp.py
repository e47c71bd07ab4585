"""``FDInterpolatedModeSum`` -- mirror of ``few.summation.fdinterp.FDInterpolatedModeSum``, the
frequency-domain mode summation behind ``GenerateEMRIWaveform(..., sum_kwargs={'output_type': 'fd'})``
(emri_pe.py:86-105).  Same call surface: ``__call__`` sizes the output from
(T, dt, pad_output, odd_len) like few's SummationBase, ``sum`` does the work, ``frequency`` holds the
grid (emri_pe.py:238), ``waveform`` the stacked (h+, hx) result.  All arithmetic runs in the CUDA
kernels behind include/emrifd.h; there is no CPU path.
"""
import numpy as np

from .. import _lib, engine
from ..utils.constants import MTSUN_SI, YRSID_SI
from ..utils.utility import fundamental_frequencies_hz


def _np(x):
    try:
        import torch
        if torch.is_tensor(x):
            return x.detach().cpu().numpy()
    except ImportError:
        pass
    if hasattr(x, "get"):
        return x.get()
    return np.asarray(x)


class FDInterpolatedModeSum:
    def __init__(self, pad_output=False, output_type="fd", odd_len=False, use_gpu=True, device=None, **kwargs):
        if output_type != "fd":
            raise ValueError("This path implements output_type='fd' only (time-domain sums are out of scope).")
        self.pad_output, self.output_type, self.odd_len = pad_output, output_type, odd_len
        self.use_gpu = True
        self._device = device
        self.frequency = None
        self.waveform = None
        self.num_pts = self.num_pts_pad = 0
        self.last_batch = None

    @property
    def handle(self):
        return _lib.get_handle(self._device)

    def __getstate__(self):
        # picklable by reconstruction (the reference forks pool workers that each own a generator, emri_pe.py:545):
        # device buffers and the handle are per process and are rebuilt lazily
        d = dict(self.__dict__)
        d.update(frequency=None, waveform=None, last_batch=None)
        return d

    # -- few.utils.baseclasses.SummationBase.__call__ (output sizing) ------------
    def _size_output(self, t_first, t_last, T, dt):
        n_pts = int(T * YRSID_SI / dt)
        T_s = n_pts * dt
        if T_s < t_last:
            num_pts = int((T_s - t_first) / dt) + 1
            num_pts_pad = 0
        else:
            num_pts = int((t_last - t_first) / dt) + 1
            num_pts_pad = int((T_s - t_first) / dt) + 1 - num_pts if self.pad_output else 0
        if self.odd_len and (num_pts + num_pts_pad) % 2 == 0:
            num_pts_pad += 1
        self.num_pts, self.num_pts_pad, self.dt = num_pts, num_pts_pad, dt

    def _grid(self, f_arr, dt):
        """A1: (N, val, fpos_dev) and ``self.frequency`` for an explicit two-sided f_arr or the implicit fftfreq grid."""
        import torch
        h = self.handle
        if f_arr is not None:
            f_host = _np(f_arr)
            N, fpos = engine.grid_from_frequency(f_host)
            self.frequency = torch.as_tensor(f_host, dtype=torch.float64).to(h.torch_device)
            return N, 0.0, self.frequency[(N - 1) // 2:].contiguous()
        N = self.num_pts + self.num_pts_pad
        if N % 2 == 0 or N < 3:
            raise ValueError("The frequency grid must have odd length: use sum_kwargs=dict(odd_len=True).")
        val = 1.0 / (N * dt)
        k = torch.arange(-(N - 1) // 2, (N - 1) // 2 + 1, dtype=torch.float64, device=h.torch_device)
        self.frequency = k * val   # == fftshift(fftfreq(N, dt)) bit for bit (numpy multiplies k by 1/(N dt))
        return N, val, None

    def sum_device_batch(self, db, t_first, t_last, T=1.0, dt=10.0, include_minus_m=True, f_arr=None, mask_positive=False):
        """Same as ``__call__`` for a batch whose packed inputs already live on the device
        (``FastSchwarzschildEccentricFlux.prepare_batch_device``).  Returns ``[B, 2, n_out]`` (row w = vstack((h+, hx)))."""
        self._size_output(float(t_first), float(t_last), T, dt)
        N, val, fpos_dev = self._grid(f_arr, dt)
        engine.run_waveform(db, N, val, fpos_dev, include_minus_m=include_minus_m, mask_positive=mask_positive)
        self.handle.status()
        self.last_batch = db
        self.waveform = db.last_out[0]
        return db.last_out

    def __call__(self, t, *args, T=1.0, dt=10.0, **kwargs):
        t_host = _np(t)
        self._size_output(t_host[0].item(), t_host[-1].item(), T, dt)
        self.sum(t, *args, dt=dt, **kwargs)
        return self.waveform

    # -- FDInterpolatedModeSum.sum ---------------------------------------------------------------
    def sum(self, t, teuk_modes, ylms, Phi_phi, Phi_r, m_arr, n_arr, M, p, e, *args, include_minus_m=True,
            separate_modes=False, dt=10.0, f_arr=None, mask_positive=False, scale=1.0, cos2psi=1.0, sin2psi=0.0,
            **kwargs):
        import torch
        if separate_modes:
            raise ValueError("separate_modes is not available on this path.")
        h = self.handle
        t_h, p_h, e_h = _np(t).astype(np.float64), _np(p).astype(np.float64), _np(e).astype(np.float64)
        # A2: Schwarzschild fundamental frequencies at the sparse points (host, L values)
        f_phi, f_r = fundamental_frequencies_hz(p_h, e_h, M)
        item = dict(t=t_h, teuk_modes=_np(teuk_modes), ylms=_np(ylms), Phi_phi=_np(Phi_phi), Phi_r=_np(Phi_r),
                    m_arr=_np(m_arr), n_arr=_np(n_arr), f_phi=f_phi, f_r=f_r, scale=scale, cos2psi=cos2psi,
                    sin2psi=sin2psi)
        pb = engine.PackedBatch([item])
        N, val, fpos_dev = self._grid(f_arr, dt)   # A1: frequency grid
        db = engine.DeviceBatch(pb, h)
        hp, hc, _ = engine.run_waveform(db, N, val, fpos_dev, include_minus_m=include_minus_m,
                                        mask_positive=mask_positive)
        h.status()
        self.last_batch = db
        self.waveform = db.last_out[0]   # vstack((h+, hx)): [2, N] or [2, (N+1)/2], written in place by the kernel
        return self.waveform

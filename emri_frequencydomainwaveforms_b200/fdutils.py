"""FD adapters of the reference's FDutils.py that sit on the hot path.

* ``get_sensitivity(f)``  -- FDutils.py:4-5,21-33: not-a-knot cubic spline through the
  LISA_Alloc_Sh table, evaluated WITH extrapolation (f = 0 included).  Built and evaluated on the GPU
  with the same spline kernels as the waveform (data/lisa_alloc_sh.npy is the table, converted by
  scripts/convert_reference_data.py).
* ``get_fd_waveform_fromFD`` -- FDutils.py:105-139: call the generator, keep f >= 0, zero outside
  ``non_zero_mask``.  When the mask is exactly ``frequency >= 0`` the generator is asked for
  ``mask_positive=True`` so no boolean gather pass is needed.
* ``get_convolution`` / ``get_fd_windowed`` -- FDutils.py:35-47,66-101 (the step right
  after the path when ``window_flag=1``): the reference's ``convolve(hstack((a[1:], a)), b, 'valid')/len(b)`` is
  the circular convolution (a (*) b)/N, an O(N^2) direct sum there; here it is three cuFFT calls
  (``torch.fft``; a library FFT, as in the reference's own commented FFT variant FDutils.py:83-85).
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_psd_spline = None


def _spline():
    global _psd_spline
    if _psd_spline is None:
        from .summation.interpolatedmodesum import CubicSplineInterpolant
        S = np.load(os.path.join(_DATA, "lisa_alloc_sh.npy"))
        _psd_spline = CubicSplineInterpolant(S[:, 0], S[:, 1])
    return _psd_spline


def get_sensitivity(f):
    """LISA sensitivity S_n(f) [s]; returns the input's array type (numpy in -> numpy out)."""
    import torch
    out = _spline()(f)[0]
    if torch.is_tensor(f):
        return out
    return out.cpu().numpy() if np.ndim(f) else float(out.cpu().numpy())


def _as_dev(x, dtype):
    import torch
    from . import _lib
    dev = _lib.get_handle().torch_device
    if torch.is_tensor(x):
        return x.to(device=dev, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(dev)


def get_convolution(a, b):
    """``convolve(hstack((a[1:], a)), b, mode='valid') / len(b)`` (FDutils.py:35-47) for equal-length 1D arrays:
    out[k] = (1/N) sum_j a[(k - j) mod N] b[j], evaluated as ifft(fft(a) fft(b)) / N on the GPU."""
    import torch
    a = _as_dev(a, torch.complex128)
    b = _as_dev(b, torch.complex128)
    if a.ndim != 1 or a.shape != b.shape:
        raise ValueError("get_convolution needs two 1D arrays of equal length.")
    return torch.fft.ifft(torch.fft.fft(a) * torch.fft.fft(b)) / a.shape[0]


def get_fd_windowed(signal, window, window_in_fd=False):
    """Convolve the FD channels [h+, hx] with the DFT of a time-domain window (FDutils.py:66-101)."""
    import torch
    if window is None:
        return [signal[0], signal[1]]
    fft_window = _as_dev(window, torch.complex128)
    if not window_in_fd:
        fft_window = torch.fft.fft(fft_window)
    cw = torch.conj(fft_window)
    return [get_convolution(cw, signal[0]), get_convolution(cw, signal[1])]


class get_fd_waveform_fromFD:
    def __init__(self, waveform_generator, positive_frequency_mask, dt, non_zero_mask=None, window=None,
                 window_in_fd=False):
        self.waveform_generator = waveform_generator
        self.positive_frequency_mask = positive_frequency_mask
        self.non_zero_mask = non_zero_mask
        self.window = window
        self.window_in_fd = window_in_fd
        m = np.asarray(positive_frequency_mask.cpu() if hasattr(positive_frequency_mask, "cpu")
                       else positive_frequency_mask)
        n = len(m)
        self._is_upper_half = bool(n % 2 == 1 and m[(n - 1) // 2:].all() and not m[: (n - 1) // 2].any())

    def __call__(self, *args, **kwargs):
        import torch
        if self.window is None and self._is_upper_half and not kwargs.get("mask_positive", False):
            ch1, ch2 = self.waveform_generator(*args, mask_positive=True, **kwargs)
        else:
            chans = get_fd_windowed(self.waveform_generator(*args, **kwargs), self.window, window_in_fd=self.window_in_fd)
            pm = self.positive_frequency_mask
            mask = torch.as_tensor(np.asarray(pm.cpu() if hasattr(pm, "cpu") else pm), device=chans[0].device)
            ch1, ch2 = chans[0][mask], chans[1][mask]
        if self.non_zero_mask is not None:
            nz = torch.as_tensor(np.asarray(self.non_zero_mask.cpu() if hasattr(self.non_zero_mask, "cpu")
                                            else self.non_zero_mask), device=ch1.device)
            ch1 = torch.where(nz, ch1, torch.zeros_like(ch1))
            ch2 = torch.where(nz, ch2, torch.zeros_like(ch2))
        return [ch1, ch2]


# ---- TD <-> FD comparison utilities (FDutils.py:49-64,142-178) -----------------------
def get_fft_td_windowed(signal, window, dt):
    """``fftshift(fft(signal[i] * window)) * dt`` for the two polarisations (FDutils.py:49-64), as cuFFT calls."""
    import torch
    w = _as_dev(window, torch.float64)
    out = []
    for s in signal[:2]:
        s = _as_dev(s, torch.complex128)
        if s.ndim != 1 or s.shape != w.shape:
            raise ValueError("signal channels and window must be 1D arrays of equal length.")
        out.append(torch.fft.fftshift(torch.fft.fft(s * w)) * dt)
    return out


class get_fd_waveform_fromTD:
    """Frequency-domain channels from a TIME-domain generator (FDutils.py:142-178): FFT of the windowed channels, f >= 0
    kept, zero outside ``non_zero_mask``.  The TD generator itself (few's TD summation) is outside this package; any
    callable returning ``[h+, hx]`` (or ``h+ - i hx``) on a uniform grid works."""

    def __init__(self, waveform_generator, positive_frequency_mask, dt, non_zero_mask=None, window=None):
        self.waveform_generator = waveform_generator
        self.positive_frequency_mask = positive_frequency_mask
        self.dt = dt
        self.non_zero_mask = non_zero_mask
        pm = positive_frequency_mask
        pm = np.asarray(pm.cpu() if hasattr(pm, "cpu") else pm)
        self.window = np.ones(len(pm)) if window is None else window

    def __call__(self, *args, **kwargs):
        import torch
        data = self.waveform_generator(*args, **kwargs)
        if not isinstance(data, (list, tuple)):       # complex h+ - i hx (check_mode_by_mode.py:247)
            data = _as_dev(data, torch.complex128)
            data = [data.real, -data.imag]
        chans = get_fft_td_windowed(data, self.window, self.dt)
        pm = self.positive_frequency_mask
        mask = torch.as_tensor(np.asarray(pm.cpu() if hasattr(pm, "cpu") else pm), device=chans[0].device)
        ch1, ch2 = chans[0][mask], chans[1][mask]
        if self.non_zero_mask is not None:
            nz = self.non_zero_mask
            nz = torch.as_tensor(np.asarray(nz.cpu() if hasattr(nz, "cpu") else nz), device=ch1.device)
            ch1 = torch.where(nz, ch1, torch.zeros_like(ch1))
            ch2 = torch.where(nz, ch2, torch.zeros_like(ch2))
        return [ch1, ch2]

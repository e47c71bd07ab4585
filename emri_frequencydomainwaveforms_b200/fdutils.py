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
  the circular convolution (a (*) b)/N.  The DFT of a Hann-type window is a narrow band, so the product path is a
  banded stencil kernel (emrifd_window_taps / emrifd_band_convolve: direct summation of the 2H+1 central taps, one
  read + one write of the signal, certified truncation bound); general windows fall back to three cuFFT calls
  (``torch.fft``; the reference's own commented FFT variant FDutils.py:83-85).
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


WINDOW_HMAX = 128          # widest band tried (taps -H..H); EMRIFD_WINDOW_MAX_TAPS = 256 is the C-ABI limit
WINDOW_RTOL = 1e-7         # default truncation bound of the banded path (see get_fd_windowed); 0 = always the exact FFT evaluation
_band_cache = {}


def _fft_convolution(a, b):
    """Exact evaluation as three library FFTs (cuFFT through torch.fft): the fallback for windows whose DFT is not band-limited."""
    import torch
    return torch.fft.ifft(torch.fft.fft(a) * torch.fft.fft(b)) / a.shape[0]


def _choose_band(e_band, e_tot, rtol):
    """Smallest H whose dropped taps hold <= rtol^2 of the DFT's energy (cumulative band energies e_band[H]); None if none does.
    The tail is a difference of O(1) numbers, so bounds below ~3e-8 cannot be certified."""
    if not (rtol >= 3e-8) or not (e_tot > 0.0):
        return None
    tail = np.sqrt(np.maximum(e_tot - e_band, 0.0) / e_tot)
    ok = np.nonzero(tail <= rtol)[0]
    return (int(ok[0]), float(tail[ok[0]])) if len(ok) else None


def window_band(window, window_in_fd=False, rtol=WINDOW_RTOL):
    """Band of the convolution kernel a = conj(fft(window)) (FDutils.py:87-93): returns ``(taps, H, bound)`` with ``taps`` a device
    complex128 tensor of a[i], i = -H..H, and ``bound`` the certified relative L2 weight of the dropped taps -- or ``None`` when no
    band of <= WINDOW_HMAX taps reaches ``rtol`` (then the caller uses the FFT evaluation).  A real time-domain window never goes
    through an FFT here: its central taps are summed directly on the GPU (emrifd_window_taps)."""
    import torch
    from . import _lib
    h = _lib.get_handle()
    key = (id(window), bool(window_in_fd), float(rtol), getattr(window, "_version", None))
    hit = _band_cache.get(key)
    if hit is not None and hit[0] is window:
        return hit[1]
    N = int(window.shape[0])
    Hm = min(WINDOW_HMAX, (N - 1) // 2)
    res = None
    if not window_in_fd:
        w = window if torch.is_tensor(window) else np.asarray(window)
        if (w.dtype.is_complex if torch.is_tensor(w) else np.iscomplexobj(w)):
            res = None                                   # complex time-domain window: general case, FFT evaluation
        else:
            wd = _as_dev(w, torch.float64).contiguous()
            taps = torch.empty(2 * (Hm + 2), dtype=torch.float64, device=h.torch_device)
            h.check(h.lib.emrifd_window_taps(h.h, wd.data_ptr(), N, Hm, taps.data_ptr()))
            th = taps.cpu().numpy()
            W = th[0:2 * (Hm + 1):2] + 1j * th[1:2 * (Hm + 1):2]          # W_0 .. W_Hm
            e_tot = N * th[2 * (Hm + 1)]
            p2 = np.abs(W) ** 2
            e_band = p2[0] + 2.0 * np.concatenate([[0.0], np.cumsum(p2[1:])])
            ch = _choose_band(e_band, e_tot, rtol)
            if ch is not None:
                H, bound = ch
                a = np.concatenate([W[H:0:-1], np.conj(W[:H + 1])])      # a_i = conj(W_i); W_{-i} = conj(W_i) for a real window
                res = (torch.from_numpy(np.ascontiguousarray(a)).to(h.torch_device), H, bound)
    else:
        fw = _as_dev(window, torch.complex128).contiguous()
        en = torch.empty(Hm + 2, dtype=torch.float64, device=h.torch_device)
        h.check(h.lib.emrifd_band_energy(h.h, fw.data_ptr(), N, Hm, en.data_ptr()))
        eh = en.cpu().numpy()
        ch = _choose_band(np.cumsum(eh[:Hm + 1]), float(eh.sum()), rtol)
        if ch is not None:
            H, bound = ch
            idx = torch.arange(-H, H + 1, device=h.torch_device) % N
            res = (torch.conj(fw[idx]).resolve_conj().contiguous(), H, bound)
    if len(_band_cache) > 8:
        _band_cache.clear()
    _band_cache[key] = (window, res)
    return res


def band_convolve(taps, H, signals, out_lo=0, out_n=None):
    """out[c][k - out_lo] = (1/N) sum_{i=-H..H} taps[i + H] signals[c][(k - i) mod N] on the GPU (emrifd_band_convolve): one read and
    one write of the signal.  ``signals``: complex128 [nch, N] device tensor."""
    import torch
    from . import _lib
    h = _lib.get_handle()
    sig = signals.contiguous()
    nch, N = sig.shape
    out_n = N - out_lo if out_n is None else out_n
    out = torch.empty((nch, out_n), dtype=torch.complex128, device=sig.device)
    h.check(h.lib.emrifd_band_convolve(h.h, taps.data_ptr(), int(H), sig.data_ptr(), int(nch), int(N), int(out_lo), int(out_n),
                                       out.data_ptr()))
    return out


def get_convolution(a, b, rtol=WINDOW_RTOL):
    """``convolve(hstack((a[1:], a)), b, mode='valid') / len(b)`` (FDutils.py:35-47) for equal-length 1D arrays:
    out[k] = (1/N) sum_j a[(k - j) mod N] b[j].  When ``a`` is band-limited around index 0 (mod N) -- the conjugated DFT of a window
    -- to within ``rtol`` (relative L2 weight of the dropped entries, measured on the device) it is applied as a banded stencil;
    otherwise (or with ``rtol=0``) as ifft(fft(a) fft(b)) / N."""
    import torch
    from . import _lib
    a = _as_dev(a, torch.complex128)
    b = _as_dev(b, torch.complex128)
    if a.ndim != 1 or a.shape != b.shape:
        raise ValueError("get_convolution needs two 1D arrays of equal length.")
    N = int(a.shape[0])
    Hm = min(WINDOW_HMAX, (N - 1) // 2)
    if rtol and N >= 3:
        h = _lib.get_handle()
        a = a.contiguous()
        en = torch.empty(Hm + 2, dtype=torch.float64, device=h.torch_device)
        h.check(h.lib.emrifd_band_energy(h.h, a.data_ptr(), N, Hm, en.data_ptr()))
        eh = en.cpu().numpy()
        ch = _choose_band(np.cumsum(eh[:Hm + 1]), float(eh.sum()), rtol)
        if ch is not None:
            H = ch[0]
            idx = torch.arange(-H, H + 1, device=a.device) % N
            return band_convolve(a[idx].contiguous(), H, b[None, :])[0]
    return _fft_convolution(a, b)


def get_fd_windowed(signal, window, window_in_fd=False, rtol=WINDOW_RTOL, out_lo=0, out_n=None):
    """Convolve the FD channels [h+, hx] with the DFT of a time-domain window (FDutils.py:66-101).

    The DFT of the windows the reference's scripts use is a narrow band (a symmetric Hann of the 1-yr grid length keeps all but
    3e-8 of its L2 weight within |i| <= 4), so the convolution runs as a banded stencil kernel; the dropped taps change
    out[k] by at most ``rtol`` * sqrt(sum|fft(window)|^2) * ||signal||_2 / N (Cauchy-Schwarz).  Windows that are not band-limited
    to ``rtol`` within WINDOW_HMAX taps (and ``rtol=0``) take the exact FFT evaluation.  ``out_lo``/``out_n`` restrict the output
    to a contiguous index range (the f >= 0 half)."""
    import torch
    if window is None:
        return [signal[0], signal[1]]
    s0, s1 = _as_dev(signal[0], torch.complex128), _as_dev(signal[1], torch.complex128)
    band = window_band(window, window_in_fd, rtol) if rtol else None
    if band is not None and s0.ndim == 1 and s0.shape == s1.shape and s0.shape[0] == window.shape[0]:
        taps, H, _ = band
        out = band_convolve(taps, H, torch.stack([s0, s1]), out_lo, out_n)
        return [out[0], out[1]]
    fft_window = _as_dev(window, torch.complex128)
    if not window_in_fd:
        fft_window = torch.fft.fft(fft_window)
    cw = torch.conj(fft_window)
    hi = None if out_n is None else out_lo + out_n
    return [_fft_convolution(cw, s0)[out_lo:hi], _fft_convolution(cw, s1)[out_lo:hi]]


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
        elif self.window is not None and self._is_upper_half:
            # windowed, f >= 0 wanted: the convolution kernel writes only that half (no boolean gather pass)
            sig = self.waveform_generator(*args, **kwargs)
            n = len(self.positive_frequency_mask)
            ch1, ch2 = get_fd_windowed(sig, self.window, window_in_fd=self.window_in_fd, out_lo=(n - 1) // 2, out_n=(n + 1) // 2)
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

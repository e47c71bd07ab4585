"""``inner_product`` / ``snr`` -- mirror of LISAanalysistools/lisatools/diagnostic.py:14-186 for
frequency-domain signals, computed by ``emrifd_inner_product`` (one fused FP64 reduction on the GPU):

    <a|b> = 4 * sum_channels sum_k dx_k * Re(conj(a_k) b_k) / S_k,
    dx_k = f_k - f_{k-1}, dx_0 = dx_1  (right summation rule, diagnostic.py:95-110).
"""
import numpy as np

from .. import _lib


def _dev(x, handle, dtype):
    import torch
    if torch.is_tensor(x):
        return x.to(device=handle.torch_device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(handle.torch_device)


def inner_product(sig1, sig2, dt=None, df=None, f_arr=None, PSD="lisasens", PSD_args=(), PSD_kwargs={},
                  normalize=False, use_gpu=True, complex=False, device=None):
    import torch
    if df is None and dt is None and f_arr is None:
        raise ValueError("Must provide either df, dt or f_arr keyword arguments.")
    if not isinstance(sig1, list):
        sig1 = [sig1]
    if not isinstance(sig2, list):
        sig2 = [sig2]
    if len(sig1) != len(sig2):
        raise ValueError("Signal 1 has {} channels. Signal 2 has {} channels. Must be equal.".format(
            len(sig1), len(sig2)))
    h = _lib.get_handle(device)
    if dt is not None:
        # time-domain inputs (diagnostic.py:49-67): zero-pad the shorter signal, rfft * dt (cuFFT), drop the DC bin
        import warnings
        ta = [_dev(s, h, torch.float64) for s in sig1]
        tb = [_dev(s, h, torch.float64) for s in sig2]
        length = max(ta[0].shape[0], tb[0].shape[0])
        if ta[0].shape[0] != tb[0].shape[0]:
            warnings.warn("The two signals are two different lengths in the time domain. Zero padding smaller array.")
            ta = [torch.nn.functional.pad(s, (0, length - s.shape[0])) for s in ta]
            tb = [torch.nn.functional.pad(s, (0, length - s.shape[0])) for s in tb]
        sig1 = [torch.fft.rfft(s)[1:] * dt for s in ta]
        sig2 = [torch.fft.rfft(s)[1:] * dt for s in tb]
        f_arr = torch.fft.rfftfreq(length, dt, dtype=torch.float64, device=h.torch_device)[1:]
    a = torch.stack([_dev(s, h, torch.complex128) for s in sig1])
    b = torch.stack([_dev(s, h, torch.complex128) for s in sig2])
    nch, n = a.shape
    if b.shape != a.shape:
        raise ValueError("Length of all channels must match.")
    if df is not None and dt is None:
        freqs = (torch.arange(n, dtype=torch.float64, device=h.torch_device) + 1) * df   # ignores DC (+1)
    else:
        freqs = _dev(f_arr, h, torch.float64)
    if isinstance(PSD, str):
        raise ValueError("String PSD names resolve through lisatools.sensitivity, which is outside this path: "
                         "pass PSD=<array> (e.g. fdutils.get_sensitivity(f)) or PSD=None.")
    elif PSD is None:
        psd = None
    else:
        psd = _dev(PSD, h, torch.float64)
        if psd.numel() != n:
            raise ValueError("PSD array length must match the signals.")
    if freqs.numel() != n:
        raise ValueError("f_arr length must match the signals.")

    def ip(x, y):
        out = torch.empty(2, dtype=torch.float64, device=h.torch_device)
        h.check(h.lib.emrifd_inner_product(h.h, x.data_ptr(), y.data_ptr(), nch, n, freqs.data_ptr(),
                                           _lib.ptr(psd), out.data_ptr()))
        o = out.cpu().numpy()
        return o[0] + 1j * o[1]

    out = ip(a, b)
    if not complex:
        out = out.real
    norm = 1.0
    if normalize is True:
        norm = np.sqrt(ip(a, a).real * ip(b, b).real)
    elif isinstance(normalize, str):
        if normalize == "sig1":
            norm = ip(a, a).real
        elif normalize == "sig2":
            norm = ip(b, b).real
        else:
            raise ValueError("If normalizing with respect to sig1 or sig2, normalize kwarg must either be 'sig1' or 'sig2'.")
    elif normalize is not False:
        raise ValueError("Normalize must be True, False, 'sig1', or 'sig2'.")
    return out / norm


def snr(sig1, *args, data=None, use_gpu=True, **kwargs):
    sig2 = sig1 if data is None else data
    return np.sqrt(inner_product(sig1, sig2, *args, use_gpu=use_gpu, **kwargs))

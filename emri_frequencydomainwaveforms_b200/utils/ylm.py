"""Spin-weight -2 spherical harmonics (mirror of ``few.utils.ylm.GetYlms``).

Reference usage: ``ylm_gen = GetYlms(assume_positive_m=True)``; ``ylm_gen(l, m, theta, phi)``
returns ``[Y_{l m}] ++ [Y_{l,-m}]`` (Tutorial_FD_construction_single_mode.ipynb:87,597-611).
Host-side (tens of values per waveform); SURVEY.md section 8f lists a device version as "next".
"""
from math import factorial, pi

import numpy as np


def _wigner_d(l, mp, m, beta):
    """Wigner small-d d^l_{mp,m}(beta) by the explicit finite sum (exact at beta = 0, pi)."""
    cb, sb = np.cos(beta / 2.0), np.sin(beta / 2.0)
    pref = np.sqrt(float(factorial(l + mp) * factorial(l - mp) * factorial(l + m) * factorial(l - m)))
    tot = 0.0
    for k in range(max(0, m - mp), min(l + m, l - mp) + 1):
        den = factorial(l + m - k) * factorial(k) * factorial(l - k - mp) * factorial(k - m + mp)
        tot += (-1.0) ** (k - m + mp) / den * cb ** (2 * l - 2 * k + m - mp) * sb ** (2 * k - m + mp)
    return pref * tot


def spin_weighted_ylm(s, l, m, theta, phi):
    """sY_lm(theta, phi) = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta) e^{i m phi}."""
    if abs(m) > l or l < abs(s):
        return 0.0j
    return ((-1.0) ** s) * np.sqrt((2 * l + 1) / (4.0 * pi)) * _wigner_d(l, m, -s, theta) * np.exp(1j * m * phi)


class GetYlms:
    def __init__(self, assume_positive_m=False, use_gpu=False, **kwargs):
        self.assume_positive_m = assume_positive_m

    def __call__(self, l_in, m_in, theta, phi):
        l_in = np.asarray(l_in).astype(int)
        m_in = np.asarray(m_in).astype(int)
        if self.assume_positive_m:
            if np.any(m_in < 0):
                raise ValueError("Code is assuming positive m values. The input has negative m.")
            l = np.concatenate([l_in, l_in])
            m = np.concatenate([m_in, -m_in])
        else:
            l, m = l_in, m_in
        return np.array([spin_weighted_ylm(-2, int(a), int(b), theta, phi) for a, b in zip(l, m)],
                        dtype=np.complex128)


def ylm_batch_device(l_arr, m_arr, neg_src, theta, phi, handle, lmax=10, cache=None):
    """Device version for a batch of viewing angles (emrifd_ylm_batch):
    returns torch complex128 [B, M + Mneg] = [Y_{l_i m_i}] ++ [Y_{l,-m} of the m > 0 modes] per walker, i.e.
    ``GetYlms(assume_positive_m=True)`` already expanded to the mode basis.  l_arr, m_arr, neg_src: device int32."""
    import torch
    dev = handle.torch_device
    th = torch.as_tensor(np.atleast_1d(np.asarray(theta, dtype=np.float64))).to(dev)
    ph = torch.as_tensor(np.atleast_1d(np.asarray(phi, dtype=np.float64))).to(dev)
    B, M, Mneg = th.shape[0], l_arr.shape[0], neg_src.shape[0]
    out = torch.empty((B, M + Mneg), dtype=torch.complex128, device=dev)
    handle.check(handle.lib.emrifd_ylm_batch(handle.h, l_arr.data_ptr(), m_arr.data_ptr(), M, neg_src.data_ptr(), Mneg,
                                             th.data_ptr(), ph.data_ptr(), B, int(lmax), out.data_ptr()))
    return out

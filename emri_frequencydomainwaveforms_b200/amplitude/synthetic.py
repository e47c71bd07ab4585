"""Synthetic Teukolsky-amplitude producer with FEW's mode-index layout.

The reference evaluates ``few.amplitude.romannet.RomanAmplitude`` (ROMAN network; weights are a
Zenodo download, absent offline).  This stand-in keeps what the hot path depends
on: the (l, m, n) layout -- l = 2..10, m = 0..l, n = -30..30, 3843 modes -- a ``[L, 3843]``
complex128 return, smooth dependence on (p, e) and >= 6 decades of dynamic range across modes.
It is an input producer, not part of the accelerated path.
"""
import numpy as np

LMAX, NMAX = 10, 30


def mode_index_arrays(lmax=LMAX, nmax=NMAX):
    l_arr, m_arr, n_arr = [], [], []
    for l in range(2, lmax + 1):
        for m in range(0, l + 1):
            for n in range(-nmax, nmax + 1):
                l_arr.append(l), m_arr.append(m), n_arr.append(n)
    return (np.asarray(l_arr, dtype=np.int32), np.asarray(m_arr, dtype=np.int32),
            np.asarray(n_arr, dtype=np.int32))


class SyntheticAmplitude:
    """``amp(p, e)`` -> ``teuk_modes[L, num_modes]`` complex128 (call shape of RomanAmplitude)."""

    def __init__(self, lmax=LMAX, nmax=NMAX, **kwargs):
        self.lmax, self.nmax = lmax, nmax
        self.l_arr, self.m_arr, self.n_arr = mode_index_arrays(lmax, nmax)
        self.num_teuk_modes = len(self.l_arr)
        lm = self.l_arr.astype(np.int64) * 100 + self.m_arr
        _, first, inverse = np.unique(lm, return_index=True, return_inverse=True)
        self.unique_l = self.l_arr[first]
        self.unique_m = self.m_arr[first]
        self.inverse_lm = inverse

    def _tables(self):
        if not hasattr(self, "_cmode"):
            l, m, n = self.l_arr.astype(np.float64), self.m_arr.astype(np.float64), self.n_arr.astype(np.float64)
            self._cmode = (0.6 ** (l - 2.0)) / (1.0 + (l - m)) * np.exp(1j * (0.4 * l - 0.25 * m + 0.15 * n))
            self._lidx = (self.l_arr - 2).astype(np.intp)
            self._midx = self.m_arr.astype(np.intp)
            self._nidx = (self.n_arr + self.nmax).astype(np.intp)
        return self._cmode, self._lidx, self._midx, self._nidx

    def __call__(self, p, e, *args, specific_modes=None, **kwargs):
        """A_lmn(p, e) = c_lm p^{-l/2} exp(-(n - n0)^2 / 2 sigma^2) exp(i (0.4 l - 0.25 m + 0.15 n + 8/p + e n / 6)),
        n0 = 2.5 e (1 + 0.2 m)/sqrt(1 - e), sigma = 0.35 + 3.5 e -- evaluated through small separable tables
        ([L, l], [L, m, n], [L, n]) that are gathered per mode."""
        p = np.atleast_1d(np.asarray(p, dtype=np.float64))
        e = np.atleast_1d(np.asarray(e, dtype=np.float64))
        cmode, lidx, midx, nidx = self._tables()
        lv = np.arange(2, self.lmax + 1, dtype=np.float64)
        mv = np.arange(0, self.lmax + 1, dtype=np.float64)
        nv = np.arange(-self.nmax, self.nmax + 1, dtype=np.float64)
        P = p[:, None] ** (-lv[None, :] / 2.0)                                           # [L, l]
        n0 = (2.5 * e / np.sqrt(1.0 - e))[:, None] * (1.0 + 0.2 * mv)[None, :]           # [L, m]
        sig = 0.35 + 3.5 * e
        ENV = np.exp(-((nv[None, None, :] - n0[:, :, None]) ** 2) / (2.0 * sig * sig)[:, None, None])   # [L, m, n]
        EPH = np.exp(1j * ((e / 6.0)[:, None] * nv[None, :] + (8.0 / p)[:, None]))      # [L, n]
        out = P[:, lidx] * ENV[:, midx, nidx] * (EPH[:, nidx] * cmode[None, :])
        return self._finish(out, specific_modes)

    def device_call(self, p, e, device, handle=None):
        """Same amplitudes evaluated on the GPU by ``emrifd_synth_amplitude`` (torch complex128 [L, num_modes]); values agree
        with the host version to rounding (different pow/exp/sincos implementations).  ``p``, ``e``: numpy or torch."""
        import torch
        from .. import _lib
        h = handle or _lib.get_handle(device.index if hasattr(device, "index") and device.index is not None else None)
        dev = h.torch_device
        cmode, lidx, midx, nidx = self._tables()
        if getattr(self, "_dev_tables", None) is None or self._dev_tables[0].device != dev:
            up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
            self._dev_tables = (torch.as_tensor(np.ascontiguousarray(cmode)).to(dev), up(self.l_arr), up(self.m_arr), up(self.n_arr))
        cm, la, ma, na = self._dev_tables
        if not torch.is_tensor(p):
            p = torch.as_tensor(np.asarray(p, dtype=np.float64))
        if not torch.is_tensor(e):
            e = torch.as_tensor(np.asarray(e, dtype=np.float64))
        p = p.to(device=dev, dtype=torch.float64).contiguous()
        e = e.to(device=dev, dtype=torch.float64).contiguous()
        out = torch.empty((p.shape[0], self.num_teuk_modes), dtype=torch.complex128, device=dev)
        h.check(h.lib.emrifd_synth_amplitude(h.h, p.data_ptr(), e.data_ptr(), p.shape[0], la.data_ptr(), ma.data_ptr(), na.data_ptr(),
                                             cm.data_ptr(), self.num_teuk_modes, int(self.lmax), int(self.nmax), out.data_ptr()))
        return out

    def _finish(self, out, specific_modes):
        if specific_modes is not None:
            res = {}
            for (ll, mm, nn) in specific_modes:
                idx = np.where((self.l_arr == ll) & (self.m_arr == abs(mm)) & (self.n_arr == nn))[0][0]
                res[(ll, mm, nn)] = out[:, idx].copy()
            return res
        return out

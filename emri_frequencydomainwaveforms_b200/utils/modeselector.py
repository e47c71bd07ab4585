"""Mode selection by power (mirror of ``few.utils.modeselector.ModeSelector``).

Host-side; sits immediately before the hot path (SURVEY.md section 8f rank 1 lists a device version
as "next").  Semantics: power = |[A, conj(A[:, m>0])] * ylms|^2 per time sample, sort descending,
cumulative sum, keep while cumsum < (1 - eps) * total (first always kept), union over time samples,
fold -m picks back onto their +m mode so +-m stay together.
"""
import numpy as np


class ModeSelector:
    def __init__(self, m0mask, use_gpu=False, **kwargs):
        self.m0mask = np.asarray(m0mask, dtype=bool)       # True where m != 0
        self.num_m_zero_up = len(self.m0mask)
        self.num_m_1_up = int(self.m0mask.sum())
        self.num_m0 = self.num_m_zero_up - self.num_m_1_up

    def __call__(self, teuk_modes, ylms, modeinds, eps=1e-5):
        """teuk_modes [L, M]; ylms [M + M_{m>0}] ( +m block, then -m block for m>0 modes );
        modeinds = [l_arr, m_arr, n_arr].  Returns (teuk_modes_kept, ylms_kept [2K], ls, ms, ns)."""
        zero_up = self.num_m_zero_up
        full = np.concatenate([teuk_modes, np.conj(teuk_modes[:, self.m0mask])], axis=1)
        # |A Y|^2 with every operation individually rounded (the device kernel does exactly the same arithmetic)
        ar, ai, yr, yi = full.real, full.imag, ylms.real[None, :], ylms.imag[None, :]
        re, im = ar * yr - ai * yi, ar * yi + ai * yr
        power = re * re + im * im
        picked = None
        ntot = power.shape[1]
        ktop = 768
        if ntot > 4 * ktop:
            # only the strongest few hundred harmonics can be picked: sort those, and check that they reach the threshold
            part = np.argpartition(power, ntot - ktop, axis=1)[:, ntot - ktop:]
            psub = np.take_along_axis(power, part, axis=1)
            order = np.argsort(psub, axis=1)[:, ::-1]
            inds_top = np.take_along_axis(part, order, axis=1)
            cs = np.cumsum(np.take_along_axis(psub, order, axis=1), axis=1)
            thresh = power.sum(axis=1)[:, None] * (1.0 - eps)
            if np.all(cs[:, -1] >= thresh):
                keep_sorted = np.ones_like(cs, dtype=bool)
                keep_sorted[:, 1:] = cs[:, :-1] < thresh
                picked = np.unique(inds_top[keep_sorted])
        if picked is None:
            inds_sort = np.argsort(power, axis=1)[:, ::-1]
            power_sorted = np.take_along_axis(power, inds_sort, axis=1)
            cumsum = np.cumsum(power_sorted, axis=1)
            thresh = cumsum[:, -1][:, None] * (1.0 - eps)
            keep_sorted = np.ones_like(cumsum, dtype=bool)
            keep_sorted[:, 1:] = cumsum[:, :-1] < thresh
            picked = np.unique(inds_sort[keep_sorted])
        # fold -m picks onto their +m partner
        m_nonzero_idx = np.where(self.m0mask)[0]
        neg = picked >= zero_up
        keep_modes = np.unique(np.concatenate([picked[~neg], m_nonzero_idx[picked[neg] - zero_up]]))
        # position of each kept mode inside the -m block (m = 0 modes reuse their +m ylm)
        pos_in_neg = np.cumsum(self.m0mask) - 1
        neg_idx = np.where(self.m0mask[keep_modes], zero_up + pos_in_neg[keep_modes], keep_modes)
        ylms_kept = np.concatenate([ylms[keep_modes], ylms[neg_idx]])
        l_arr, m_arr, n_arr = modeinds
        return (teuk_modes[:, keep_modes], ylms_kept, l_arr[keep_modes], m_arr[keep_modes], n_arr[keep_modes])

    # ---- device version (emrifd_mode_select) -----------------------------------------
    def select_device(self, teuk_dev, samp_walker, ylm_full, eps=1e-5, handle=None):
        """teuk_dev: torch complex128 [nsamp, M] (all walkers' time samples back to back) on the GPU;
        samp_walker: int32 [nsamp]; ylm_full: complex128 [B, M + M_{m>0}].  Returns a torch uint8 [B, M] keep mask
        (union over each walker's samples, -m picks folded onto their +m partner)."""
        import torch
        from .. import _lib
        h = handle or _lib.get_handle()
        dev = h.torch_device
        if not hasattr(self, "_neg_src_dev") or self._neg_src_dev.device != dev:
            self._neg_src_dev = torch.as_tensor(np.where(self.m0mask)[0].astype(np.int32)).to(dev)
        teuk_dev = teuk_dev.to(device=dev, dtype=torch.complex128).contiguous()
        sw = torch.as_tensor(np.asarray(samp_walker, dtype=np.int32)).to(dev)
        yl = torch.as_tensor(np.ascontiguousarray(ylm_full, dtype=np.complex128)).to(dev)
        B, nsamp, M = yl.shape[0], teuk_dev.shape[0], teuk_dev.shape[1]
        if M != self.num_m_zero_up or yl.shape[1] != M + self.num_m_1_up:
            raise ValueError("teuk_modes / ylms do not match the mode basis of this selector")
        flags = torch.empty((B, M), dtype=torch.uint8, device=dev)
        h.check(h.lib.emrifd_mode_select(h.h, teuk_dev.data_ptr(), nsamp, M, sw.data_ptr(), yl.data_ptr(),
                                         self._neg_src_dev.data_ptr(), self.num_m_1_up, B, float(eps), flags.data_ptr()))
        return flags

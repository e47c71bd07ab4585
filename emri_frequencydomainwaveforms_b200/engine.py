"""Host-side packing + launch logic above the C-ABI (include/emrifd.h).

A *walker* is one waveform evaluation: its sparse trajectory (t, Phi_phi, Phi_r, f_phi, f_r), the
kept Teukolsky modes ``teuk_modes[L, K]``, their (m, n) and ``ylms[2K]``.  Walkers of a batch are
ragged (L and K differ); they are packed back to back and described by ``emrifd_walker_t`` records.
PyTorch is used only to own device buffers and streams.
"""
import numpy as np

from . import _lib
from ._lib import WALKER_DTYPE, BRANCH_DTYPE, MAX_BRANCHES, INCLUDE_MINUS_M, MASK_POSITIVE


class PackedBatch:
    """Packed host arrays of a ragged walker batch (layout documented in include/emrifd.h)."""

    def __init__(self, items):
        B = len(items)
        if B == 0:
            raise ValueError("empty batch")
        w = np.zeros(B, dtype=WALKER_DTYPE)
        ko = to = mo = co = 0
        for i, it in enumerate(items):
            L, K = it["teuk_modes"].shape
            if len(it["t"]) != L or len(it["m_arr"]) != K or len(it["n_arr"]) != K or len(it["ylms"]) != 2 * K:
                raise ValueError("inconsistent walker shapes: teuk_modes [L,K], t [L], m/n [K], ylms [2K]")
            w[i] = (L, K, ko, to, mo, co, 0, it.get("scale", 1.0), it.get("cos2psi", 1.0), it.get("sin2psi", 0.0))
            ko += L
            to += L * K
            mo += K
            co += L * (2 * K + 4) * 4
        self.B, self.walkers = B, w
        self.n_knots, self.n_teuk, self.n_modes, self.n_coeff = ko, to, mo, co
        cat = lambda key, dt: np.ascontiguousarray(np.concatenate([np.asarray(it[key]).ravel() for it in items]), dtype=dt)
        self.t = cat("t", np.float64)
        self.f_phi = cat("f_phi", np.float64)
        self.f_r = cat("f_r", np.float64)
        self.Phi_phi = cat("Phi_phi", np.float64)
        self.Phi_r = cat("Phi_r", np.float64)
        self.teuk = cat("teuk_modes", np.complex128)
        self.m = cat("m_arr", np.int32)
        self.n = cat("n_arr", np.int32)
        self.ylm = cat("ylms", np.complex128)
        self.Lmax = int(w["L"].max())
        self.Kmax = int(w["K"].max())

    def h2d_bytes(self):
        return int(self.walkers.nbytes + 5 * self.t.nbytes + self.teuk.nbytes + self.m.nbytes + self.n.nbytes
                   + self.ylm.nbytes)


class DeviceBatch:
    """Device-resident copy of a PackedBatch plus its work buffers (coeff, branches)."""

    def __init__(self, pb, handle):
        import torch
        dev = handle.torch_device
        self.pb, self.handle = pb, handle
        up = lambda a: torch.from_numpy(a.view(np.float64) if a.dtype == np.complex128 else a).to(dev)
        self.t, self.f_phi, self.f_r = up(pb.t), up(pb.f_phi), up(pb.f_r)
        self.Phi_phi, self.Phi_r = up(pb.Phi_phi), up(pb.Phi_r)
        self.teuk, self.m, self.n, self.ylm = up(pb.teuk), up(pb.m), up(pb.n), up(pb.ylm)
        self.coeff = torch.empty(pb.n_coeff, dtype=torch.float64, device=dev)
        self.branches = torch.zeros(pb.n_modes * MAX_BRANCHES * BRANCH_DTYPE.itemsize, dtype=torch.uint8, device=dev)

    @classmethod
    def from_device_parts(cls, handle, walkers, tracks, teuk, m, n, ylm):
        """Batch whose packed arrays were produced on the device (device Ylm / mode selection / compaction,
        waveform.FastSchwarzschildEccentricFlux.prepare_batch_device).  ``walkers``: host WALKER_DTYPE records;
        ``tracks``: dict of device float64 tensors t, f_phi, f_r, Phi_phi, Phi_r [sum L]."""
        import torch
        self = cls.__new__(cls)
        pb = PackedBatch.__new__(PackedBatch)
        pb.B, pb.walkers = len(walkers), walkers
        pb.n_knots = int(walkers["L"].sum())
        pb.n_teuk = int((walkers["L"].astype(np.int64) * walkers["K"]).sum())
        pb.n_modes = int(walkers["K"].sum())
        pb.n_coeff = int((walkers["L"].astype(np.int64) * (2 * walkers["K"].astype(np.int64) + 4) * 4).sum())
        pb.Lmax, pb.Kmax = int(walkers["L"].max()), int(walkers["K"].max())
        self.pb, self.handle = pb, handle
        self.t, self.f_phi, self.f_r = tracks["t"], tracks["f_phi"], tracks["f_r"]
        self.Phi_phi, self.Phi_r = tracks["Phi_phi"], tracks["Phi_r"]
        self.teuk, self.m, self.n, self.ylm = teuk, m, n, ylm
        dev = handle.torch_device
        self.coeff = torch.empty(pb.n_coeff, dtype=torch.float64, device=dev)
        self.branches = torch.zeros(pb.n_modes * MAX_BRANCHES * BRANCH_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        return self

    def branches_host(self):
        """Work-list (A4 output) as a structured numpy array [n_modes, MAX_BRANCHES]."""
        return self.branches.cpu().numpy().view(BRANCH_DTYPE).reshape(self.pb.n_modes, MAX_BRANCHES)

    def coeff_host(self, i=0):
        w = self.pb.walkers[i]
        L, R = int(w["L"]), 2 * int(w["K"]) + 4
        o = int(w["coeff_off"])
        return self.coeff[o:o + L * R * 4].cpu().numpy().reshape(L, R, 4)


def group_evaluations(db):
    """Stationary points the mode sum actually solves: one per ((m, n) group, bin), i.e. the work-list records of the first
    mode of every distinct (m, n) pair of a walker (csrc/emrifd.cu group_kernel).  Returns an int64 array [B]."""
    br = db.branches_host()
    m, n = db.m.cpu().numpy(), db.n.cpu().numpy()
    per_mode = np.where(br["end"] >= br["start"], br["end"] - br["start"] + 1, 0).sum(axis=1)
    out = np.zeros(db.pb.B, dtype=np.int64)
    for i, w in enumerate(db.pb.walkers):
        o, K = int(w["mode_off"]), int(w["K"])
        _, first = np.unique(np.stack([m[o:o + K], n[o:o + K]], axis=1), axis=0, return_index=True)
        out[i] = per_mode[o:o + K][first].sum()
    return out


def grid_from_frequency(frequency):
    """Validate a two-sided frequency array: odd length, symmetric, one zero in the middle
    (FDInterpolatedModeSum asserts exactly one zero; emri_pe.py:339-342 builds f_arr this way).
    Returns (N, fpos) with fpos the f >= 0 half."""
    f = np.asarray(frequency, dtype=np.float64)
    N = len(f)
    if N == 0:
        raise ValueError("Input f_arr has zero length.")
    if N % 2 == 0 or N < 3:
        raise ValueError("f_arr must have odd length >= 3 with f = 0 in the middle.")
    zero = (N - 1) // 2
    if f[zero] != 0.0 or np.sum(f == 0.0) != 1:
        raise ValueError("f_arr must contain exactly one zero, at its centre.")
    fpos = np.ascontiguousarray(f[zero:])
    if not np.array_equal(f[:zero], -fpos[:0:-1]):
        raise ValueError("f_arr must be symmetric about zero.")
    if not np.all(np.diff(fpos) > 0):
        raise ValueError("f_arr must be strictly increasing.")
    return N, fpos


def run_waveform(db, N, val=0.0, fpos_dev=None, include_minus_m=True, mask_positive=False, like=False):
    """A3 -> A4 -> A5..A7 for a device batch.  Returns (hp, hc [B, n_out] complex128 torch, like [B,3] or None)."""
    import torch
    h, pb = db.handle, db.pb
    n_out = (N + 1) // 2 if mask_positive else N
    # one [B, 2, n_out] buffer: walker w writes h+ at w*2*n_out and hx n_out further (= vstack((h+, hx)) per walker)
    pb.walkers["out_off"] = np.arange(pb.B, dtype=np.int64) * (2 * n_out)
    out = torch.empty((pb.B, 2, n_out), dtype=torch.complex128, device=h.torch_device)
    hp, hc = out[:, 0, :], out[:, 1, :]
    like_out = torch.empty((pb.B, 3), dtype=torch.float64, device=h.torch_device) if like else None
    flags = (INCLUDE_MINUS_M if include_minus_m else 0) | (MASK_POSITIVE if mask_positive else 0)
    rc = h.lib.emrifd_fd_waveform_batch(
        h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.teuk.data_ptr(), db.f_phi.data_ptr(),
        db.f_r.data_ptr(), db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(), db.m.data_ptr(), db.n.data_ptr(),
        db.ylm.data_ptr(), int(N), float(val), _lib.ptr(fpos_dev), flags, db.coeff.data_ptr(),
        db.branches.data_ptr(), out.data_ptr(), out.data_ptr() + 16 * n_out, _lib.ptr(like_out))
    h.check(rc)
    db.last_out = out   # [B, 2, n_out]: row w is vstack((h+, hx)) of walker w
    return hp, hc, like_out


def run_loglike(db, N, val=0.0, fpos_dev=None, include_minus_m=True):
    """Fused template + likelihood for a device batch: no h(f) is written to HBM.  Returns [B,3] torch."""
    import torch
    h, pb = db.handle, db.pb
    like_out = torch.empty((pb.B, 3), dtype=torch.float64, device=h.torch_device)
    flags = (INCLUDE_MINUS_M if include_minus_m else 0) | MASK_POSITIVE
    rc = h.lib.emrifd_fd_waveform_batch(
        h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.teuk.data_ptr(), db.f_phi.data_ptr(),
        db.f_r.data_ptr(), db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(), db.m.data_ptr(), db.n.data_ptr(),
        db.ylm.data_ptr(), int(N), float(val), _lib.ptr(fpos_dev), flags, db.coeff.data_ptr(),
        db.branches.data_ptr(), None, None, like_out.data_ptr())
    h.check(rc)
    return like_out


def run_loglike_host(pb, handle, N, val=0.0, fpos_dev=None, include_minus_m=True, hp_dev=None, hc_dev=None):
    """The e2e call: HOST packed inputs in, HOST ll[B,3] out (H2D + kernels + D2H inside).
    hp_dev / hc_dev: optional device tensors [B, (N+1)/2] complex128 that receive the waveforms."""
    out = np.zeros((pb.B, 3))
    if hp_dev is not None:
        pb.walkers["out_off"] = np.arange(pb.B, dtype=np.int64) * ((N + 1) // 2)
    flags = (INCLUDE_MINUS_M if include_minus_m else 0) | MASK_POSITIVE
    rc = handle.lib.emrifd_loglike_batch_host(
        handle.h, pb.walkers.ctypes.data, pb.B, pb.t.ctypes.data, pb.teuk.ctypes.data, pb.f_phi.ctypes.data,
        pb.f_r.ctypes.data, pb.Phi_phi.ctypes.data, pb.Phi_r.ctypes.data, pb.m.ctypes.data, pb.n.ctypes.data,
        pb.ylm.ctypes.data, int(N), float(val), _lib.ptr(fpos_dev), flags, _lib.ptr(hp_dev), _lib.ptr(hc_dev),
        out.ctypes.data)
    handle.check(rc)
    return out


def walker_cost_estimate(item, df):
    """Host-side estimate of a walker's mode-sum work = stationary points to solve: for every distinct (m, n) of its modes the
    frequency bins swept by f_mn = m f_phi + n f_r along the trajectory (total variation, so both branches of a turnover count),
    in bins of width ``df``.  Used to balance walker shards (distributed.balanced_walker_assignment); the exact count comes from the
    device (group_evaluations) once the work-list exists."""
    fp, fr = np.asarray(item["f_phi"]), np.asarray(item["f_r"])
    cost = 0.0
    for m, n in set(zip(np.asarray(item["m_arr"]).tolist(), np.asarray(item["n_arr"]).tolist())):
        cost += float(np.sum(np.abs(np.diff(m * fp + n * fr))))
    return cost / df

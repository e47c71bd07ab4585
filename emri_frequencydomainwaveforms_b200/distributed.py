"""Multi-GPU sharding of the FD hot path on one 8xB200 box (one process per GPU, torch.distributed).

Two natural partitions, neither needs a collective in the data path:

* **walker sharding** (MCMC ensembles, parameter batches; emri_pe.py-style likelihood calls):
  likelihood evaluations are independent (likelihood.py:245-248), so ``params[nb, :]`` is split into
  contiguous blocks, every rank evaluates its block with the fused kernels against its own replica of
  the whitened data, and the ``nb`` log-likelihoods are gathered (``all_gather`` of a few doubles).
* **frequency-bin sharding** (one long, high-mode-count waveform): bins are independent once a thread
  owns the (+f, -f) pair, so the f >= 0 bins are cut into contiguous slices balanced by the number of
  stationary-point evaluations (not by bin count: the signal occupies a few % of the band).  Spline
  coefficients and the work-list are replicated (MBs).  The only exchange is one NCCL
  ``all_reduce(SUM)`` of three doubles per walker: sum|d-h|^2, <d|h>, <h|h>.

The helpers take the compute step as a callable so that the host logic is testable on CPU with the
``gloo`` backend (tests/test_distributed_cpu.py).
"""
import numpy as np


def shard_range(n, world, rank):
    """Contiguous block [lo, hi) of ``n`` items for ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bin_work_histogram(branches, m_arr, N):
    """Evaluations per positive bin j (both the +f and -f full-grid indices of the pair count),
    from the work-list: a difference array over [start, end] of every branch."""
    zero = (N - 1) // 2
    npos = zero + 1
    diff = np.zeros(npos + 1, dtype=np.int64)
    b = branches.reshape(-1)
    ok = b["end"] >= b["start"]
    s = np.maximum(b["start"][ok], 0)
    e = np.minimum(b["end"][ok], N - 1)
    pos = e >= zero                        # +f part: j in [max(s, zero) - zero, e - zero]
    np.add.at(diff, np.maximum(s[pos], zero) - zero, 1)
    np.add.at(diff, e[pos] - zero + 1, -1)
    neg = s < zero                         # -f part: j in [zero - min(e, zero - 1), zero - s]
    np.add.at(diff, zero - np.minimum(e[neg], zero - 1), 1)
    np.add.at(diff, zero - s[neg] + 1, -1)
    return np.cumsum(diff[:-1])


def balanced_bin_slices(work, world, base_cost=0.02, align=1024):
    """Cut [0, len(work)) into ``world`` contiguous slices of equal cost; every bin costs ``base_cost``
    (store / data read) plus its evaluations.  Slice starts are multiples of ``align`` (pass the mode-sum kernel's tile
    size ``emrifd_tile_bins()``: tiles then coincide with the precomputed per-tile sum|d~|^2 table, so empty tiles stay on the fast path).
    Returns [(j_lo, j_cnt)] * world."""
    cost = np.asarray(work, dtype=np.float64) + base_cost
    c = np.concatenate([[0.0], np.cumsum(cost)])
    edges = [0]
    for r in range(1, world):
        e = int(np.searchsorted(c, c[-1] * r / world))
        if align > 1:
            e = int(round(e / align)) * align
        edges.append(min(e, len(cost)))
    edges.append(len(cost))
    edges = np.maximum.accumulate(edges)
    return [(int(edges[r]), int(edges[r + 1] - edges[r])) for r in range(world)]


def balanced_walker_assignment(cost, world):
    """Deal walkers to ``world`` ranks so that every rank carries about the same estimated work AND the same number of walkers
    (counts differ by at most one): walkers in order of descending cost (ties by index), each to the rank with the least work so
    far among those that still have room (longest-processing-time rule with a capacity).  Returns a list of index lists, one per
    rank; deterministic, so every rank computes the same assignment without communication."""
    cost = np.asarray(cost, dtype=np.float64)
    n = len(cost)
    cap = [n // world + (1 if r < n % world else 0) for r in range(world)]
    order = sorted(range(n), key=lambda i: (-cost[i], i))
    shares = [[] for _ in range(world)]
    load = [0.0] * world
    for i in order:
        r = min((r for r in range(world) if len(shares[r]) < cap[r]), key=lambda r: (load[r], r))
        shares[r].append(i)
        load[r] += cost[i]
    return shares


def gather_walker_results(local, counts, group=None):
    """all_gather ragged per-rank result vectors (torch tensors, same dtype/device) into one tensor."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: counts[r]] for r in range(world)], dim=0)


def walker_sharded_loglike(params, compute_ll, group=None, device="cpu"):
    """Evaluate ``compute_ll(params_block) -> ll[block]`` on this rank's contiguous block of walkers and
    gather all ``nb`` values on every rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nb = params.shape[0]
    lo, hi = shard_range(nb, world, rank)
    local = np.asarray(compute_ll(params[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    counts = [shard_range(nb, world, r)[1] - shard_range(nb, world, r)[0] for r in range(world)]
    out = gather_walker_results(torch.as_tensor(local, dtype=torch.float64, device=device), counts, group)
    return out.cpu().numpy()


def bin_sharded_sums(compute_partial, slices, group=None, device="cpu"):
    """``compute_partial(j_lo, j_cnt) -> [B, 3]`` partial (ll, <d|h>, <h|h>) sums of this rank's bin slice;
    one all_reduce(SUM) makes them global.  The finalised sums are linear in the per-bin terms, so reducing
    them is exact up to FP64 summation order."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    j_lo, j_cnt = slices[rank]
    part = compute_partial(j_lo, j_cnt)
    t = part if torch.is_tensor(part) else torch.as_tensor(np.asarray(part), dtype=torch.float64, device=device)
    if j_cnt == 0:
        t = torch.zeros_like(t)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


# ---- GPU conveniences ---------------------------------------------------------------------------
def cyclic_tile_owner(world, rank):
    """(tile_first, tile_stride) of ``rank`` in the cyclic tile sharding: rank r owns tiles r, r + world, ..."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    return rank, world


def gpu_bin_sharded_loglike_cyclic(db, N, val=0.0, fpos_dev=None, include_minus_m=True, group=None):
    """Frequency-bin sharded likelihood with CYCLIC tile ownership (emrifd_batch_sum_cyclic): no work histogram, no
    device->host copy of the work-list; every rank builds the (replicated, MB-sized) splines and work-list, sums its
    interleaved tiles, and one NCCL all_reduce(SUM) of [B, 3] doubles finishes."""
    import torch
    import torch.distributed as dist
    from . import _lib
    from ._lib import INCLUDE_MINUS_M, MASK_POSITIVE
    h, pb = db.handle, db.pb
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    flags = (INCLUDE_MINUS_M if include_minus_m else 0) | MASK_POSITIVE
    h.check(h.lib.emrifd_batch_spline(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.teuk.data_ptr(),
                                      db.f_phi.data_ptr(), db.f_r.data_ptr(), db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(),
                                      db.coeff.data_ptr()))
    h.check(h.lib.emrifd_batch_segment(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(),
                                       db.m.data_ptr(), db.n.data_ptr(), int(N), float(val), _lib.ptr(fpos_dev),
                                       db.branches.data_ptr(), None))
    first, stride = cyclic_tile_owner(world, rank)
    out = torch.zeros((pb.B, 3), dtype=torch.float64, device=h.torch_device)
    h.check(h.lib.emrifd_batch_sum_cyclic(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(),
                                          db.m.data_ptr(), db.n.data_ptr(), db.ylm.data_ptr(), db.branches.data_ptr(),
                                          int(N), float(val), _lib.ptr(fpos_dev), flags, int(first), int(stride), None, None,
                                          out.data_ptr()))
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def gpu_bin_sharded_loglike(db, N, val=0.0, fpos_dev=None, include_minus_m=True, group=None, slices=None):
    """Frequency-bin sharded likelihood of a DeviceBatch replicated on every rank (NCCL all_reduce of [B,3]).
    ``slices``: reuse a previous work-balanced partition (e.g. across MCMC steps, where the work distribution
    barely moves) instead of rebuilding it from the work-list (a device->host copy + host histogram)."""
    import torch
    import torch.distributed as dist
    from . import _lib
    from ._lib import INCLUDE_MINUS_M, MASK_POSITIVE
    h, pb = db.handle, db.pb
    world = dist.get_world_size(group)
    flags = (INCLUDE_MINUS_M if include_minus_m else 0) | MASK_POSITIVE
    h.check(h.lib.emrifd_batch_spline(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.teuk.data_ptr(),
                                      db.f_phi.data_ptr(), db.f_r.data_ptr(), db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(),
                                      db.coeff.data_ptr()))
    h.check(h.lib.emrifd_batch_segment(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(),
                                       db.m.data_ptr(), db.n.data_ptr(), int(N), float(val), _lib.ptr(fpos_dev),
                                       db.branches.data_ptr(), None))
    if slices is None:
        work = bin_work_histogram(db.branches_host(), pb.m, N)
        slices = balanced_bin_slices(work, world, align=int(h.lib.emrifd_tile_bins()))

    def partial(j_lo, j_cnt):
        out = torch.zeros((pb.B, 3), dtype=torch.float64, device=h.torch_device)
        if j_cnt > 0:
            h.check(h.lib.emrifd_batch_sum(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(),
                                           db.m.data_ptr(), db.n.data_ptr(), db.ylm.data_ptr(), db.branches.data_ptr(),
                                           int(N), float(val), _lib.ptr(fpos_dev), flags, int(j_lo), int(j_cnt), None, None,
                                           out.data_ptr()))
        return out

    return bin_sharded_sums(partial, slices, group=group), slices

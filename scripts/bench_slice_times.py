"""Diagnostic for the frequency-bin sharded path (configs[3]): time the sliced mode sum of EVERY rank's slice on one GPU.
The 8-GPU likelihood time is max over these (+ spline/segment + all_reduce), so this shows the partition's real balance.
  python scripts/bench_slice_times.py [--world 8] [--T 4.0]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--T", type=float, default=4.0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from emri_frequencydomainwaveforms_b200 import _lib, engine, distributed as D
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    from emri_frequencydomainwaveforms_b200.utils.constants import YRSID_SI
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
    h = _lib.get_handle(0)
    dev = h.torch_device
    gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
    M, mu, e0, dt = 1e6, 10.0, 0.7, 10.0
    p0 = get_p_at_t(gen.inspiral_generator, args.T * 0.99, [M, mu, 0.0, e0, 1.0], xtol=1e-9, bounds=[7.2 + 2 * e0 + 0.05, 16.0 + 2 * e0])
    it = gen.prepare(M, mu, p0, e0, 1.0, -np.pi / 2, dist=1.0, T=args.T, dt=dt, eps=1e-5, mode_selection="all")
    n_pts = int(args.T * YRSID_SI / dt) + 1
    N = n_pts + 1 if n_pts % 2 == 0 else n_pts
    n = (N + 1) // 2
    val = 1.0 / (N * dt)
    db = engine.DeviceBatch(engine.PackedBatch([it]), h)
    hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)
    f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
    wf1 = torch.sqrt(torch.full((n,), val, dtype=torch.float64, device=dev) / get_sensitivity(f_pos))
    wfac = torch.stack([wf1, wf1]).contiguous()
    data_w = (torch.cat([hp, hc], dim=0) * wfac).contiguous()
    h.check(h.lib.emrifd_set_data(h.h, data_w.data_ptr(), wfac.data_ptr(), n))
    pb = db.pb
    work = D.bin_work_histogram(db.branches_host(), pb.m, N)
    slices = D.balanced_bin_slices(work, args.world)
    flags = _lib.INCLUDE_MINUS_M | _lib.MASK_POSITIVE
    out = torch.zeros((1, 3), dtype=torch.float64, device=dev)
    res = []
    for (j_lo, j_cnt) in slices:
        def run():
            h.check(h.lib.emrifd_batch_sum(h.h, pb.walkers.ctypes.data, 1, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(), db.n.data_ptr(),
                                           db.ylm.data_ptr(), db.branches.data_ptr(), N, val, None, flags, int(j_lo), int(j_cnt), None, None,
                                           out.data_ptr()))
        run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(args.reps):
            run()
        b.record()
        torch.cuda.synchronize()
        res.append({"j_lo": int(j_lo), "bins": int(j_cnt), "tiles": int((j_cnt + 1023) // 1024), "ms": a.elapsed_time(b) / args.reps,
                    "evals": int(work[j_lo:j_lo + j_cnt].sum())})
    print(json.dumps({"world": args.world, "slices": res, "max_ms": max(r["ms"] for r in res), "sum_ms": sum(r["ms"] for r in res)}))


if __name__ == "__main__":
    main()

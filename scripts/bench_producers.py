#!/usr/bin/env python
"""Stage timing of the batched producer path: host trajectories (threaded) ->
device amplitudes / Ylm / mode selection / compaction -> fused likelihood, against the host-producer path.
Workload: configs[4]-style draws (1 yr, dt = 10 s, eps = 1e-2, plunging), B walkers per call.
  python scripts/bench_producers.py [--batch 64] [--reps 5] [--eps 1e-2]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--eps", type=float, default=1e-2)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    import torch
    from emri_frequencydomainwaveforms_b200 import _hostlib, _lib, engine
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    T, dt, eps, B = 1.0, 10.0, args.eps, args.batch
    gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True),
                               return_list=True)
    base = gen.waveform_generator
    rng = np.random.default_rng(2601996)
    P = np.zeros((B, 14))
    i = 0
    while i < B:
        M = np.exp(rng.uniform(np.log(1e5), np.log(1e7)))
        mu = M * np.exp(rng.uniform(np.log(1e-6), np.log(1e-4)))
        e0 = rng.uniform(0.001, 0.7)
        try:
            p0 = get_p_at_t(base.inspiral_generator, T * 0.99, [M, mu, 0.0, e0, 1.0], xtol=1e-9, bounds=[7.2 + 2 * e0 + 0.05, 16.0 + 2 * e0])
        except ValueError:
            continue
        P[i] = [M, mu, 0.0, p0, e0, 1.0, 1.0, *rng.uniform(0.2, 2.8, 4), rng.uniform(0, 6.28), 0.0, rng.uniform(0, 6.28)]
        i += 1
    h = _lib.get_handle()
    inj = gen(*P[0], T=T, dt=dt, eps=eps, mask_positive=True)
    n = inj[0].shape[0]
    N = 2 * n - 1
    wfac = np.full((2, n), 2.0e19)
    data = np.stack([c.cpu().numpy() for c in inj]) * wfac
    res = {"batch": B, "eps": eps, "N": N, "host_threads": len(os.sched_getaffinity(0))}
    for mode in ("device", "host"):
        tm = FDTemplateModel(gen, producers=mode)
        tm.set_data(data, wfac)
        tm.get_ll(P, T=T, dt=dt, eps=eps, N=N)      # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = args.reps if mode == "device" else max(1, args.reps // 3)
        for _ in range(reps):
            ll = tm.get_ll(P, T=T, dt=dt, eps=eps, N=N)
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) / reps
        res[f"get_ll_{mode}_ms"] = 1e3 * el
        res[f"likelihoods_per_s_{mode}"] = B / el
        res[f"h2d_bytes_{mode}"] = tm.last_h2d_bytes
        res[f"ll0_{mode}"] = float(ll[0])
    # stage timing of the device-producer path
    ang = np.array([gen._transform(*row[7:11]) for row in P])
    ig = base.inspiral_generator
    for nt in sorted({1, 4, args.threads or min(len(os.sched_getaffinity(0)), 16)}):
        t0 = time.perf_counter()
        for _ in range(args.reps):
            _hostlib.trajectory_batch(P[:, 0], P[:, 1], P[:, 3], P[:, 4], P[:, 11], P[:, 13], T, ig.rtol, ig.atol, ig.max_init_len, nthreads=nt)
        res[f"trajectory_ms_{nt}thr"] = 1e3 * (time.perf_counter() - t0) / args.reps
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        db, ok = base.prepare_batch_device(P[:, 0], P[:, 1], P[:, 3], P[:, 4], ang[:, 0], ang[:, 1], dist=P[:, 6], Phi_phi0=P[:, 11],
                                           Phi_r0=P[:, 13], T=T, dt=dt, eps=eps, cos2psi=ang[:, 2], sin2psi=ang[:, 3], handle=h)
    torch.cuda.synchronize()
    res["prepare_batch_device_ms"] = 1e3 * (time.perf_counter() - t0) / args.reps
    res["knots"], res["modes_kept_mean"] = int(db.pb.n_knots), float(db.pb.n_modes / db.pb.B)
    # device stages alone (CUDA events)
    basis = base._device_basis(h)
    Mb, Mneg = base.num_teuk_modes, int(base.m0mask.sum())
    from emri_frequencydomainwaveforms_b200.utils.ylm import ylm_batch_device
    p_dev, e_dev = torch.from_numpy(db.p_e_host[0]).cuda(), torch.from_numpy(db.p_e_host[1]).cuda()
    sw = torch.from_numpy(np.repeat(np.arange(db.pb.B, dtype=np.int32), db.pb.walkers["L"])).cuda()

    def ev_time(fn):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(args.reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.reps, out
    res["amplitude_ms"], teuk_full = ev_time(lambda: base.amplitude_generator.device_call(p_dev, e_dev, h.torch_device).contiguous())
    res["ylm_ms"], ylm_full = ev_time(lambda: ylm_batch_device(basis["l"], basis["m"], basis["neg_src"], ang[:, 0], ang[:, 1], h))
    flags = torch.empty((db.pb.B, Mb), dtype=torch.uint8, device=h.torch_device)
    res["mode_select_ms"], _ = ev_time(lambda: h.check(h.lib.emrifd_mode_select(h.h, teuk_full.data_ptr(), teuk_full.shape[0], Mb, sw.data_ptr(),
                                                                                  ylm_full.data_ptr(), basis["neg_src"].data_ptr(), Mneg, db.pb.B, eps,
                                                                                  flags.data_ptr())))
    res["fused_likelihood_ms"], _ = ev_time(lambda: engine.run_loglike(db, N, 1.0 / (N * dt)))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

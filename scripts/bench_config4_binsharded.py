"""BASELINE.json configs[3] "high-mode stress": eps = 1e-5 (or every mode), e0 = 0.7, T = 4 yr, dt = 10 s full grid
(N = 12 623 261, 6.3e6 bins), ONE long waveform sharded by frequency-bin slice across the GPUs of one box with an
NCCL all_reduce of the likelihood partial sums.  Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      scripts/bench_config4_binsharded.py [--modes all|eps] [--T 4.0] [--steps 5]

Prints one JSON line (rank 0): time per likelihood (max over ranks, CUDA events), evaluations, per-rank work balance.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="all", choices=["all", "eps"])
    ap.add_argument("--T", type=float, default=4.0)
    ap.add_argument("--dt", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--sharding", default="cyclic", choices=["cyclic", "contiguous"],
                    help="cyclic: interleaved tile ownership (emrifd_batch_sum_cyclic); contiguous: work-balanced bin slices")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from emri_frequencydomainwaveforms_b200 import _lib, engine, distributed as D
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
    from emri_frequencydomainwaveforms_b200.utils.constants import YRSID_SI

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
    h = _lib.get_handle(local)
    dev = h.torch_device
    gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
    M, mu, e0 = 1e6, 10.0, 0.7
    p0 = get_p_at_t(gen.inspiral_generator, args.T * 0.99, [M, mu, 0.0, e0, 1.0], xtol=1e-9, bounds=[7.2 + 2 * e0 + 0.05, 16.0 + 2 * e0])
    it = gen.prepare(M, mu, p0, e0, 1.0, -np.pi / 2, dist=1.0, T=args.T, dt=args.dt,
                     eps=1e-5, mode_selection="all" if args.modes == "all" else None)
    n_pts = int(args.T * YRSID_SI / args.dt) + 1
    N = n_pts + 1 if n_pts % 2 == 0 else n_pts
    n = (N + 1) // 2
    val = 1.0 / (N * args.dt)
    db = engine.DeviceBatch(engine.PackedBatch([it]), h)
    # data = the waveform itself (materialised once, single GPU) -> ll must come out ~0
    hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)
    f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
    wf1 = torch.sqrt(torch.full((n,), val, dtype=torch.float64, device=dev) / get_sensitivity(f_pos))
    wfac = torch.stack([wf1, wf1]).contiguous()
    data_w = (torch.cat([hp, hc], dim=0) * wfac).contiguous()
    h.check(h.lib.emrifd_set_data(h.h, data_w.data_ptr(), wfac.data_ptr(), n))
    del hp, hc
    times = []
    slices = None
    for s in range(args.warmup + args.steps):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0_.record()
        if args.sharding == "cyclic":
            red = D.gpu_bin_sharded_loglike_cyclic(db, N, val)
        else:
            red, slices = D.gpu_bin_sharded_loglike(db, N, val, slices=slices)   # partition built in the first warm-up step, then reused
        e1_.record(); torch.cuda.synchronize()
        t = torch.tensor([e0_.elapsed_time(e1_)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if s >= args.warmup:
            times.append(t.item())
    h.status()
    work = D.bin_work_histogram(db.branches_host(), it["m_arr"], N)
    if args.sharding == "cyclic":
        tile = h.lib.emrifd_tile_bins()
        owner = (np.arange(len(work)) // tile) % world
        per_rank = [int(work[owner == r].sum()) for r in range(world)]
        slices = [(None, int((owner == r).sum())) for r in range(world)]
    else:
        per_rank = [int(work[lo:lo + c].sum()) for lo, c in slices]
    if rank == 0:
        r = red.cpu().numpy()[0]
        print(json.dumps({"config": "configs[3] high-mode stress, frequency-bin sharded", "n_gpus": world, "T_yr": args.T, "N": N,
                          "modes": int(len(it["m_arr"])), "knots": int(len(it["t"])), "evals": int(work.sum()),
                          "ms_per_likelihood": float(np.median(times)), "ms_all": times, "ll": float(r[0]), "hh": float(r[2]),
                          "sharding": args.sharding, "evals_per_rank": per_rank, "bins_per_rank": [c for _, c in slices],
                          "includes": "spline build + segmentation + sharded mode sum + NCCL all_reduce" + ("" if args.sharding == "cyclic" else
                                      " (work-balanced partition built once in warm-up and reused)")}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

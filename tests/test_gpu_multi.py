"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on a single-GPU box): frequency-bin sharding with an
NCCL all_reduce of the likelihood sums, and walker sharding, must reproduce the single-GPU result."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from helpers import make_item
        from emri_frequencydomainwaveforms_b200 import _lib, engine, distributed as D
        from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
        gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
        items = [make_item(gen, "plunge", dt=20.0), make_item(gen, "plunge", dt=20.0, Phi_phi0=0.31)]
        N = items[0]["N"]
        n = (N + 1) // 2
        val = 1.0 / (N * 20.0)
        h = _lib.get_handle(rank)
        db = engine.DeviceBatch(engine.PackedBatch(items), h)
        hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)
        w = torch.full((2, n), 2.0e19, dtype=torch.float64, device=h.torch_device)
        # data = signal + noise that is non-zero in EVERY bin: tiles no harmonic touches then carry a likelihood term too, so a
        # rank that double-counted a truncated tile (or skipped one) would show up in the all_reduced sums
        gen_ = torch.Generator(device="cpu").manual_seed(1234)
        noise = torch.complex(torch.randn((2, n), generator=gen_, dtype=torch.float64), torch.randn((2, n), generator=gen_, dtype=torch.float64))
        dw = (torch.stack([hp[0], hc[0]]) * w + 0.05 * float((hp[0].abs() * w[0]).max()) * noise.to(h.torch_device)).contiguous()
        h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), w.data_ptr(), n))
        single = engine.run_loglike(db, N, val).cpu().numpy()
        red, slices = D.gpu_bin_sharded_loglike(db, N, val)
        red_cyc = D.gpu_bin_sharded_loglike_cyclic(db, N, val)
        # walker sharding: each rank evaluates its block through the host-buffer call, results are gathered
        lo, hi = D.shard_range(len(items), world, rank)
        local = engine.run_loglike_host(engine.PackedBatch(items[lo:hi]), h, N, val)[:, 0]
        counts = [D.shard_range(len(items), world, r)[1] - D.shard_range(len(items), world, r)[0] for r in range(world)]
        gathered = D.gather_walker_results(torch.as_tensor(local, device=h.torch_device), counts).cpu().numpy()
        q.put((rank, single, red.cpu().numpy(), slices, gathered, red_cyc.cpu().numpy()))
    finally:
        dist.destroy_process_group()


def test_bin_and_walker_sharding_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, single0, red0, slices, g0, cyc0), (_, single1, red1, _, g1, cyc1) = res
    assert np.array_equal(single0, single1) and np.array_equal(red0, red1)
    scale = np.abs(single0[:, 0:1]) + np.abs(single0[:, 2:3])
    assert np.all(np.abs(red0 - single0) <= 1e-12 * scale)                  # same sums, different partial order
    assert np.array_equal(cyc0, cyc1) and np.all(np.abs(cyc0 - single0) <= 1e-12 * scale)   # cyclic tile ownership
    assert single0[0, 0] < 0 and single0[1, 0] < single0[0, 0]      # walker 0 is the injected signal (only the noise is left), walker 1 is off
    assert slices[0][1] > 0 and slices[1][1] > 0 and slices[0][1] + slices[1][1] == (len(g0) and sum(c for _, c in slices))
    assert np.allclose(g0, single0[:, 0], rtol=1e-12, atol=1e-12 * scale.max()) and np.array_equal(g0, g1)

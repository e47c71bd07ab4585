"""world_size-2 gloo tests of the multi-GPU host logic (walker sharding, bin sharding + all_reduce).
The compute step is played by the CPU oracle (tests may use it); the GPU path plugs its kernels into
the same helpers (distributed.gpu_bin_sharded_loglike, bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

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
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import make_item, oracle_waveform
        from oracle.oracle import Oracle
        from emri_frequencydomainwaveforms_b200 import distributed as D
        from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
        gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
        orc = Oracle("f64")
        it = make_item(gen, "plunge", dt=100.0)
        N = it["N"]
        n = (N + 1) // 2
        hp, hc, coeff, br, nbr = oracle_waveform(orc, it)
        rng = np.random.default_rng(0)
        w = np.full((2, n), 2.0e19)
        dw = (np.stack([hp[n - 1:], hc[n - 1:]]) + 1e-21 * (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n)))) * w
        full = orc.loglike(dw, np.stack([hp[n - 1:], hc[n - 1:]]), w)

        # ---- frequency-bin sharding: slices balanced by work, all_reduce of the three sums ----
        work = D.bin_work_histogram(br, it["m_arr"], N)
        assert work.sum() == orc.last_n_eval            # the histogram counts every stationary point once
        slices = D.balanced_bin_slices(work, world)
        assert slices[0][0] == 0 and sum(c for _, c in slices) == n and slices[1][0] == slices[0][1]

        def partial(j_lo, j_cnt):
            a, b, *_ = oracle_waveform(orc, it, out_lo=n - 1 + j_lo, out_n=j_cnt)
            return orc.loglike(dw[:, j_lo:j_lo + j_cnt], np.stack([a, b]), w[:, j_lo:j_lo + j_cnt])[None, :]

        red = D.bin_sharded_sums(partial, slices).numpy()[0]
        cost = [work[lo:lo + c].sum() for lo, c in slices]

        # ---- walker sharding: 5 walkers over 2 ranks, gathered on every rank ----
        params = np.arange(5, dtype=np.float64)[:, None] * 0.01

        def ll_block(p):
            out = []
            for row in p:
                it2 = dict(it, Phi_phi=it["Phi_phi"] + row[0])
                a, b, *_ = oracle_waveform(orc, it2, out_lo=n - 1, out_n=n)
                out.append(orc.loglike(dw, np.stack([a, b]), w)[0])
            return np.asarray(out)

        gathered = D.walker_sharded_loglike(params, ll_block)
        ref = ll_block(params) if rank == 0 else None
        q.put((rank, full, red, cost, gathered, ref))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_walker_and_bin_sharding_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, full0, red0, cost0, g0, ref0), (_, full1, red1, cost1, g1, _) = res
    assert np.array_equal(red0, red1)                                  # every rank holds the reduced sums
    assert np.allclose(red0, full0, rtol=1e-12, atol=1e-12 * abs(full0[2]))
    assert abs(cost0[0] - cost0[1]) <= 0.05 * (cost0[0] + cost0[1])    # work-balanced, not bin-balanced
    assert np.array_equal(g0, g1) and g0.shape == (5,)
    assert np.allclose(g0, ref0, rtol=1e-13)


def test_shard_helpers():
    from emri_frequencydomainwaveforms_b200 import distributed as D
    for n in (0, 1, 5, 16, 1024):
        for world in (1, 2, 4, 8):
            blocks = [D.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    work = np.zeros(1000, dtype=np.int64)
    work[100:200] = 50
    sl = D.balanced_bin_slices(work, 4, align=1)
    assert sum(c for _, c in sl) == 1000 and all(100 <= lo <= 200 for lo, _ in sl[1:])
    work = np.zeros(100000, dtype=np.int64)
    work[20000:30000] = 7
    sl = D.balanced_bin_slices(work, 8)                        # tile-aligned starts (mode-sum tiles of 1024 bins)
    assert sum(c for _, c in sl) == 100000 and all(lo % 1024 == 0 for lo, _ in sl)
    assert all(19000 <= lo <= 31000 for lo, _ in sl[1:])


def test_cyclic_tile_ownership_covers_every_tile_once():
    from emri_frequencydomainwaveforms_b200 import distributed as D
    for world in (1, 2, 8):
        ntiles = 1541
        seen = np.zeros(ntiles, dtype=int)
        for r in range(world):
            first, stride = D.cyclic_tile_owner(world, r)
            seen[first::stride] += 1
        assert np.all(seen == 1)
    with pytest.raises(ValueError):
        D.cyclic_tile_owner(4, 4)


def test_balanced_walker_assignment():
    """Cost-sorted capacity-constrained dealing of walkers to ranks: equal counts, every walker exactly once, and a far better balance than
    contiguous blocks on a heavy-tailed cost distribution (the bench's synthetic draws vary 10x in work)."""
    from emri_frequencydomainwaveforms_b200 import distributed as D
    rng = np.random.default_rng(4)
    for world in (2, 4, 8):
        cost = np.exp(rng.normal(size=64 * world))
        sh = D.balanced_walker_assignment(cost, world)
        assert sorted(i for s in sh for i in s) == list(range(len(cost))) and all(len(s) == 64 for s in sh)
        tot = np.array([cost[s].sum() for s in sh])
        blocks = cost.reshape(world, 64).sum(axis=1)
        assert tot.max() / tot.mean() < 1.02 and tot.max() / tot.mean() < blocks.max() / blocks.mean()
    assert D.balanced_walker_assignment([3.0, 1.0, 2.0], 1) == [[0, 2, 1]]
    assert [len(s) for s in D.balanced_walker_assignment(np.ones(10), 4)] == [3, 3, 2, 2]


def test_bench_deals_every_pooled_walker_exactly_once(monkeypatch):
    """bench.bench_batches for N > 1: the pools of all ranks are dealt per batch index by estimated cost -- every pooled walker goes
    to exactly one rank, every rank gets B walkers per batch, and the per-rank work estimates are balanced to a few per cent."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    rng = np.random.default_rng(11)
    B, nbatch, world = 6, 2, 4

    def fake_pool(g, B_, nbatch_=bench.NBATCH, wait_s=0.0):
        r = np.random.default_rng(100 + g)
        pool = []
        for k in range(nbatch_):
            items = []
            for w in range(B_):
                L = 12
                scale = np.exp(r.normal())
                items.append({"f_phi": np.linspace(1e-3, 1e-3 * (1 + scale), L), "f_r": np.linspace(7e-4, 7e-4 * (1 + 0.5 * scale), L),
                              "m_arr": np.array([2, 2, 1]), "n_arr": np.array([0, 1, -1]), "tag": (g, k, w)})
            pool.append(items)
        return pool

    monkeypatch.setattr(bench, "draw_pool", fake_pool)
    shares = [bench.bench_batches(r, B, nbatch, world=world) for r in range(world)]
    from emri_frequencydomainwaveforms_b200 import engine
    df = 1.0 / (bench.grid_len() * bench.DT)
    for k in range(nbatch):
        tags = [it["tag"] for r in range(world) for it in shares[r][k]]
        assert len(tags) == world * B and len(set(tags)) == world * B and all(t[1] == k for t in tags)
        assert all(len(shares[r][k]) == B for r in range(world))
        tot = np.array([sum(engine.walker_cost_estimate(it, df) for it in shares[r][k]) for r in range(world)])
        assert tot.max() / tot.mean() < 1.10
    assert bench.bench_batches(0, B, nbatch, world=1)[0][0]["tag"] == (0, 0, 0)     # one GPU: pool 0 as drawn

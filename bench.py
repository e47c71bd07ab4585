#!/usr/bin/env python
"""bench.py -- FD EMRI waveform + likelihood throughput on B200 (driver contract in the task prompt).

Workload (BASELINE.json configs[4] "batch throughput", per-walker settings of configs[1]): synthetic parameter draws
(ln M in [ln 1e5, ln 1e7], ln eta in [ln 1e-6, ln 1e-4], e0 in [0.001, 0.7], p0 fixed so the plunge is at 0.99 T;
check_mode_by_mode.py:125-136,194-213), T = 1 yr, dt = 10 s, eps = 1e-2, N = 3 155 815 (N+ = 1 577 908 bins).  Every walker
of a batch is a DISTINCT draw and NBATCH pre-built batches rotate across the steps (NBATCH x B draws per GPU).
One *step* = one pass of the hot path over a batch of B walkers per GPU: spline build -> segmentation -> (m, n) grouping ->
SPA mode sum writing h+(f), hx(f) on f >= 0 (32 B/bin) fused with the PSD-weighted <d|h>, <h|h>, |d-h|^2 reductions against
an injected signal; with more than one GPU every step issues the all_gather of its B log-likelihoods (the one collective
walker sharding needs; asynchronous, overlapping the next step's kernels, drained before the timed region closes).  `value` = walkers (waveform + likelihood) per second with the packed sparse inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (H2D of the sparse inputs and D2H of the likelihoods inside the timed
region).  Extra keys measured in the same run: `cfg1` (BASELINE configs[0] system, sparse support, HBM-bound), `cfg4_binsharded`
(BASELINE configs[3]: one 4-yr all-mode waveform, frequency-bin sharded over the N GPUs + NCCL all_reduce, strong scaling)
and, for N > 1, `multi_gpu_parity`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

`--impl reference` times the CPU arm on the SAME walkers (same seeds, batch size and rotation): oracle/emrifd_cpu_fast.c, an
optimised double-precision CPU implementation of the path (validated against the binary128 oracle), all host threads.
"""
import argparse
import json
import os
import pickle
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 2601996          # check_mode_by_mode.py:47-48
T_YR, DT, EPS = 1.0, 10.0, 1e-2
NBATCH = 4              # pre-built batches that rotate across the steps


def counters():
    """Hardware-counter figures of the dominant kernel from the committed ncu capture (profiles/r2_counters.json) and the
    flop count of the oracle's inner loop (itemised there)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_counters.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------
# synthetic workload (host side, outside every timed region: the trajectory ODE stays on the host)
# ---------------------------------------------------------------------------------------------
def draw_walkers(n_distinct, n_total, seed, T=T_YR, dt=DT, eps=EPS, workload="plunge"):
    from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # the stand-in producer notice: said once in config["data_note"]
        gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
    rng = np.random.default_rng(seed)
    base = []
    tries = 0
    if workload == "cfg1":   # BASELINE.json configs[0]: M=1e6, mu=10, p0=12, e0=0.35 (does not plunge within 1 yr: sparse support)
        it = gen.prepare(1e6, 10.0, 12.0, 0.35, np.pi / 3, -np.pi / 2, dist=1.0, T=T, dt=dt, eps=eps)
        it["raw"] = (1e6, 10.0, 12.0, 0.35, np.pi / 3)
        base.append(it)
    while len(base) < n_distinct and tries < 50 * n_distinct and workload != "cfg1":
        tries += 1
        M = np.exp(rng.uniform(np.log(1e5), np.log(1e7)))
        mu = M * np.exp(rng.uniform(np.log(1e-6), np.log(1e-4)))
        e0 = rng.uniform(0.001, 0.7)
        theta = np.arccos(rng.uniform(-1, 1))
        try:
            p0 = get_p_at_t(gen.inspiral_generator, T * 0.99, [M, mu, 0.0, e0, 1.0], xtol=1e-9,
                            bounds=[7.2 + 2 * e0 + 0.05, 16.0 + 2 * e0])
            it = gen.prepare(M, mu, p0, e0, theta, -np.pi / 2, dist=1.0, T=T, dt=dt, eps=eps)
        except ValueError:
            continue
        it["raw"] = (M, mu, p0, e0, theta)
        base.append(it)
    if not base:
        raise RuntimeError("no valid parameter draw")
    items = []
    for i in range(n_total):
        b = base[i % len(base)]
        it = dict(b)
        # distinct initial phases: they shift Phi_phi(t), Phi_r(t) by constants, no new ODE solve needed
        dphi, dr = rng.uniform(0, 2 * np.pi), rng.uniform(0, 2 * np.pi)
        it["Phi_phi"] = b["Phi_phi"] + dphi
        it["Phi_r"] = b["Phi_r"] + dr
        it["raw"] = b["raw"] + (dphi, dr)      # (M, mu, p0, e0, theta, Phi_phi0, Phi_r0): the same walker as raw parameters
        items.append(it)
    return items


def _pool_path(g, B, nbatch):
    return os.path.join(tempfile.gettempdir(), f"emrifd_bench_walkers_r{g}_B{B}_n{nbatch}_s{SEED}.pkl")


def draw_pool(g, B, nbatch=NBATCH, wait_s=0.0):
    """The NBATCH batches of B distinct draws with seed index g (cached on disk: the reference arm, which runs first on the same
    box, and the B200 arm draw the very same walkers; with several ranks every rank draws its own pool and reads the others').
    wait_s > 0: another process is drawing this pool right now -- poll for its file before falling back to drawing it here."""
    path = _pool_path(g, B, nbatch)
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                return pickle.load(f)
        except Exception:
            pass
        if time.time() - t0 >= wait_s:
            break
        time.sleep(0.5)
    batches = [draw_walkers(B, B, SEED + 1000 * g + 17 * k) for k in range(nbatch)]
    try:
        with open(path + f".tmp{os.getpid()}", "wb") as f:
            pickle.dump(batches, f)
        os.replace(path + f".tmp{os.getpid()}", path)
    except Exception:
        pass
    return batches


def bench_batches(rank, B, nbatch=NBATCH, world=1):
    """The NBATCH batches of B walkers that `rank` evaluates.  One GPU: pool 0 as drawn.  Several GPUs: the draws of all ranks'
    pools are dealt out per batch index, heaviest first, each to the least-loaded rank (distributed.balanced_walker_assignment on
    engine.walker_cost_estimate, the bins swept by a walker's harmonics), so that every rank carries the same work: a random split
    leaves the heaviest of 8 ranks 13-18 % above the mean, which a synchronous ensemble step pays in full."""
    own = draw_pool(rank, B, nbatch)
    if world == 1:
        return own
    from emri_frequencydomainwaveforms_b200 import distributed as D, engine
    # the other ranks draw their pools at this very moment (torchrun): poll for their files; without them (a single process asked
    # for --gpus N) draw everything here
    wait = 300.0 if int(os.environ.get("WORLD_SIZE", "1")) > 1 else 0.0
    pools = [own if g == rank else draw_pool(g, B, nbatch, wait_s=wait) for g in range(world)]
    df = 1.0 / (grid_len() * DT)
    out = []
    for k in range(nbatch):
        flat = [(g, w) for g in range(world) for w in range(B)]
        cost = np.array([engine.walker_cost_estimate(pools[g][k][w], df) for g, w in flat])
        mine = D.balanced_walker_assignment(cost, world)[rank]
        out.append([pools[flat[i][0]][k][flat[i][1]] for i in mine])
    return out


def grid_len(T=T_YR, dt=DT):
    from emri_frequencydomainwaveforms_b200.utils.constants import YRSID_SI
    n = int(T * YRSID_SI / dt) + 1
    return n + 1 if n % 2 == 0 else n


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def workload_config(batch, workload="plunge"):
    desc = ("configs[4]/[1]: synthetic parameter draws (M 1e5-1e7, eta 1e-6-1e-4, e0 0.001-0.7, p0 set to plunge at 0.99 T), "
            "FD waveform on f>=0 + PSD-weighted likelihood, T=1 yr, dt=10 s, eps=1e-2, N=3155815")
    if workload == "cfg1":
        desc = ("configs[0] system (M=1e6, mu=10, p0=12, e0=0.35; no plunge within 1 yr, sparse support), walkers differ in initial phases, "
                "FD waveform on f>=0 + PSD-weighted likelihood, T=1 yr, dt=10 s, eps=1e-2, N=3155815")
    return {"workload": desc,
            "walkers_per_gpu_per_step": batch, "distinct_draws_per_gpu": batch * NBATCH, "batches_rotating": NBATCH,
            "seed": f"{SEED} + 1000*pool + 17*batch (N > 1: the pools of all ranks dealt out per batch by estimated cost, equal counts and equal work per rank)", "T_yr": T_YR, "dt_s": DT, "eps": EPS, "N": grid_len(),
            "l2": "per-step output (B x 50.5 MB) and inputs exceed the 126 MB L2; no explicit flush needed",
            "parallelism": "walker-sharded; when N > 1 every step issues the NCCL all_gather of its log-likelihoods (asynchronous: it overlaps the next "
                            "step's kernels; the last one is drained inside the timed region)",
            "data_note": "trajectories / amplitudes from the package's offline stand-in producers (FEW's data files are absent)"}


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle/emrifd_cpu_fast.c on the same walkers
# ---------------------------------------------------------------------------------------------
def injected_signal_host(fast, N, val):
    """The injected data exactly as the B200 arm builds them (walker 0 of draw_walkers(1, 1, SEED), LISA PSD), on the host."""
    from scipy.interpolate import CubicSpline
    tab = np.load(os.path.join(ROOT, "emri_frequencydomainwaveforms_b200", "data", "lisa_alloc_sh.npy"))
    n = (N + 1) // 2
    inj = draw_walkers(1, 1, SEED)[0]
    hp, hc, _, _ = fast.sum(inj, N, val)
    wf1 = np.sqrt(val / CubicSpline(tab[:, 0], tab[:, 1])(np.arange(n) * val))
    wfac = np.stack([wf1, wf1])
    return np.stack([hp, hc]) * wfac, wfac


def run_reference(args, guard):
    """--impl reference: the path's CPU implementation (optimised double port, all host threads) on the same batches."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        draw_pool(rank, args.batch)      # (rank 0 deals the walkers out of every rank's pool; nothing is timed here)
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm is meant to use every host thread it can
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    from oracle.oracle import FastCPU
    fast = FastCPU()
    fast.set_num_threads(host_threads())
    cores = fast.num_threads()
    N = grid_len()
    val = 1.0 / (N * DT)
    B = args.batch
    batches = bench_batches(0, B, world=max(1, args.gpus))
    data_w, wfac = injected_signal_host(fast, N, val)

    def step(i):
        out = np.zeros((B, 3))
        for w, it in enumerate(batches[i % NBATCH]):
            hp, hc, like, _ = fast.sum(it, N, val, data_w=data_w, wfac=wfac, want_h=True)   # h+, hx materialised + likelihood, as the GPU step
            out[w] = like
        return out

    for w in range(args.warmup):
        step(w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        ll = step(args.warmup + s)
    el = time.perf_counter() - t0
    value = B * args.steps / el
    line = {"impl": "reference", "metric": "fd_waveform_likelihoods_per_s", "value": value, "unit": "walkers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(B),
            "cpu_baseline": {"value": value, "unit": "walkers/s", "cores": cores, "kind": "port",
                             "sample": f"the B200 arm's own rank-0 batches ({B} walkers/step, {NBATCH} rotating batches), "
                                       "oracle/emrifd_cpu_fast.c: optimised f64 CPU implementation (one solve per (m,n) group, warm-started Newton, "
                                       "OpenMP over bin tiles, -O3 -march=native, FMA), validated <= 1e-9 against the binary128 oracle"},
            "e2e": {"value": value, "unit": "walkers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "ll_checksum": float(np.sum(ll[:, 0]))}
    guard.emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class StdoutGuard:
    """Everything written to fd 1 while the bench runs (NCCL's version banner, library chatter) is sent to stderr, so that
    stdout carries exactly ONE line: the JSON result."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


def main():
    guard = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="walkers per GPU per step (all distinct draws)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg1 / cfg4 / from-parameters legs (profiling runs)")
    ap.add_argument("--workload", default="plunge", choices=["plunge", "cfg1"],
                    help="plunge: the headline batch of plunging draws (FP64-bound); cfg1: configs[0] system, sparse support (HBM-bound)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "b200" else 0)
    if args.impl == "reference":
        return run_reference(args, guard)

    import torch
    import ctypes as C
    from emri_frequencydomainwaveforms_b200 import _lib, engine, distributed as D
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    h = _lib.get_handle(local_rank)
    dev = h.torch_device

    B = args.batch
    N = grid_len()
    n = (N + 1) // 2
    val = 1.0 / (N * DT)
    if args.workload == "cfg1":
        batches = [draw_walkers(1, B, SEED + 1000 * rank + 17 * k, workload="cfg1") for k in range(NBATCH)]
    else:
        batches = bench_batches(rank, B, world=world)
    pbs = [engine.PackedBatch(items) for items in batches]
    dbs = [engine.DeviceBatch(pb, h) for pb in pbs]
    for pb in pbs:
        pb.walkers["out_off"] = np.arange(B, dtype=np.int64) * n

    # injected data = walker 0 of a fixed draw (same on every rank and in the CPU arm), whitened with the LISA PSD
    def set_injection():
        dbi = engine.DeviceBatch(engine.PackedBatch(draw_walkers(1, 1, SEED)), h)
        hp0, hc0, _ = engine.run_waveform(dbi, N, val, mask_positive=True)
        f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
        wf1 = torch.sqrt(torch.full((n,), val, dtype=torch.float64, device=dev) / get_sensitivity(f_pos))
        wf = torch.stack([wf1, wf1]).contiguous()
        dw = (torch.cat([hp0, hc0], dim=0) * wf).contiguous()
        h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), wf.data_ptr(), n))
        return dw, wf
    data_w, wfac = set_injection()

    hp = torch.empty((B, n), dtype=torch.complex128, device=dev)
    hc = torch.empty((B, n), dtype=torch.complex128, device=dev)
    like = torch.empty((B, 3), dtype=torch.float64, device=dev)
    ll_vec = torch.empty(B, dtype=torch.float64, device=dev)
    ll_all = torch.empty(B * world, dtype=torch.float64, device=dev)
    flags = _lib.INCLUDE_MINUS_M | _lib.MASK_POSITIVE

    pending = [None]

    def gather_ll(src):
        """The collective of walker sharding: every rank ends a step with all B x N log-likelihoods (what Eryn's
        compute_log_like returns to the ensemble move, ensemble.py:1283-1318).  Issued asynchronously on NCCL's stream so that it
        overlaps the next step's kernels (steps are independent batches); the previous step's gather is waited for before its
        buffers are reused, and the last one before the timed region closes (drain_ll)."""
        if pending[0] is not None:
            pending[0].wait()
        ll_vec.copy_(src)
        pending[0] = dist.all_gather_into_tensor(ll_all, ll_vec, async_op=True)

    def drain_ll():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None

    def step_device(i):
        pb, db = pbs[i % NBATCH], dbs[i % NBATCH]
        h.check(h.lib.emrifd_fd_waveform_batch(
            h.h, pb.walkers.ctypes.data, B, db.t.data_ptr(), db.teuk.data_ptr(), db.f_phi.data_ptr(), db.f_r.data_ptr(),
            db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(), db.m.data_ptr(), db.n.data_ptr(), db.ylm.data_ptr(), N, val, None,
            flags, db.coeff.data_ptr(), db.branches.data_ptr(), hp.data_ptr(), hc.data_ptr(), like.data_ptr()))
        if dist is not None:
            gather_ll(like[:, 0])

    like_host = np.zeros((B, 3))
    like_host_t = torch.from_numpy(like_host)

    def step_e2e(i):
        pb = pbs[i % NBATCH]
        h.check(h.lib.emrifd_loglike_batch_host(
            h.h, pb.walkers.ctypes.data, B, pb.t.ctypes.data, pb.teuk.ctypes.data, pb.f_phi.ctypes.data, pb.f_r.ctypes.data,
            pb.Phi_phi.ctypes.data, pb.Phi_r.ctypes.data, pb.m.ctypes.data, pb.n.ctypes.data, pb.ylm.ctypes.data, N, val, None,
            flags, hp.data_ptr(), hc.data_ptr(), like_host.ctypes.data))
        if dist is not None:     # host results -> device -> all_gather -> host: what a multi-process sampler would do
            gather_ll(like_host_t[:, 0].to(dev, non_blocking=False))
            drain_ll()
            ll_all.cpu()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, first=0):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = h.launch_count()
        t0 = time.perf_counter()
        ev0.record()
        for s in range(steps):
            fn(first + s)
        if dist is not None:
            drain_ll()           # the last step's all_gather belongs to the timed region
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(ev0.elapsed_time(ev1), 0.0)
        launches = h.launch_count() - l0
        barrier()
        tt = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt[0].item(), tt[1].item(), launches

    # ---- warm-up, then the device-resident timed region --------------------------------------
    for w in range(args.warmup):
        step_device(w)
    h.status()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    ms_dev, _, launches = timed(step_device, args.steps, first=args.warmup)
    # ---- e2e through the host-buffer C-ABI call ----------------------------------------------
    for w in range(2):
        step_e2e(w)
    _, ms_e2e_wall, _ = timed(step_e2e, args.steps, first=2)
    clocks = sampler.finish()
    h.status()
    last = (args.steps + 1) % NBATCH          # batch of the last e2e step: the device path must reproduce its numbers
    step_device(last)
    ll_dev = like[:, 0].cpu().numpy()
    assert np.all(np.isfinite(ll_dev)) and np.allclose(ll_dev, like_host[:, 0], rtol=1e-12, atol=1e-9), "device and e2e paths disagree"

    # ---- dominant kernel (empty_tile + mode_sum bracket) timed live with CUDA events on its own stream --------------
    h.check(h.lib.emrifd_sum_kernel_time(h.h, 1, None, None))
    ksteps = min(max(args.steps, NBATCH), 32)
    for s in range(ksteps):
        step_device(s)
    kms, kmain, kl = C.c_double(), C.c_double(), C.c_int64()
    h.check(h.lib.emrifd_sum_kernel_times(h.h, 0, C.byref(kms), C.byref(kmain), C.byref(kl)))
    k_avg_ms = kms.value / max(kl.value, 1)          # classification + mode_sum_kernel with empty_tile_kernel underneath, up to their join
    k_main_ms = kmain.value / max(kl.value, 1)       # mode_sum_kernel as it runs in the step (the zero-fill stream beside it)
    # the same kernel with the zero-fill moved behind it (what an ncu capture sees)
    h.check(h.lib.emrifd_set_overlap(h.h, 0))
    h.check(h.lib.emrifd_sum_kernel_time(h.h, 1, None, None))
    for s in range(ksteps):
        step_device(s)
    h.check(h.lib.emrifd_sum_kernel_times(h.h, 0, C.byref(kms), C.byref(kmain), C.byref(kl)))
    k_alone_ms = kmain.value / max(kl.value, 1)
    k_pair_seq_ms = kms.value / max(kl.value, 1)
    h.check(h.lib.emrifd_set_overlap(h.h, 1))

    # work counters, averaged over the rotating batches: per-(l,m,n) evaluations (SURVEY's unit), MBE, and the stationary
    # points the kernel actually solves (one per (m, n) group and bin)
    evals = mbe = gevals = 0
    solves_per_batch = []
    for pb, db in zip(pbs, dbs):
        nev = torch.zeros((B, 2), dtype=torch.int64, device=dev)
        h.check(h.lib.emrifd_batch_segment(h.h, pb.walkers.ctypes.data, B, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(),
                                           db.n.data_ptr(), N, val, None, db.branches.data_ptr(), nev.data_ptr()))
        e_, m_ = [int(x) for x in nev.sum(dim=0).cpu().numpy()]
        evals += e_ / NBATCH
        mbe += m_ / NBATCH
        solves_per_batch.append(int(engine.group_evaluations(db).sum()))
        gevals += solves_per_batch[-1] / NBATCH

    gfl = C.c_double()
    h.check(h.lib.emrifd_bench_fp64_fma(h.h, 4096, C.byref(gfl)))
    fp64_peak = gfl.value / 1e3

    extras = {}
    if not args.no_extras:
        extras["e2e_from_parameters"] = leg_from_parameters(args, h, batches, N, n, val, B, world, dist, barrier, dev, data_w, wfac)
        if args.workload == "plunge":
            extras["cfg1"] = leg_cfg1(h, rank, world, dist, barrier, N, n, val, B, flags, hp, hc, like, dev)
        extras["cfg4_binsharded"] = leg_cfg4(h, rank, world, dist, barrier, dev, fp64_peak)
        if dist is not None:
            h.check(h.lib.emrifd_set_data(h.h, data_w.data_ptr(), wfac.data_ptr(), n))
            extras["multi_gpu_parity"] = leg_parity(h, rank, world, dist, N, n, val, dev)
        h.check(h.lib.emrifd_set_data(h.h, data_w.data_ptr(), wfac.data_ptr(), n))

    if dist is not None:
        drain_ll()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    cnt = counters()
    # Algorithmic bytes of the dominant kernel: the 32 B/bin of h+, hx that must be written (SURVEY 8d "waveform").  The
    # fused likelihood's 48 B/bin of data reads are NOT charged: tiles no harmonic touches take their sum |d~|^2 from a
    # table precomputed at emrifd_set_data, so most of those reads never happen.
    alg_bytes = 32.0 * n * B
    ach_gbs = alg_bytes / (k_avg_ms * 1e-3) / 1e9
    # FP64 roofline from EXECUTED flops: (dfma*2 + dmul + dadd) per solved stationary point as counted by ncu on this kernel
    # (profiles/r2_counters.json, smsp__sass_thread_inst_executed_op_d*_pred_on) x the stationary points solved per launch
    fl_exec = float(cnt.get("executed_flops_per_solve", 171.0))
    ach_tflops = fl_exec * gevals / (k_main_ms * 1e-3) / 1e12      # every stationary point is solved inside mode_sum_kernel
    fl_orc = float(cnt.get("oracle_flops_per_mode_eval", 376.0))
    value = world * B * args.steps / (ms_dev * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e_wall * 1e-3)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]
        traffic = tr["dram_bytes_per_launch"] * B / tr["batch"]
    except Exception:
        pass
    roof_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
    roof_fp64 = {"bound": "fp64", "achieved": ach_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tflops / fp64_peak,
                 "traffic": traffic,
                 "flops": "EXECUTED FP64 flops (2 dfma + dmul + dadd, ncu) per solved stationary point x stationary points solved per launch",
                 "executed_flops_per_solve": fl_exec, "solves_per_launch": gevals,
                 "fp64_pipe_active_ncu": cnt.get("fp64_pipe_active"), "issue_active_ncu": cnt.get("issue_active"),
                 "thread_instructions_per_solve_ncu": cnt.get("thread_instructions_per_solve"), "ncu_capture": cnt.get("capture"),
                 "mode_evals_per_launch": evals, "mbe_per_launch": mbe,
                 "reference_formulation_equivalent": {
                     "what": "rate at which the reference formulation's work (one evaluation per (l,m,n) mode and bin, flops counted from the "
                             "oracle's inner loop, profiles/roofline.json) is retired -- NOT a hardware utilisation, may exceed the peak",
                     "oracle_flops_per_mode_eval": fl_orc, "tflops_equivalent": fl_orc * evals / (k_main_ms * 1e-3) / 1e12,
                     "survey_300_per_mbe_convention_tflops": 300.0 * mbe / (k_main_ms * 1e-3) / 1e12},
                 "peak_source": "emrifd_bench_fp64_fma: CUDA-core DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                "no tensor cores on this path)"}
    binding, other = (roof_fp64, roof_hbm) if roof_fp64["frac"] >= roof_hbm["frac"] else (roof_hbm, roof_fp64)
    step_ms = ms_dev / args.steps
    roof_fp64.update({"kernel": "mode_sum_kernel<true,true,2,true> (warp-specialised persistent CTAs; every stationary point is solved here), timed as "
                                "it runs in the step: empty_tile_kernel's store stream shares the SMs with it",
                      "kernel_ms": k_main_ms, "kernel_share_of_step": k_main_ms / step_ms,
                      "kernel_alone": {"kernel_ms": k_alone_ms, "frac": fl_exec * gevals / (k_alone_ms * 1e-3) / 1e12 / fp64_peak,
                                       "launch_pair_ms": k_pair_seq_ms,
                                       "what": "the same kernel with the zero-fill behind it on the same stream (emrifd_set_overlap(0)): the "
                                               "configuration an ncu capture sees; the step is slower that way"}})
    roof_hbm.update({"kernel": "empty_tile_kernel<true,true,2> + mode_sum_kernel<true,true,2,true> (the launch pair that writes h+, hx: store stream "
                               "for the empty tiles, persistent CTAs for the others)",
                     "kernel_ms": k_avg_ms, "kernel_share_of_step": k_avg_ms / step_ms})
    binding = dict(binding, launch_pair_ms=k_avg_ms, launch_pair_share_of_step=k_avg_ms / step_ms)

    line = {
        "metric": "fd_waveform_likelihoods_per_s", "value": value, "unit": "walkers/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(B, args.workload),
        "e2e": {"value": e2e_value, "unit": "walkers/s", "h2d_bytes_per_step": pbs[0].h2d_bytes(),
                "d2h_bytes_per_step": int(like_host.nbytes) + (8 * B * world if world > 1 else 0),
                "ms_per_step": ms_e2e_wall / args.steps,
                "call": "emrifd_loglike_batch_host (host packed sparse inputs -> H2D -> spline/segment/group/sum+likelihood -> D2H ll)"
                        + (" + all_gather of ll" if world > 1 else "")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": binding,
        "roofline_other_roof": other,
        "work": {"evals_per_walker": evals / B, "group_evals_per_walker": gevals / B, "mbe_per_walker": mbe / B,
                 "solves_per_batch": solves_per_batch,
                 "modes_per_walker": float(np.mean([pb.n_modes for pb in pbs])) / B, "knots_per_walker": float(np.mean([pb.n_knots for pb in pbs])) / B},
    }
    line.update(extras)
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is an N = 1 figure (rank 0 would stall the other ranks' exit)
        try:
            from oracle.oracle import FastCPU
            fast = FastCPU()
            fast.set_num_threads(host_threads())
            dwh, wfh = data_w.cpu().numpy().view(np.complex128).reshape(2, n), wfac.cpu().numpy()
            t0 = time.perf_counter()
            done, worst = 0, 0.0
            order = [last] + [k for k in range(NBATCH) if k != last]
            while time.perf_counter() - t0 < 10.0:            # a bounded sample: whole batches until ~10 s of CPU work
                for k in order:
                    for w, it in enumerate(batches[k]):
                        _, _, lk, _ = fast.sum(it, N, val, data_w=dwh, wfac=wfh, want_h=True)
                        if k == last and done < B:            # the batch the GPU evaluated last: same walkers, same data
                            worst = max(worst, abs(lk[0] - ll_dev[w]) / abs(lk[2] + abs(lk[0])))
                        done += 1
                    if time.perf_counter() - t0 >= 10.0:
                        break
            el = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": done / el, "unit": "walkers/s", "cores": fast.num_threads(), "kind": "port",
                                    "sample": f"{done} walkers (whole batches of this run's rotating draws, the GPU's last batch first), "
                                              f"oracle/emrifd_cpu_fast.c (optimised f64 CPU implementation, OpenMP over bin tiles, -march=native), {el:.1f} s",
                                    "max_rel_ll_difference_vs_gpu": worst}
        except Exception as exc:   # test infrastructure: never let it break the bench line
            line["cpu_baseline"] = {"value": None, "unit": "walkers/s", "cores": 0, "kind": "port", "sample": f"unavailable: {exc}"}
    guard.emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# extra legs (same run, outside the headline timed regions)
# ---------------------------------------------------------------------------------------------
def leg_from_parameters(args, h, batches, N, n, val, B, world, dist, barrier, dev, data_w, wfac):
    """Public-API path from RAW PARAMETERS: FDTemplateModel.get_ll(params[NBATCH*B, 14]) -- host trajectory ODE (threaded,
    cores / ranks threads) -> H2D of the sparse tracks -> device amplitudes / Ylm / mode selection / compaction -> spline /
    segment / sum + likelihood -> D2H of ll.  Everything a user's likelihood call pays is inside."""
    import torch
    try:
        import warnings
        from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
        from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True),
                                       return_list=True, frame="source")
        model = FDTemplateModel(gen)
        model.set_data(data_w, wfac)        # the injection of the headline legs
        raw = np.array([it["raw"] for items in batches for it in items])
        P = np.zeros((len(raw), 14))
        P[:, 0], P[:, 1], P[:, 3], P[:, 4], P[:, 5], P[:, 6] = raw[:, 0], raw[:, 1], raw[:, 2], raw[:, 3], 1.0, 1.0
        # source frame: (qS, phiS) are the viewing angles (theta, phi) themselves
        P[:, 7], P[:, 8], P[:, 11], P[:, 13] = raw[:, 4], -np.pi / 2, raw[:, 5], raw[:, 6]
        P = np.tile(P, (4, 1))              # configs[4]: 1024 walkers per likelihood call (16 chunks of 64: the pipeline reaches steady state)
        kw = dict(T=T_YR, dt=DT, eps=EPS, N=N)
        model.get_ll(P, **kw)               # warm-up of the pipelined path (side handles, allocator pools of the side streams)
        reps = 6
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            ll = model.get_ll(P, **kw)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        barrier()
        tt = torch.tensor([el], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return {"value": world * len(P) * reps / tt[0].item(), "unit": "walkers/s", "walkers_per_call": len(P), "ms_per_call": 1e3 * tt[0].item() / reps,
                "h2d_bytes_per_call": int(model.last_h2d_bytes), "d2h_bytes_per_call": int(ll.nbytes + 4 * len(P)),
                "host_threads": host_threads(), "finite": bool(np.all(np.isfinite(ll))),
                "call": "FDTemplateModel.get_ll(params) (the lisatools Likelihood plug-in): host trajectory ODE (threaded, chunks of 64 walkers "
                        "pipelined against the device work of the previous chunk) -> H2D tracks -> device amplitudes/Ylm/mode selection/compaction "
                        "-> spline/segment/group/sum+likelihood -> D2H ll"}
    except Exception as exc:   # auxiliary figure: never let it break the bench line
        return {"value": None, "unit": "walkers/s", "error": str(exc)[:300]}


def leg_cfg1(h, rank, world, dist, barrier, N, n, val, B, flags, hp, hc, like, dev):
    """BASELINE configs[0] system (sparse support: the HBM-bound regime), B walkers per GPU that differ in their phases."""
    import ctypes as C
    import torch
    from emri_frequencydomainwaveforms_b200 import engine
    try:
        items = draw_walkers(1, B, SEED + 7 + 1000 * rank, workload="cfg1")
        pb = engine.PackedBatch(items)
        db = engine.DeviceBatch(pb, h)
        pb.walkers["out_off"] = np.arange(B, dtype=np.int64) * n

        def step():
            h.check(h.lib.emrifd_fd_waveform_batch(
                h.h, pb.walkers.ctypes.data, B, db.t.data_ptr(), db.teuk.data_ptr(), db.f_phi.data_ptr(), db.f_r.data_ptr(),
                db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(), db.m.data_ptr(), db.n.data_ptr(), db.ylm.data_ptr(), N, val, None,
                flags, db.coeff.data_ptr(), db.branches.data_ptr(), hp.data_ptr(), hc.data_ptr(), like.data_ptr()))
        for _ in range(3):
            step()
        reps = 20
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        h.check(h.lib.emrifd_sum_kernel_time(h.h, 1, None, None))
        ev0.record()
        for _ in range(reps):
            step()
        ev1.record()
        torch.cuda.synchronize()
        kms, kmain, kl = C.c_double(), C.c_double(), C.c_int64()
        h.check(h.lib.emrifd_sum_kernel_times(h.h, 0, C.byref(kms), C.byref(kmain), C.byref(kl)))
        tt = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt[0].item() / reps
        k_ms = kms.value / max(kl.value, 1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        gbs = 32.0 * n * B / (k_ms * 1e-3) / 1e9
        return {"walkers_per_s": world * B / (ms * 1e-3), "ms_per_step": ms, "kernel_ms": k_ms, "mode_sum_ms": kmain.value / max(kl.value, 1),
                "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak,
                "hbm_peak": hbm_peak, "algorithmic_bytes_per_launch": 32.0 * n * B, "group_evals_per_walker": float(engine.group_evaluations(db).mean()),
                "what": "configs[0] system (M=1e6, mu=10, p0=12, e0=0.35, 1 yr, eps=1e-2): waveform on f>=0 + likelihood, "
                        f"{B} walkers/GPU/step; kernel = empty_tile + mode_sum bracket, charged 32 B/bin of h+, hx written"}
    except Exception as exc:
        return {"walkers_per_s": None, "error": str(exc)[:300]}


def cfg4_walker():
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
    M, mu, e0, T = 1e6, 10.0, 0.7, 4.0
    p0 = get_p_at_t(gen.inspiral_generator, T * 0.99, [M, mu, 0.0, e0, 1.0], xtol=1e-9, bounds=[7.2 + 2 * e0 + 0.05, 16.0 + 2 * e0])
    return gen.prepare(M, mu, p0, e0, 1.0, -np.pi / 2, dist=1.0, T=T, dt=DT, mode_selection="all"), T


def leg_cfg4(h, rank, world, dist, barrier, dev, fp64_peak):
    """BASELINE configs[3]: ONE 4-yr waveform with all 3843 (l,m,n) modes (671 (m,n) groups), N = 12 623 261, likelihood
    sharded by frequency-bin tile over the N GPUs (cyclic tile ownership) + one NCCL all_reduce of three doubles.  Strong
    scaling: the 1-rank time is measured in the same run (every rank does the full sum once)."""
    import torch
    from emri_frequencydomainwaveforms_b200 import engine, distributed as D
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    try:
        it, T = cfg4_walker()
        N = grid_len(T)
        n = (N + 1) // 2
        val = 1.0 / (N * DT)
        db = engine.DeviceBatch(engine.PackedBatch([it]), h)
        hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)      # data = the waveform itself -> ll ~ 0
        f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
        wf1 = torch.sqrt(torch.full((n,), val, dtype=torch.float64, device=dev) / get_sensitivity(f_pos))
        wf = torch.stack([wf1, wf1]).contiguous()
        dw = (torch.cat([hp, hc], dim=0) * wf).contiguous()
        h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), wf.data_ptr(), n))
        del hp, hc

        def time_it(fn, reps=5, warm=2):
            ts = []
            for s in range(warm + reps):
                barrier()
                e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0_.record()
                out = fn()
                e1_.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0_.elapsed_time(e1_)], dtype=torch.float64, device=dev)
                if dist is not None:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if s >= warm:
                    ts.append(t.item())
            return float(np.median(ts)), out
        ms1, single = time_it(lambda: engine.run_loglike(db, N, val))          # full sum on one GPU (every rank does the same work)
        single = single.cpu().numpy()[0]
        br = db.branches_host()
        evals = int(np.where(br["end"] >= br["start"], br["end"] - br["start"] + 1, 0).sum())
        gevals = int(engine.group_evaluations(db)[0])
        out = {"modes": int(len(it["m_arr"])), "knots": int(len(it["t"])), "N": N, "mode_evals": evals, "solves": gevals,
               "ms_per_likelihood_1rank": ms1, "ll_1rank": float(single[0]), "hh": float(single[2]),
               "g_solves_per_s_1rank": gevals / (ms1 * 1e-3) / 1e9,
               "includes": "spline build + segmentation + (m,n) grouping + mode sum + likelihood reduction"}
        if dist is not None:
            msN, red = time_it(lambda: D.gpu_bin_sharded_loglike_cyclic(db, N, val))
            red = red.cpu().numpy()[0]
            small = torch.zeros((1, 3), dtype=torch.float64, device=dev)
            ar = []
            for s in range(25):
                barrier()
                e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0_.record()
                dist.all_reduce(small, op=dist.ReduceOp.SUM)
                e1_.record()
                torch.cuda.synchronize()
                if s >= 5:
                    ar.append(e0_.elapsed_time(e1_) * 1e3)
            ok = bool(abs(red[0] - single[0]) <= 1e-12 * abs(single[2]) and abs(red[2] - single[2]) <= 1e-12 * abs(single[2])
                      and abs(red[1] - single[1]) <= 1e-12 * abs(single[2]))
            out.update({"n_gpus": world, "ms_per_likelihood": msN, "speedup_vs_1rank": ms1 / msN, "all_reduce_us": float(np.median(ar)),
                        "sharding": "cyclic tile ownership (emrifd_batch_sum_cyclic) + NCCL all_reduce(SUM) of [1,3] doubles",
                        "parity_vs_1rank_1e-12": ok, "ll": float(red[0])})
        else:
            out.update({"n_gpus": 1, "ms_per_likelihood": ms1, "speedup_vs_1rank": 1.0})
        return out
    except Exception as exc:
        return {"ms_per_likelihood": None, "error": str(exc)[:300]}


def leg_parity(h, rank, world, dist, N, n, val, dev):
    """N > 1 only (the 1-GPU test box always skips tests/test_gpu_multi.py): the walker-sharded all_gather of ll and the
    frequency-bin sharded likelihood + NCCL all_reduce must equal the rank-local single-GPU result to 1e-12."""
    import torch
    from emri_frequencydomainwaveforms_b200 import engine, distributed as D
    try:
        nw = 2 * world
        items = draw_walkers(4, nw, SEED + 99)                 # the SAME walkers on every rank
        db = engine.DeviceBatch(engine.PackedBatch(items), h)
        local_all = engine.run_loglike(db, N, val).cpu().numpy()
        scale = np.abs(local_all[:, 0:1]) + np.abs(local_all[:, 2:3])
        # walker sharding through the host-buffer call + all_gather
        lo, hi = D.shard_range(nw, world, rank)
        mine = engine.run_loglike_host(engine.PackedBatch(items[lo:hi]), h, N, val)[:, 0]
        counts = [D.shard_range(nw, world, r)[1] - D.shard_range(nw, world, r)[0] for r in range(world)]
        gathered = D.gather_walker_results(torch.as_tensor(mine, device=dev), counts).cpu().numpy()
        ok_w = bool(np.all(np.abs(gathered - local_all[:, 0]) <= 1e-12 * scale[:, 0]))
        # bin sharding, both partitions
        cyc = D.gpu_bin_sharded_loglike_cyclic(db, N, val).cpu().numpy()
        con, _ = D.gpu_bin_sharded_loglike(db, N, val)
        con = con.cpu().numpy()
        ok_c = bool(np.all(np.abs(cyc - local_all) <= 1e-12 * scale))
        ok_s = bool(np.all(np.abs(con - local_all) <= 1e-12 * scale))
        flag = torch.tensor([float(ok_w and ok_c and ok_s)], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert flag.item() == 1.0, f"multi-GPU parity failed on some rank (rank {rank}: walker {ok_w}, cyclic {ok_c}, contiguous {ok_s})"
        return True
    except AssertionError:
        raise
    except Exception as exc:
        return {"ok": False, "error": str(exc)[:300]}


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- FD EMRI waveform + likelihood throughput on B200 (driver contract in the task prompt).

Workload (BASELINE.json configs[4] "batch throughput", per-walker settings of configs[1]):
synthetic parameter draws (ln M in [ln 1e5, ln 1e7], ln eta in [ln 1e-6, ln 1e-4], e0 in [0.001, 0.7],
p0 fixed so the plunge is at 0.99 T; check_mode_by_mode.py:125-136,194-213), T = 1 yr, dt = 10 s,
eps = 1e-2, N = 3 155 815 (N+ = 1 577 908 bins).  One *step* = one pass of the hot path over a batch
of B walkers per GPU: spline build -> segmentation -> SPA mode sum writing h+(f), hx(f) on f >= 0
(32 B/bin) fused with the PSD-weighted <d|h>, <h|h>, |d-h|^2 reductions against an injected signal.
`value` = walkers (waveform + likelihood) per second with the packed sparse inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (H2D of the sparse inputs and D2H of the
likelihoods inside the timed region).  Walkers are sharded across GPUs (weak scaling, no collective
in the data path).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
"""
import argparse
import json
import os
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
# Algorithmic FP64 flop per stationary-point evaluation: SURVEY.md section 8(d)'s itemisation (cubic solve 60,
# 4 Horner splines 24, fdot/fddot 20, arg + divisions 25, K_1/3 factor 70, phase + sincos 65, complex products 36).
# One evaluation serves BOTH signs of m (the mirrored term is the conjugate), so we charge 300 per evaluation, not
# per SURVEY "MBE" (mode, sign of m, bin); the MBE-convention figure is reported separately.
FLOPS_PER_EVAL = 300.0
SURVEY_FLOPS_PER_MBE = 300.0


# ---------------------------------------------------------------------------------------------
# synthetic workload (host side, outside every timed region: the trajectory ODE stays on the host)
# ---------------------------------------------------------------------------------------------
def draw_walkers(n_distinct, n_total, seed, T=T_YR, dt=DT, eps=EPS, workload="plunge"):
    from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux, viewing_angles
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
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


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's double build (OpenMP) on a bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_oracle_rate(items, N, dt, data_w, wfac, max_seconds=20.0, min_walkers=2):
    from oracle.oracle import Oracle
    orc = Oracle("f64")
    orc.lib.orc_set_num_threads(host_threads())
    cores = orc.lib.orc_num_threads()
    n = (N + 1) // 2
    val = 1.0 / (N * dt)
    t0 = time.perf_counter()
    done = 0
    for it in items:
        hp, hc, *_ = orc.fd_sum(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"],
                                it["f_phi"], it["f_r"], N, val, scale=it["scale"], out_lo=n - 1, out_n=n)
        orc.loglike(data_w, np.stack([hp, hc]), wfac)
        done += 1
        if done >= min_walkers and time.perf_counter() - t0 > max_seconds:
            break
    el = time.perf_counter() - t0
    return done / el, cores, done, el


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference(args, guard):
    """--impl reference: the reference algorithm's CPU implementation (oracle port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm is meant to use every host thread it can
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    from oracle import oracle as orc_mod
    orc_mod.build()
    N = grid_len()
    per_step = 2
    items = draw_walkers(min(4, per_step * 2), per_step * (args.steps + args.warmup), SEED)
    n = (N + 1) // 2
    rng = np.random.default_rng(1)
    wfac = np.full((2, n), 1e18)
    data_w = np.zeros((2, n), dtype=np.complex128)
    for w in range(args.warmup):
        cpu_oracle_rate(items[w * per_step:(w + 1) * per_step], N, DT, data_w, wfac, max_seconds=1e9, min_walkers=per_step)
    t0 = time.perf_counter()
    cores = 1
    for s in range(args.steps):
        lo = (args.warmup + s) * per_step
        _, cores, _, _ = cpu_oracle_rate(items[lo:lo + per_step], N, DT, data_w, wfac, max_seconds=1e9, min_walkers=per_step)
    el = time.perf_counter() - t0
    value = per_step * args.steps / el
    line = {"impl": "reference", "metric": "fd_waveform_likelihoods_per_s", "value": value, "unit": "walkers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(per_step),
            "cpu_baseline": {"value": value, "unit": "walkers/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} walkers/step x {args.steps} steps of the bench workload, oracle f64 build, OpenMP"},
            "e2e": {"value": value, "unit": "walkers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    guard.emit(json.dumps(line))


def workload_config(batch, workload="plunge"):
    desc = ("configs[4]/[1]: synthetic parameter draws (M 1e5-1e7, eta 1e-6-1e-4, e0 0.001-0.7, p0 set to plunge at 0.99 T), "
            "FD waveform on f>=0 + PSD-weighted likelihood, T=1 yr, dt=10 s, eps=1e-2, N=3155815")
    if workload == "cfg1":
        desc = ("configs[0] system (M=1e6, mu=10, p0=12, e0=0.35; no plunge within 1 yr, sparse support), walkers differ in initial phases, "
                "FD waveform on f>=0 + PSD-weighted likelihood, T=1 yr, dt=10 s, eps=1e-2, N=3155815")
    return {"workload": desc,
            "walkers_per_gpu_per_step": batch, "T_yr": T_YR, "dt_s": DT, "eps": EPS, "N": grid_len(),
            "l2": "per-step output (B x 50.5 MB) and inputs exceed the 126 MB L2; no explicit flush needed",
            "parallelism": "walker-sharded, no data-path collective"}


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
    ap.add_argument("--batch", type=int, default=64, help="walkers per GPU per step")
    ap.add_argument("--distinct", type=int, default=8, help="distinct (M, mu, e0, p0) draws per GPU (phases vary per walker)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="plunge", choices=["plunge", "cfg1"],
                    help="plunge: the headline batch of plunging draws (FP64-bound); cfg1: configs[0] system, sparse support (HBM-bound)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "b200" else 0)
    if args.impl == "reference":
        return run_reference(args, guard)

    import torch
    import ctypes as C
    from emri_frequencydomainwaveforms_b200 import _lib, engine
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
    items = draw_walkers(args.distinct, B, SEED + 1000 * rank, workload=args.workload)
    pb = engine.PackedBatch(items)
    db = engine.DeviceBatch(pb, h)
    pb.walkers["out_off"] = np.arange(B, dtype=np.int64) * n

    # injected data = walker 0 of rank 0's draw set (same on every rank), whitened with the LISA PSD
    inj_items = draw_walkers(1, 1, SEED)
    dbi = engine.DeviceBatch(engine.PackedBatch(inj_items), h)
    hp0, hc0, _ = engine.run_waveform(dbi, N, val, mask_positive=True)
    f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
    psd = get_sensitivity(f_pos)
    dfv = torch.full((n,), val, dtype=torch.float64, device=dev)
    wf1 = torch.sqrt(dfv / psd)
    wfac = torch.stack([wf1, wf1]).contiguous()
    data_w = (torch.cat([hp0, hc0], dim=0) * wfac).contiguous()
    h.check(h.lib.emrifd_set_data(h.h, data_w.data_ptr(), wfac.data_ptr(), n))
    del dbi, hp0, hc0

    hp = torch.empty((B, n), dtype=torch.complex128, device=dev)
    hc = torch.empty((B, n), dtype=torch.complex128, device=dev)
    like = torch.empty((B, 3), dtype=torch.float64, device=dev)
    flags = _lib.INCLUDE_MINUS_M | _lib.MASK_POSITIVE

    def step_device():
        h.check(h.lib.emrifd_fd_waveform_batch(
            h.h, pb.walkers.ctypes.data, B, db.t.data_ptr(), db.teuk.data_ptr(), db.f_phi.data_ptr(), db.f_r.data_ptr(),
            db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(), db.m.data_ptr(), db.n.data_ptr(), db.ylm.data_ptr(), N, val, None,
            flags, db.coeff.data_ptr(), db.branches.data_ptr(), hp.data_ptr(), hc.data_ptr(), like.data_ptr()))

    like_host = np.zeros((B, 3))

    def step_e2e():
        h.check(h.lib.emrifd_loglike_batch_host(
            h.h, pb.walkers.ctypes.data, B, pb.t.ctypes.data, pb.teuk.ctypes.data, pb.f_phi.ctypes.data, pb.f_r.ctypes.data,
            pb.Phi_phi.ctypes.data, pb.Phi_r.ctypes.data, pb.m.ctypes.data, pb.n.ctypes.data, pb.ylm.ctypes.data, N, val, None,
            flags, hp.data_ptr(), hc.data_ptr(), like_host.ctypes.data))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = h.launch_count()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            fn()
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
    for _ in range(args.warmup):
        step_device()
    h.status()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    ms_dev, _, launches = timed(step_device, args.steps)
    # ---- e2e through the host-buffer C-ABI call ----------------------------------------------
    for _ in range(2):
        step_e2e()
    _, ms_e2e_wall, _ = timed(step_e2e, args.steps)
    clocks = sampler.finish()
    h.status()

    # ---- e2e from RAW PARAMETERS through the generator's batched public call: host trajectories (threaded native ODE)
    #      -> H2D of the sparse tracks -> device amplitudes / Ylm / mode selection / compaction -> the same spline /
    #      segment / sum + likelihood launches -> D2H of ll.  Everything a user's get_ll(params) pays is inside.
    e2e_par = None
    try:
        from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
        genp = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
        raw = np.array([it["raw"] for it in items])

        def step_params():
            dbp, okp = genp.prepare_batch_device(raw[:, 0], raw[:, 1], raw[:, 2], raw[:, 3], raw[:, 4], -np.pi / 2, dist=1.0,
                                                 Phi_phi0=raw[:, 5], Phi_r0=raw[:, 6], T=T_YR, dt=DT, eps=EPS, handle=h)
            dbp.pb.walkers["out_off"] = np.arange(dbp.pb.B, dtype=np.int64) * n
            h.check(h.lib.emrifd_fd_waveform_batch(
                h.h, dbp.pb.walkers.ctypes.data, dbp.pb.B, dbp.t.data_ptr(), dbp.teuk.data_ptr(), dbp.f_phi.data_ptr(), dbp.f_r.data_ptr(),
                dbp.Phi_phi.data_ptr(), dbp.Phi_r.data_ptr(), dbp.m.data_ptr(), dbp.n.data_ptr(), dbp.ylm.data_ptr(), N, val, None,
                flags, dbp.coeff.data_ptr(), dbp.branches.data_ptr(), hp.data_ptr(), hc.data_ptr(), like.data_ptr()))
            return like[:dbp.pb.B].cpu().numpy(), dbp

        for _ in range(2):
            llp, dbp = step_params()
        psteps = max(3, args.steps // 2)
        barrier()
        t0 = time.perf_counter()
        for _ in range(psteps):
            llp, dbp = step_params()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        barrier()
        tt = torch.tensor([el], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_par = {"value": world * B * psteps / tt[0].item(), "unit": "walkers/s", "ms_per_step": 1e3 * tt[0].item() / psteps,
                   "h2d_bytes_per_step": int(dbp.h2d_bytes), "d2h_bytes_per_step": int(llp.nbytes + 4 * B),
                   "host_threads": len(os.sched_getaffinity(0)),
                   "call": "FastSchwarzschildEccentricFlux.prepare_batch_device(raw parameters) + emrifd_fd_waveform_batch: host trajectory ODE "
                           "(threaded) -> H2D tracks -> device amplitudes/Ylm/mode selection/compaction -> spline/segment/sum+likelihood -> D2H ll",
                   "modes_per_walker": dbp.pb.n_modes / dbp.pb.B}
    except Exception as exc:   # auxiliary figure: never let it break the bench line
        e2e_par = {"value": None, "unit": "walkers/s", "error": str(exc)[:200]}

    # ---- dominant kernel (mode_sum) timed live with CUDA events on its own stream --------------
    h.check(h.lib.emrifd_sum_kernel_time(h.h, 1, None, None))
    ksteps = min(args.steps, 32)
    for _ in range(ksteps):
        step_device()
    kms, kl = C.c_double(), C.c_int64()
    h.check(h.lib.emrifd_sum_kernel_time(h.h, 0, C.byref(kms), C.byref(kl)))
    k_avg_ms = kms.value / max(kl.value, 1)

    # work counters: stationary-point evaluations and MBE of this rank's batch
    nev = torch.zeros((B, 2), dtype=torch.int64, device=dev)
    h.check(h.lib.emrifd_batch_segment(h.h, pb.walkers.ctypes.data, B, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(),
                                       db.n.data_ptr(), N, val, None, db.branches.data_ptr(), nev.data_ptr()))
    evals, mbe = [int(x) for x in nev.sum(dim=0).cpu().numpy()]
    gevals = int(engine.group_evaluations(db).sum())    # stationary points actually solved: one per ((m, n) group, bin)
    ll_dev = like[:, 0].cpu().numpy()
    assert np.all(np.isfinite(ll_dev)) and np.allclose(ll_dev, like_host[:, 0], rtol=1e-12, atol=1e-9), "device and e2e paths disagree"

    gfl = C.c_double()
    h.check(h.lib.emrifd_bench_fp64_fma(h.h, 4096, C.byref(gfl)))

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
    # Algorithmic bytes of the dominant kernel: the 32 B/bin of h+, hx that must be written (SURVEY 8d "waveform").  The
    # fused likelihood's 48 B/bin of data reads are NOT charged: tiles no harmonic touches take their sum |d~|^2 from a
    # table precomputed at emrifd_set_data, so most of those reads never happen (charging them gave frac > 1).
    alg_bytes = 32.0 * n * B
    ach_gbs = alg_bytes / (k_avg_ms * 1e-3) / 1e9
    ach_tflops = FLOPS_PER_EVAL * evals / (k_avg_ms * 1e-3) / 1e12
    fp64_peak = gfl.value / 1e3
    value = world * B * args.steps / (ms_dev * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e_wall * 1e-3)
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (scaled by batch), else null
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]
        traffic = tr["dram_bytes_per_launch"] * B / tr["batch"]
    except Exception:
        pass
    roof_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
    roof_fp64 = {"bound": "fp64", "achieved": ach_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tflops / fp64_peak,
                 "traffic": traffic, "flops_per_eval": FLOPS_PER_EVAL, "stationary_point_evals_per_launch": evals,
                 "mbe_per_launch": mbe, "achieved_mbe_convention_300_per_mbe": SURVEY_FLOPS_PER_MBE * mbe / (k_avg_ms * 1e-3) / 1e12,
                 "peak_source": "emrifd_bench_fp64_fma: CUDA-core DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                "no tensor cores on this path)"}
    binding, other = (roof_fp64, roof_hbm) if roof_fp64["frac"] >= roof_hbm["frac"] else (roof_hbm, roof_fp64)
    common = {"kernel": "empty_tile_kernel<true,true> + mode_sum_kernel<true,true> (one event bracket: store stream for empty tiles, "
                        "persistent CTAs for the others)", "kernel_ms": k_avg_ms, "kernel_share_of_step": k_avg_ms / (ms_dev / args.steps)}
    binding = dict(binding, **common)

    line = {
        "metric": "fd_waveform_likelihoods_per_s", "value": value, "unit": "walkers/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(B, args.workload),
        "e2e": {"value": e2e_value, "unit": "walkers/s", "h2d_bytes_per_step": pb.h2d_bytes(), "d2h_bytes_per_step": int(like_host.nbytes),
                "ms_per_step": ms_e2e_wall / args.steps,
                "call": "emrifd_loglike_batch_host (host packed sparse inputs -> H2D -> spline/segment/sum+likelihood -> D2H ll)"},
        "e2e_from_parameters": e2e_par,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": binding,
        "roofline_other_roof": other,
        "work": {"evals_per_walker": evals / B, "group_evals_per_walker": gevals / B, "mbe_per_walker": mbe / B, "modes_per_walker": pb.n_modes / B,
                 "knots_per_walker": pb.n_knots / B},
    }
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is an N = 1 figure (rank 0 would stall the other ranks' exit)
        try:
            from oracle import oracle as orc_mod
            orc_mod.build()
            rate, cores, done, el = cpu_oracle_rate(items, N, DT, data_w.cpu().numpy().view(np.complex128).reshape(2, n),
                                                    wfac.cpu().numpy(), max_seconds=15.0)
            line["cpu_baseline"] = {"value": rate, "unit": "walkers/s", "cores": cores, "kind": "port",
                                    "sample": f"first {done} walkers of the rank-0 batch, oracle f64 build (OpenMP over bins), {el:.1f} s"}
        except Exception as exc:   # the oracle is test infrastructure: never let it break the bench line
            line["cpu_baseline"] = {"value": None, "unit": "walkers/s", "cores": 0, "kind": "port", "sample": f"unavailable: {exc}"}
    guard.emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

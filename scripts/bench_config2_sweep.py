"""BASELINE.json configs[1]: the check_mode_by_mode.py sweep -- Tobs = 1 yr, dt = 10 s, eps = 1e-2, fixed inspiral
(p0 set so that the plunge is at 0.99 Tobs), nsteps = 10 prior draws with seed 2601996
(check_mode_by_mode.py:47-48,125-136,168-213).  For every point: FD waveform through the generator on the full grid and
on the 1 % grid f_arr = fftshift(fftfreq(int(0.01 N), dt)) (:233-241), the "check 1 ==" identity h = h+ - i hx (:247),
and -- instead of the TD comparison, which is another summation class -- accuracy against the CPU oracle:
  * per-bin error / max|h| against the binary128 oracle on every 64th bin,
  * mismatch 1 - <a|b>/sqrt(<a|a><b|b>) against the double (OpenMP) oracle on the full grid, LISA PSD (:299-306),
  * timing: full Python call, device part (CUDA events around the C-ABI call), double oracle on the host cores.
Prints one JSON line per point and a summary line.   python scripts/bench_config2_sweep.py [--nsteps 10]
"""
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
    ap.add_argument("--nsteps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from emri_frequencydomainwaveforms_b200 import engine, _lib
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
    from emri_frequencydomainwaveforms_b200.lisatools.diagnostic import inner_product
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
    from oracle.oracle import Oracle

    SEED = 2601996
    rng = np.random.RandomState(SEED)
    Tobs, dt, eps = 1.0, 10.0, 1e-2
    sum_kwargs = dict(pad_output=True, output_type="fd", odd_len=True)
    few_gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=sum_kwargs, use_gpu=True, return_list=False)
    few_gen_list = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=sum_kwargs, use_gpu=True, return_list=True)
    base = few_gen_list.waveform_generator
    oq, od = Oracle("quad"), Oracle("f64")
    h = _lib.get_handle()
    a = np.pi / 3
    emri_kwargs = dict(T=Tobs, dt=dt, eps=eps)
    done, failed, rows = 0, 0, []
    while done < args.nsteps and done + failed < 20 * args.nsteps:
        M = np.exp(rng.uniform(np.log(1e5), np.log(1e7)))
        mu = M * np.exp(rng.uniform(np.log(1e-6), np.log(1e-4)))
        e0 = rng.uniform(0.001, 0.7)
        Phi_phi0, Phi_r0 = rng.uniform(0, 2 * np.pi, 2)
        try:
            p0 = get_p_at_t(base.inspiral_generator, Tobs * 0.99, [M, mu, 0.0, e0, 1.0], xtol=2e-12)
            inj = np.array([M, mu, 0.0, p0, e0, 1.0, 1.0, a, a, a, a, Phi_phi0, 0.0, Phi_r0])
            few_gen(*inj, **emri_kwargs)                                  # warm-up of this shape
            torch.cuda.synchronize()
            tic = time.perf_counter(); hfd = few_gen(*inj, **emri_kwargs); torch.cuda.synchronize(); fd_time = time.perf_counter() - tic
            N = hfd.shape[0]
            f1 = np.fft.fftshift(np.fft.fftfreq(int(N * 0.01) | 1, dt))    # odd length (the generator needs one f = 0 bin)
            few_gen(*inj, f_arr=f1, **emri_kwargs); torch.cuda.synchronize()
            tic = time.perf_counter(); few_gen(*inj, f_arr=f1, **emri_kwargs); torch.cuda.synchronize(); fd_time_ds = time.perf_counter() - tic
            sig = few_gen_list(*inj, **emri_kwargs)
            check1 = complex((torch.conj(sig[0] - 1j * sig[1]) * hfd).sum() / (torch.conj(hfd) * hfd).sum())
        except ValueError:
            failed += 1
            continue
        # device part alone
        theta, phi, c2, s2 = few_gen_list._transform(a, a, a, a)
        it = base.prepare(M, mu, p0, e0, theta, phi, dist=1.0, Phi_phi0=Phi_phi0, Phi_r0=Phi_r0, **emri_kwargs)
        it["cos2psi"], it["sin2psi"] = c2, s2
        db = engine.DeviceBatch(engine.PackedBatch([it]), h)
        val = 1.0 / (N * dt)
        dev_ms = 1e30
        hp = hc = None
        for _ in range(4):   # first pass warms the caching allocator; best of the rest
            del hp, hc
            db.last_out = None
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record(); hp, hc, _ = engine.run_waveform(db, N, val); e1_.record(); torch.cuda.synchronize()
            dev_ms = min(dev_ms, e0_.elapsed_time(e1_))
        hp, hc = hp[0].cpu().numpy(), hc[0].cpu().numpy()
        zero, n = (N - 1) // 2, (N + 1) // 2
        # binary128 oracle on every 64th bin
        fpos = (np.arange(n, dtype=np.float64) * val)[::64]
        hq, cq, *_ = oq.fd_sum(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"], it["f_phi"], it["f_r"],
                               2 * len(fpos) - 1, 0.0, fpos, scale=it["scale"], cos2psi=c2, sin2psi=s2)
        sel = zero + 64 * np.arange(-(len(fpos) - 1), len(fpos))
        err = max(np.max(np.abs(hp[sel] - hq)), np.max(np.abs(hc[sel] - cq))) / max(np.max(np.abs(hp)), np.max(np.abs(hc)))
        # double oracle, full grid, timed (CPU baseline of this point) + mismatch with the LISA PSD
        tic = time.perf_counter()
        hd, cd, *_ = od.fd_sum(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"], it["f_phi"], it["f_r"],
                               N, val, scale=it["scale"], cos2psi=c2, sin2psi=s2, out_lo=zero, out_n=n)
        cpu_time = time.perf_counter() - tic
        f_pos = np.arange(n) * val
        psd = get_sensitivity(f_pos)
        ov = inner_product([hp[zero:], hc[zero:]], [hd, cd], f_arr=f_pos, PSD=psd, normalize=True)
        snr2 = inner_product([hp[zero:], hc[zero:]], [hp[zero:], hc[zero:]], f_arr=f_pos, PSD=psd)
        row = {"M": M, "mu": mu, "p0": p0, "e0": e0, "L": int(len(it["t"])), "K": int(len(it["m_arr"])), "N": int(N), "evals": int(od.last_n_eval),
               "fd_time_s": fd_time, "fd_time_1pct_grid_s": fd_time_ds, "device_ms": dev_ms, "oracle_f64_cpu_s": cpu_time,
               "check1": [check1.real, check1.imag], "per_bin_err_vs_binary128": float(err), "mismatch_vs_f64_oracle": float(1.0 - ov),
               "snr_dist1Gpc": float(np.sqrt(snr2))}
        rows.append(row)
        print(json.dumps(row))
        done += 1
    print(json.dumps({"summary": "configs[1] check_mode_by_mode-style sweep", "points": done, "failed_draws": failed,
                      "max_per_bin_err": max(r["per_bin_err_vs_binary128"] for r in rows),
                      "max_abs_mismatch": max(abs(r["mismatch_vs_f64_oracle"]) for r in rows),
                      "median_device_ms": float(np.median([r["device_ms"] for r in rows])),
                      "median_fd_call_s": float(np.median([r["fd_time_s"] for r in rows])),
                      "median_oracle_cpu_s": float(np.median([r["oracle_f64_cpu_s"] for r in rows])),
                      "cpu_threads": od.lib.orc_num_threads()}))


if __name__ == "__main__":
    main()

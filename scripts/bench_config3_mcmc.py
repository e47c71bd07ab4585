"""BASELINE.json configs[2]: emri_pe.py-style batched likelihood.  Tobs = 2 yr, M = 1e6, mu = 10, e0 = 0.35, p0 fixed so
the plunge is at 0.99 Tobs (emri_pe.py:620-636), eps = 1e-2, FD injection, downsample = 100 (emri_pe.py:322-391),
16 walkers drawn N(injection, cov/(2.4*6)) with seed 2601996 (emri_pe.py:65-66,440-444), evaluated through
Likelihood(parameter_transforms, fill_data_noise=True) -> FDTemplateModel.get_ll (the plugin hook).

Reports the time of one likelihood batch split into (a) the host-side producers that stay on the CPU per north_star
(trajectory ODE, amplitudes, Ylm, mode selection, packing) and (b) the accelerated path through the host-buffer C-ABI
call (H2D + spline + segmentation + fused mode-sum/likelihood + D2H).   python scripts/bench_config3_mcmc.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from emri_frequencydomainwaveforms_b200 import engine, _lib
    from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity, get_fd_waveform_fromFD
    from emri_frequencydomainwaveforms_b200.lisatools.diagnostic import snr
    from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel, Likelihood
    from emri_frequencydomainwaveforms_b200.utils.transform import TransformContainer
    from emri_frequencydomainwaveforms_b200.utils.utility import get_p_at_t
    from emri_frequencydomainwaveforms_b200.trajectory.inspiral import EMRIInspiral
    from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform

    SEED = 2601996
    np.random.seed(SEED)
    Tobs, dt, eps, nwalkers, downsample = 2.0, 10.0, 1e-2, 16, 100
    M, mu, e0 = 1e6, 10.0, 0.35
    qK = phiK = qS = phiS = np.pi / 3
    dist, Phi_phi0, Phi_r0 = 2.4539054256, np.pi / 3, np.pi / 3
    traj = EMRIInspiral(func="SchwarzEccFlux")
    p0 = get_p_at_t(traj, Tobs * 0.99, [M, mu, 0.0, e0, 1.0], xtol=2e-12)
    few_gen_list = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True),
                                        use_gpu=True, return_list=True)
    inj14 = np.array([M, mu, 0.1, p0, e0, 1.0, dist, qS, phiS, qK, phiK, Phi_phi0, 0.0, Phi_r0])
    emri_kwargs = dict(T=Tobs, dt=dt, eps=eps)
    # full-grid injection, non-zero mask, down-sampled symmetric f_arr (emri_pe.py:237-245,333-349)
    t0 = time.perf_counter()
    sig_fd = few_gen_list(*inj14, mask_positive=True, **emri_kwargs)
    torch.cuda.synchronize()
    t_full = time.perf_counter() - t0
    frequency = few_gen_list.waveform_generator.create_waveform.frequency
    fixed_freq = frequency[frequency >= 0.0].cpu().numpy()
    non_zero = (sig_fd[0].abs() > 1e-50).cpu().numpy()
    end_f = fixed_freq[non_zero].max()
    num = int(non_zero.sum() / downsample)
    p_freq = np.linspace(0.0, end_f * 1.01, num=num)
    newfreq = np.hstack((-p_freq[::-1][:-1], p_freq))
    f_arr_ds = newfreq[newfreq >= 0.0]
    emri_kwargs_ds = dict(emri_kwargs, f_arr=newfreq)
    like_gen_ds = get_fd_waveform_fromFD(few_gen_list, newfreq >= 0.0, dt)
    check = like_gen_ds(*inj14, **emri_kwargs_ds)
    snr_ds = float(snr(check, PSD=get_sensitivity(f_arr_ds), f_arr=f_arr_ds))
    # transforms exactly as emri_pe.py:161-206
    fill_dict = {"ndim_full": 14, "fill_values": np.array([0.0, 1.0, dist, qS, phiS, qK, phiK, 0.0]),
                 "fill_inds": np.array([2, 5, 6, 7, 8, 9, 10, 12])}
    tc = TransformContainer(fill_dict)
    model = FDTemplateModel(few_gen_list, f_arr=newfreq)
    like = Likelihood(model, 2, f_arr=f_arr_ds, parameter_transforms={"emri": tc}, fill_data_noise=True, subset=24)
    like.inject_signal(data_stream=check, noise_fn=[get_sensitivity, get_sensitivity], noise_kwargs=[{}, {}])
    inj6 = np.array([np.log(M), np.log(mu / M), p0, e0, Phi_phi0, Phi_r0])
    cov = np.load(os.path.join(ROOT, "emri_frequencydomainwaveforms_b200", "data", "walker_covariance.npy")) / (2.4 * 6)
    start = np.random.multivariate_normal(inj6, cov, size=nwalkers)
    ll_inj = like(inj6[None, :], **emri_kwargs)
    ll = like(start, **emri_kwargs)           # warm-up + result
    # timing: whole call, then the accelerated part alone on the packed batch of the same walkers
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        like(start, **emri_kwargs)
    t_call = (time.perf_counter() - t0) / reps
    # the same call with the NumPy host producers (amplitudes, Ylm, mode selection on the CPU)
    model_h = FDTemplateModel(few_gen_list, f_arr=newfreq, producers="host")
    like_h = Likelihood(model_h, 2, f_arr=f_arr_ds, parameter_transforms={"emri": tc}, fill_data_noise=True, subset=24)
    like_h.inject_signal(data_stream=check, noise_fn=[get_sensitivity, get_sensitivity], noise_kwargs=[{}, {}])
    ll_h = like_h(start, **emri_kwargs)
    t0 = time.perf_counter()
    like_h(start, **emri_kwargs)
    t_call_host = time.perf_counter() - t0
    p14 = tc.both_transforms(start)
    t0 = time.perf_counter()
    items, ok = model.prepare_batch(p14, **emri_kwargs)
    t_host = time.perf_counter() - t0
    pb = engine.PackedBatch(items)
    h = _lib.get_handle()
    N = len(newfreq)
    for _ in range(3):
        engine.run_loglike_host(pb, h, N, 0.0, model._fpos_dev)
    reps2 = 50
    t0 = time.perf_counter()
    for _ in range(reps2):
        out = engine.run_loglike_host(pb, h, N, 0.0, model._fpos_dev)
    t_gpu = (time.perf_counter() - t0) / reps2
    print(json.dumps({
        "config": "configs[2] emri_pe.py-style batched likelihood (Tobs=2 yr, downsample=100, 16 walkers)",
        "p0": p0, "N_full": int(frequency.shape[0]), "n_downsampled": int(len(f_arr_ds)), "snr_downsampled": snr_ds,
        "modes_injection": int(few_gen_list.waveform_generator.num_modes_kept), "walkers": nwalkers, "valid": int(ok.sum()),
        "ll_injection": float(ll_inj[0]), "ll_walkers_min_max": [float(np.nanmin(ll)), float(np.nanmax(ll))],
        "full_grid_waveform_s": t_full,
        "likelihood_call_s": t_call, "host_producers_s": t_host, "accelerated_path_s": t_gpu,
        "likelihoods_per_s_accelerated_path": nwalkers / t_gpu, "likelihoods_per_s_whole_call": nwalkers / t_call,
        "producers": model.producers, "likelihood_call_s_host_producers": t_call_host,
        "likelihoods_per_s_whole_call_host_producers": nwalkers / t_call_host,
        "max_abs_ll_diff_device_vs_host_producers": float(np.nanmax(np.abs(ll - ll_h))),
        "h2d_bytes": pb.h2d_bytes(), "gpu_launches_per_call": 5}))


if __name__ == "__main__":
    main()

"""Repeat the pipelined FDTemplateModel.get_ll(params) many times and report any exception / non-finite result."""
import os, sys, time, traceback, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import torch
import bench
from emri_frequencydomainwaveforms_b200 import _lib, engine
from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
from emri_frequencydomainwaveforms_b200.lisatools.likelihood import FDTemplateModel
h = _lib.get_handle(0); dev = h.torch_device
B = 64
N = bench.grid_len(); n = (N + 1) // 2; val = 1.0 / (N * bench.DT)
batches = bench.bench_batches(0, B)
dbi = engine.DeviceBatch(engine.PackedBatch(bench.draw_walkers(1, 1, bench.SEED)), h)
hp0, hc0, _ = engine.run_waveform(dbi, N, val, mask_positive=True)
f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
wf1 = torch.sqrt(torch.full((n,), val, dtype=torch.float64, device=dev) / get_sensitivity(f_pos))
wf = torch.stack([wf1, wf1]).contiguous(); dw = (torch.cat([hp0, hc0], dim=0) * wf).contiguous()
gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True), return_list=True, frame="source")
model = FDTemplateModel(gen); model.set_data(dw, wf)
raw = np.array([it["raw"] for items in batches for it in items])
P = np.zeros((len(raw), 14))
P[:, 0], P[:, 1], P[:, 3], P[:, 4], P[:, 5], P[:, 6] = raw[:, 0], raw[:, 1], raw[:, 2], raw[:, 3], 1.0, 1.0
P[:, 7], P[:, 8], P[:, 11], P[:, 13] = raw[:, 4], -np.pi / 2, raw[:, 5], raw[:, 6]
P = np.tile(P, (4, 1))
kw = dict(T=bench.T_YR, dt=bench.DT, eps=bench.EPS, N=N)
ref = model.get_ll(P, **kw)
bad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    try:
        ll = model.get_ll(P, **kw)
        if not np.array_equal(ll, ref):
            bad += 1; print("rep", rep, "differs: max", np.nanmax(np.abs(ll - ref)), "nan", int(np.isnan(ll).sum()), flush=True)
    except Exception:
        bad += 1; print("rep", rep, "EXCEPTION"); traceback.print_exc()
print("done, bad =", bad)

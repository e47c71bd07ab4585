"""cProfile of GenerateEMRIWaveform.__call__ for one 1-yr waveform (host-side overhead of the public API)."""
import cProfile, pstats, sys, os, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from emri_frequencydomainwaveforms_b200.waveform import GenerateEMRIWaveform
gen = GenerateEMRIWaveform("FastSchwarzschildEccentricFlux", sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True), return_list=True)
args = (1e6, 10.0, 0.0, 12.0, 0.35, 1.0, 1.0, np.pi / 3, np.pi / 4, np.pi / 3, np.pi / 4, 0.0, 0.0, 0.0)
kw = dict(T=1.0, dt=10.0, eps=1e-2, mask_positive=True)
for _ in range(5):
    out = gen(*args, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    out = gen(*args, **kw)
torch.cuda.synchronize()
print("ms per call", 1e3 * (time.perf_counter() - t0) / 50)
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    out = gen(*args, **kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

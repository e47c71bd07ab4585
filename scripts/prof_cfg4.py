"""One rank's configs[3] likelihood (4 yr, all 3843 modes) a few times: for ncu launch lists / captures.  python scripts/prof_cfg4.py [reps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from emri_frequencydomainwaveforms_b200 import _lib, engine
from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
h = _lib.get_handle(0)
dev = h.torch_device
it, T = bench.cfg4_walker()
N = bench.grid_len(T); n = (N + 1) // 2; val = 1.0 / (N * bench.DT)
db = engine.DeviceBatch(engine.PackedBatch([it]), h)
hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)
f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
wf1 = torch.sqrt(torch.full((n,), val, dtype=torch.float64, device=dev) / get_sensitivity(f_pos))
wf = torch.stack([wf1, wf1]).contiguous()
dw = (torch.cat([hp, hc], dim=0) * wf).contiguous()
h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), wf.data_ptr(), n))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = engine.run_loglike(db, N, val)
    torch.cuda.synchronize(); print("ms", 1e3 * (time.perf_counter() - t0), out.cpu().numpy())

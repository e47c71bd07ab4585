"""Stage-by-stage GPU-vs-oracle diagnostics (prints, never asserts).  Run on the GPU box."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import CASES, make_item, oracle_waveform, rel_err
from oracle.oracle import Oracle
from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
from emri_frequencydomainwaveforms_b200.summation.fdinterp import FDInterpolatedModeSum
from emri_frequencydomainwaveforms_b200 import _lib

print(torch.cuda.get_device_name(0))
gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
oq = Oracle("quad")
h = _lib.get_handle()
import ctypes as C
g = C.c_double()
h.check(h.lib.emrifd_bench_fp64_fma(h.h, 4096, C.byref(g)))
print("FP64 FMA peak GFLOP/s:", g.value)
for name in CASES:
    it = make_item(gen, name)
    t0 = time.time()
    hp_o, hc_o, coeff_o, br_o, nbr_o = oracle_waveform(oq, it)
    t1 = time.time()
    s = FDInterpolatedModeSum(pad_output=True, output_type="fd", odd_len=True)
    out = s(it["t"], it["teuk_modes"], it["ylms"], it["Phi_phi"], it["Phi_r"], it["m_arr"], it["n_arr"],
            it["M"], it["p"], it["e"], T=it["T"], dt=it["dt"], scale=it["scale"]).cpu().numpy()
    coeff_g = s.last_batch.coeff_host(0)
    br_g = s.last_batch.branches_host()
    print(f"== {name}: L={len(it['t'])} K={it['teuk_modes'].shape[1]} N={it['N']} evals={oq.last_n_eval} oracle {t1-t0:.2f}s")
    print("  coeff equal:", np.array_equal(coeff_g, coeff_o), "max abs diff", np.max(np.abs(coeff_g - coeff_o)),
          "n diff", int(np.sum(coeff_g != coeff_o)))
    for key in ("mode", "dir", "ja", "jb", "closed_end", "start", "end", "xa", "xb", "Fa", "Fb"):
        eq = np.array_equal(br_g[key], br_o[key])
        if not eq:
            bad = np.argwhere(br_g[key] != br_o[key])
            print("  branch field", key, "DIFFERS at", bad[:5].tolist(), br_g[key][tuple(bad[0])], br_o[key][tuple(bad[0])])
    print("  branches equal:", all(np.array_equal(br_g[k], br_o[k]) for k in br_o.dtype.names if k != "pad"))
    print("  hp rel err", rel_err(out[0], hp_o), "hc rel err", rel_err(out[1], hc_o),
          "support equal", np.array_equal(out[0] != 0, hp_o != 0), "nnz", int(np.sum(hp_o != 0)))
    if rel_err(out[0], hp_o) > 1e-10:
        d = np.abs(out[0] - hp_o)
        i = int(np.argmax(d))
        print("   worst bin", i, "f", s.frequency[i].item(), out[0][i], hp_o[i])
        print("   frac bins > 1e-10:", np.mean(d > 1e-10 * np.max(np.abs(hp_o))))

"""Debug statistics of the mode sum on the bench workload (needs a -DSUM_STATS build: EMRIFD_LIB=variants/stats.so)."""
import ctypes as C, json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from emri_frequencydomainwaveforms_b200 import _lib, engine
h = _lib.get_handle(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
wl = sys.argv[2] if len(sys.argv) > 2 else "plunge"
N = bench.grid_len(); n = (N + 1) // 2; val = 1.0 / (N * bench.DT)
items = bench.draw_walkers(1, B, bench.SEED, workload="cfg1") if wl == "cfg1" else bench.bench_batches(0, B)[0]
pb = engine.PackedBatch(items); db = engine.DeviceBatch(pb, h)
pb.walkers["out_off"] = np.arange(B, dtype=np.int64) * n
hp = torch.empty((B, n), dtype=torch.complex128, device="cuda"); hc = torch.empty_like(hp)
out = (C.c_uint64 * 16)()
h.lib.emrifd_debug_stats(out)
flags = _lib.INCLUDE_MINUS_M | _lib.MASK_POSITIVE
h.check(h.lib.emrifd_fd_waveform_batch(h.h, pb.walkers.ctypes.data, B, db.t.data_ptr(), db.teuk.data_ptr(), db.f_phi.data_ptr(), db.f_r.data_ptr(),
        db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(), db.m.data_ptr(), db.n.data_ptr(), db.ylm.data_ptr(), N, val, None, flags,
        db.coeff.data_ptr(), db.branches.data_ptr(), hp.data_ptr(), hc.data_ptr(), None))
h.lib.emrifd_debug_stats(out)
names = ["tiles", "passes", "sub_entries", "warp_sub_visits", "thread_pair_evals", "bins_accumulated", "robust_pair_solves", "empty_passes", "warp_sub_visits_with_work"]
st = {k: int(out[i]) for i, k in enumerate(names)}
st["group_evals"] = int(engine.group_evaluations(db).sum())
st["tiles_total"] = B * ((n + h.lib.emrifd_tile_bins() - 1) // h.lib.emrifd_tile_bins())
print(json.dumps(st))

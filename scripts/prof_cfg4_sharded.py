"""Stage timing of the cyclic bin-sharded configs[3] likelihood under torchrun (events on every rank, max over ranks)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from emri_frequencydomainwaveforms_b200 import _lib, engine, distributed as D
from emri_frequencydomainwaveforms_b200.fdutils import get_sensitivity
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = _lib.get_handle(lr); dev = h.torch_device
it, T = bench.cfg4_walker()
N = bench.grid_len(T); n = (N + 1) // 2; val = 1.0 / (N * bench.DT)
db = engine.DeviceBatch(engine.PackedBatch([it]), h); pb = db.pb
hp, hc, _ = engine.run_waveform(db, N, val, mask_positive=True)
f_pos = torch.arange(n, dtype=torch.float64, device=dev) * val
wf1 = torch.sqrt(torch.full((n,), val, dtype=torch.float64, device=dev) / get_sensitivity(f_pos))
wf = torch.stack([wf1, wf1]).contiguous(); dw = (torch.cat([hp, hc], dim=0) * wf).contiguous()
h.check(h.lib.emrifd_set_data(h.h, dw.data_ptr(), wf.data_ptr(), n))
flags = _lib.INCLUDE_MINUS_M | _lib.MASK_POSITIVE
out = torch.zeros((1, 3), dtype=torch.float64, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
def stage_a():
    h.check(h.lib.emrifd_batch_spline(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.teuk.data_ptr(), db.f_phi.data_ptr(), db.f_r.data_ptr(), db.Phi_phi.data_ptr(), db.Phi_r.data_ptr(), db.coeff.data_ptr()))
    h.check(h.lib.emrifd_batch_segment(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(), db.n.data_ptr(), int(N), float(val), None, db.branches.data_ptr(), None))
def stage_b():
    h.check(h.lib.emrifd_batch_sum_cyclic(h.h, pb.walkers.ctypes.data, pb.B, db.t.data_ptr(), db.coeff.data_ptr(), db.m.data_ptr(), db.n.data_ptr(), db.ylm.data_ptr(), db.branches.data_ptr(), int(N), float(val), None, flags, rank, world, None, None, out.data_ptr()))
res = []
for rep in range(8):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter()
    e[0].record(); stage_a(); e[1].record(); stage_b(); e[2].record(); dist.all_reduce(out, op=dist.ReduceOp.SUM); e[3].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    h.check(h.lib.emrifd_sum_kernel_time(h.h, 1, None, None))
    res.append([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[0].elapsed_time(e[3]), 1e3 * (t1 - t0)])
r = torch.tensor(res[3:], dtype=torch.float64, device=dev).median(dim=0).values
dist.all_reduce(r, op=dist.ReduceOp.MAX)
import ctypes as C
stage_b(); kms, kmain, kl = C.c_double(), C.c_double(), C.c_int64()
h.check(h.lib.emrifd_sum_kernel_times(h.h, 0, C.byref(kms), C.byref(kmain), C.byref(kl)))
km = torch.tensor([kms.value / max(kl.value, 1), kmain.value / max(kl.value, 1)], dtype=torch.float64, device=dev)
dist.all_reduce(km, op=dist.ReduceOp.MAX)
if rank == 0:
    print("world", world, "ms: spline+segment %.3f | groups+pieces+sum+finalize %.3f | all_reduce %.3f | total %.3f | host enqueue %.3f | pair kernel %.3f mode_sum %.3f" % (*r.tolist(), *km.tolist()))
dist.destroy_process_group()

"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_evals]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
for V in rows[2:]:
    d = dict(zip(H, V))
    g = lambda k: float(d[k].replace(",", "")) if k in d and d[k] not in ("", "n/a") else float("nan")
    print("kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    cyc = g("sm__cycles_elapsed.avg")
    print(f"  duration {g('gpu__time_duration.sum'):.3f} {U[H.index('gpu__time_duration.sum')]}  regs {d.get('launch__registers_per_thread')}  "
          f"occupancy(warps active %) {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}")
    print(f"  dram read {g('dram__bytes_read.sum'):.4g} {U[H.index('dram__bytes_read.sum')]} write {g('dram__bytes_write.sum'):.4g} {U[H.index('dram__bytes_write.sum')]}  "
          f"dram% {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.2f}")
    print(f"  fp64 pipe active % {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f}  issue active % {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}")
    dfma, dmul, dadd = (g(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") for k in ("dfma", "dmul", "dadd"))
    inst = g("smsp__inst_executed.sum"); tpi = g("smsp__thread_inst_executed_per_inst_executed.ratio")
    print(f"  per cycle (chip): dfma {dfma:.0f} dmul {dmul:.0f} dadd {dadd:.0f} -> {2*dfma+dmul+dadd:.0f} flop/cycle of {148*64*2} peak "
          f"({(2*dfma+dmul+dadd)/(148*128)*100:.1f}%)   fp64 share of thread-inst {(dfma+dmul+dadd)*cyc/(inst*tpi)*100:.1f}%")
    if len(sys.argv) > 2:
        ne = float(sys.argv[2])
        print(f"  per evaluation: {(2*dfma+dmul+dadd)*cyc/ne:.1f} flop, {inst*tpi/ne:.1f} thread-instructions")
    for k in H:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            v = g(k)
            if v > 0.15:
                print(f"    stall {k.split('stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")

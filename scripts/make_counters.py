"""profiles/r2_counters.json + profiles/r2_modesum_ncu_summary.txt from a full ncu capture of mode_sum_kernel.
usage: python scripts/make_counters.py gpurun_out/X_modesum.ncu-rep gpurun_out/X_ncu_full.log
The bench line in the log (bench.py --batch 16 --steps 2 --warmup 3 --no-extras under ncu -s 4 -c 1) tells how many stationary points
the captured launch solved: launches 0..3 are the injection and the three warm-up steps, so the capture is batch index 3."""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, log = sys.argv[1], sys.argv[2]
line = [l for l in open(log) if l.startswith('{"metric"')][-1]
work = json.loads(line)["work"]
solves = work["solves_per_batch"][3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, V = rows[0], rows[2]
d = dict(zip(H, V))
g = lambda k: float(d[k].replace(",", ""))
cyc = g("sm__cycles_elapsed.avg")
dfma, dmul, dadd = (g(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc for k in ("dfma", "dmul", "dadd"))
inst = g("smsp__inst_executed.sum") * g("smsp__thread_inst_executed_per_inst_executed.ratio")
old = json.load(open(os.path.join(ROOT, "profiles", "r2_counters.json")))
out = {
    "_doc": "Hardware-counter figures of mode_sum_kernel<true,true,2,true> that bench.py's roofline uses, from the committed ncu --set full capture; "
            "and the flop count of the reference formulation's inner loop (oracle).",
    "capture": f"profiles/r2_modesum_ncu_summary.txt (ncu --set full --clock-control none -k regex:mode_sum_kernel -s 4 -c 1, bench.py --batch 16 --steps 2 "
               f"--warmup 3 --no-extras; the captured launch is rotating batch 3: {solves} stationary points solved)",
    "kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size"),
    "duration_us": g("gpu__time_duration.sum") / (1e3 if H and rows[1][H.index("gpu__time_duration.sum")] in ("nsecond", "ns") else 1.0),
    "solves_in_capture": solves,
    "executed_flops_per_solve": round((2 * dfma + dmul + dadd) / solves, 1),
    "fp64_thread_instructions_per_solve": round((dfma + dmul + dadd) / solves, 1),
    "thread_instructions_per_solve": round(inst / solves, 1),
    "fp64_pipe_active": round(g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100, 3),
    "issue_active": round(g("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100, 3),
    "fp64_share_of_thread_instructions": round((dfma + dmul + dadd) / inst, 3),
    "registers_per_thread": int(g("launch__registers_per_thread")),
    "dram_bytes_read": g("dram__bytes_read.sum") * (1e6 if rows[1][H.index("dram__bytes_read.sum")] == "Mbyte" else 1.0),
    "dram_bytes_write": g("dram__bytes_write.sum") * (1e6 if rows[1][H.index("dram__bytes_write.sum")] == "Mbyte" else 1.0),
    "oracle_flops_per_mode_eval": old["oracle_flops_per_mode_eval"],
    "oracle_flops_itemisation": old["oracle_flops_itemisation"],
}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_counters.json"), "w"), indent=1)
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, str(solves)], capture_output=True, text=True).stdout
lines = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines2.py"), rep, "40"], capture_output=True, text=True).stdout
with open(os.path.join(ROOT, "profiles", "r2_modesum_ncu_summary.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:mode_sum_kernel -s 4 -c 1 python bench.py --batch 16 --steps 2 --warmup 3 "
            f"--no-cpu-baseline --no-extras\n# captured launch: rotating batch 3, {solves} stationary points solved ('per evaluation' = per solved stationary point)\n")
    f.write(summ + "\n# stall samples per source line (emrifd.cu), top 40\n" + "\n".join(l[:230] for l in lines.splitlines()) + "\n")
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", "r2_modesum_ncu_details.txt"), "w").write(det)
print(json.dumps({k: v for k, v in out.items() if k not in ("oracle_flops_itemisation", "_doc")}, indent=1))

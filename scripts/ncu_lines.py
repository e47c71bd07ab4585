"""Rank CUDA source lines / SASS opcodes of an .ncu-rep by instructions and stall samples.
usage: python scripts/ncu_lines.py rep.ncu-rep n_evals [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, nev = sys.argv[1], float(sys.argv[2]) / 32
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fl = lambda x: float(x) if x not in ("", "-", "...") else 0.0
out, ops, tot_s, tot_i = [], defaultdict(float), 0.0, 0.0
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        s, n = fl(r[4]), fl(r[7])
        out.append((int(r[0]), r[1].strip()[:100], s, n)); tot_s += s; tot_i += n
    elif len(r) > 7 and r[2].startswith("0x"):
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3].strip())
        if m: ops[m.group(2)] += fl(r[7])
print("total warp-inst per warp-eval", round(tot_i / nev, 1))
for ln, src, s, n in sorted(out, key=lambda x: -x[3])[:top]:
    print(f"{ln:5d} {s/tot_s*100:5.1f}% smp {n/tot_i*100:5.1f}% inst ({n/nev:6.1f}/eval) | {src}")
print({k: round(v / nev, 1) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:26]})

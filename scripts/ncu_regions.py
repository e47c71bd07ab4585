"""Instruction mix per source-line region.  usage: python scripts/ncu_regions.py rep n_evals name:lo-hi ..."""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, nev = sys.argv[1], float(sys.argv[2]) / 32
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fl = lambda x: float(x) if x not in ("", "-", "...") else 0.0
cur, per, smp = None, defaultdict(lambda: defaultdict(float)), defaultdict(float)
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        cur = int(r[0]); smp[cur] += fl(r[4])
    elif len(r) > 7 and r[2].startswith("0x") and cur:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3].strip())
        if m: per[cur][m.group(2)] += fl(r[7]) / nev
ts = sum(smp.values())
for spec in sys.argv[3:]:
    name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
    tot = defaultdict(float)
    for ln in per:
        if lo <= ln <= hi:
            for k, v in per[ln].items(): tot[k] += v
    s = sum(v for ln, v in smp.items() if lo <= ln <= hi)
    print(f"{name:18s} {sum(tot.values()):6.1f} inst/eval  {s/ts*100:5.1f}% samples ", {k: round(v, 1) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:12]})

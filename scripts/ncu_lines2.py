"""Per-source-line stall samples of an ncu capture.  usage: python scripts/ncu_lines2.py rep [top]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[2]
ix = {k: i for i, k in enumerate(H)}
stall_cols = [k for k in H if k.startswith("stall_") and "Not Issued" not in k]
fl = lambda x: float(x) if x not in ("", "-", "...") else 0.0
lines = {}
cur = None
per = defaultdict(lambda: defaultdict(float))
inst = defaultdict(float)
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        cur = int(r[0]); lines[cur] = r[1]
    elif len(r) > 7 and r[2].startswith("0x") and cur:
        for k in stall_cols:
            per[cur][k] += fl(r[ix[k]])
        inst[cur] += fl(r[ix["Instructions Executed"]])
tot = sum(sum(v.values()) for v in per.values())
bytype = defaultdict(float)
for v in per.values():
    for k, x in v.items(): bytype[k] += x
print("total samples", tot, {k: round(100 * v / tot, 1) for k, v in sorted(bytype.items(), key=lambda x: -x[1])[:10]})
for ln, v in sorted(per.items(), key=lambda x: -sum(x[1].values()))[:top]:
    s = sum(v.values())
    print(f"{ln:5d} {100*s/tot:5.1f}% inst {inst[ln]:10.0f} ", {k[6:]: round(100 * x / tot, 1) for k, x in sorted(v.items(), key=lambda x: -x[1])[:4] if x > 0}, "|", lines[ln].strip()[:90])

"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum).  usage: python scripts/launch_shares.py launches.csv [skip_first_n]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        H = r; st = i + 1; break
ix = {k: i for i, k in enumerate(H)}
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[st:]:
    if len(r) < len(H) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[ix['Kernel Name']].split('(')[0][:70]
    tot[name] += float(r[ix['Metric Value']].replace(',', '')) / 1e3; cnt[name] += 1
T = sum(tot.values())
ours = {k: v for k, v in tot.items() if 'at::' not in k and 'nccl' not in k.lower()}
To = sum(ours.values())
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:25]:
    print(f"{v:10.1f} us {cnt[k]:4d}x  avg {v/cnt[k]:8.1f} us  {100*v/To:5.1f}% of own  {k}")
print("total", round(T, 1), "us; own kernels", round(To, 1), "us")

"""profiles/traffic.json (bench.py's roofline.traffic) from ncu launch lists.
usage: python scripts/make_traffic.py plunge=gpurun_out/launches_X.csv cfg1=gpurun_out/launches_X_cfg1.csv"""
import csv, json, os, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MULT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out = {"_doc": "dram__bytes_read.sum + dram__bytes_write.sum of the mode-sum launch pair (empty_tile_kernel<true,true> + mode_sum_kernel<true,true>) "
               "averaged over the launches of one ncu launch-list pass of bench.py --batch 64 (cold-cache, serialised); bench.py scales by batch"}
for spec in sys.argv[1:]:
    wl, path = spec.split("=")
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            H = r; st = i + 1; break
    ix = {k: i for i, k in enumerate(H)}
    agg = defaultdict(lambda: defaultdict(float)); cnt = defaultdict(int)
    for r in rows[st:]:
        if len(r) < len(H):
            continue
        n = r[ix['Kernel Name']].split('(')[0]
        if not (n.startswith('void mode_sum_kernel<1, 1') or n.startswith('void empty_tile_kernel<1, 1')):
            continue
        m = r[ix['Metric Name']]
        if m.startswith('dram__bytes'):
            agg[n][m] += float(r[ix['Metric Value']].replace(',', '')) * MULT.get(r[ix['Metric Unit']], 1)
        elif m == 'gpu__time_duration.sum':
            cnt[n] += 1
    tot, src = 0.0, []
    for n in sorted(agg):
        rd, wr = agg[n]['dram__bytes_read.sum'] / cnt[n], agg[n]['dram__bytes_write.sum'] / cnt[n]
        tot += rd + wr
        src.append(f"{n}: {rd/1e6:.1f} MB read + {wr/1e6:.1f} MB write per launch ({cnt[n]} launches)")
    out[wl] = {"batch": 64, "dram_bytes_per_launch": tot, "source": f"profiles/{os.path.basename(path)}: " + "; ".join(src)}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

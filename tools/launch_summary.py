"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list (per kernel and per launch)."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith('==')]))
def ns(row):
    return float(row['Metric Value'].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6}[row['Metric Unit']]
tot = sum(ns(r) for r in rows)
agg = collections.OrderedDict()
for r in rows:
    key = re.sub(r'\(.*', '', r['Kernel Name']).replace('void mstr::', '').replace('mstr::', '')
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += ns(r)
print("launches %d, total %.3f ms (cold-cache, serialised: compare shares)" % (len(rows), tot / 1e6))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %3d %9.3f ms %5.1f%%" % (k[:58], n, t / 1e6, 100 * t / tot))
if "--all" in sys.argv:
    for i, r in enumerate(rows):
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void mstr::', '')[:52]
        print("%3d %-52s grid %-14s %9.1f us" % (i, name, r['Grid Size'], ns(r) / 1e3))

"""Per-kernel totals of the SECOND half of an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of a script that runs the
same call twice.  usage: python profiles/launch_summary.py list.csv "title" """
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
data = [(r[ki], float(r[vi].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1}.get(r[ui], 1e-6)) for r in rows[1:] if len(r) > vi]
half = data[len(data) // 2:]
agg = collections.OrderedDict()
for k, v in half:
    k = k.split('(')[0][:78]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in half)
print(f"{sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}: {tot:.3f} ms of kernel time over {len(half)} launches (ncu-serialised, cold caches; second of two identical calls)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {v:8.3f} ms  {100 * v / tot:5.1f}%  x{n:<3d} {k}")

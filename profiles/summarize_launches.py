"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel for ONE step.
usage: python profiles/summarize_launches.py launches.csv [step_index]   (steps are delimited by plane_stats_kernel)"""
import collections
import csv
import re
import sys


def main(path, step=-2):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            v, u = float(row['Metric Value'].replace(',', '')), row['Metric Unit']
            ms = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
            rows.append((re.sub(r'\(.*', '', row['Kernel Name'])[:90], ms))
    marks = [i for i, (k, _) in enumerate(rows) if 'plane_stats' in k]
    a, b = marks[step], marks[step + 1]
    agg = collections.OrderedDict()
    for k, ms in rows[a:b]:
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"one step = launches {a}..{b - 1} of {path}: {b - a} launches, {total:.3f} ms of kernel time (ncu-serialised, cold caches)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.3f} ms {100 * v[1] / total:5.1f}%  x{v[0]:<3d} {k}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -2)

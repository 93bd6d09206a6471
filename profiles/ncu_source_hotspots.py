"""Top source lines of a kernel by executed warp instructions and stall samples.
usage: python profiles/ncu_source_hotspots.py report.ncu-rep kernel-regex [n]"""
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name', f'regex:{kernel}',
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fpath, hdr, agg, idx = None, None, [], {}
    for r in rows:
        if not r:
            continue
        if r[0] in ('File Path', 'File Name'):
            fpath = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            hdr = r
        elif hdr and r[0].isdigit() and len(r) == len(hdr):
            inst = float(r[hdr.index('Instructions Executed')] or 0)
            smp = float(r[hdr.index('# Samples')] or 0)
            if inst or smp:
                key = (fpath, int(r[0]))
                if key not in idx:
                    idx[key] = len(agg)
                    agg.append([0.0, 0.0, fpath, int(r[0]), r[1].strip()[:110]])
                agg[idx[key]][0] += inst
                agg[idx[key]][1] += smp
    ti, ts = sum(a[0] for a in agg), sum(a[1] for a in agg)
    print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}")
    for inst, smp, f, ln, src in sorted(agg, reverse=True)[:top]:
        print(f"{100 * inst / ti:5.1f}% inst {100 * smp / max(ts, 1):5.1f}% smp  {f}:{ln:<4d} {src}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)

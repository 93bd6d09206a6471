"""Stall samples of a warp-specialised kernel split by role: the SASS of the first matching launch is cut at its USETMAXREG
instructions (every role starts with setmaxnreg) and, per region, the stall reasons and the hottest instructions are listed.
usage: python profiles/ncu_sass_roles.py report.ncu-rep kernel-regex [top-n]"""
import csv
import subprocess
import sys


def main(path, kernel, top=12):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name', f'regex:{kernel}',
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, body = None, []
    for r in rows:
        if r and r[0] == 'Address':
            if hdr is not None:
                break
            hdr = r
        elif hdr and len(r) == len(hdr):
            body.append(r)
    i_s, i_i, i_src = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    total = sum(float(r[i_s] or 0) for r in body)
    cuts = [0] + [k for k, r in enumerate(body) if 'USETMAXREG' in r[i_src]] + [len(body)]
    print(f"{len(body)} instructions, {total:.0f} samples, regions cut at USETMAXREG: {cuts}")
    for a, b in zip(cuts[:-1], cuts[1:]):
        reg = body[a:b]
        smp = sum(float(r[i_s] or 0) for r in reg)
        if smp < 0.005 * total:
            continue
        ins = sum(float(r[i_i] or 0) for r in reg)
        by = sorted(((sum(float(r[i] or 0) for r in reg), h) for i, h in reasons), reverse=True)[:6]
        ops = {}
        for r in reg:
            op = r[i_src].split()[0] if not r[i_src].startswith('@') else r[i_src].split()[1]
            ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + float(r[i_i] or 0)
        print(f"\n== instructions {a}..{b}: {100 * smp / total:.1f}% of samples, {ins:.3g} warp instructions; "
              + ", ".join(f"{h[6:]} {100 * v / max(smp, 1):.0f}%" for v, h in by))
        print("   executed: " + ", ".join(f"{k} {v:.3g}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:10]))
        for r in sorted(reg, key=lambda r: -float(r[i_s] or 0))[:top]:
            why = max(reasons, key=lambda ih: float(r[ih[0]] or 0))[1][6:]
            print(f"   {100 * float(r[i_s] or 0) / total:5.2f}%  {why:14s} {r[i_src][:100]}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 12)

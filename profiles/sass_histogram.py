"""SASS opcode histogram per kernel of libnfe_b200.so (and of the L2 gather micro-benchmark): the evidence for what the
kernels are built from — UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
LDGSTS (cp.async), UTMALDG (TMA; only the micro-benchmark that retired gather4 has it), generic LD/ST vs LDS/STS.

    python profiles/sass_histogram.py > profiles/sass_r02_opcodes.txt        (build container; needs cuobjdump)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGETS = [os.path.join(ROOT, "nerffaceediting_b200", "lib", "libnfe_b200.so"), os.path.join(ROOT, "profiles", "microbench", "_bin", "l2_gather")]
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "USETMAXREG", "LDG", "STG", "RED", "REDG", "ATOMG", "LDS", "STS", "LD", "ST",
       "MUFU", "F2FP", "FFMA2", "FFMA", "HMMA", "NANOSLEEP", "BAR", "LDL", "STL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    for path in TARGETS:
        if not os.path.exists(path):
            continue
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        print(f"==== {os.path.relpath(path, ROOT)}   (arch lines: {sorted(set(re.findall(r'arch = (sm_\w+)', out)))})")
        kernels = collections.OrderedDict()
        cur = None
        for line in out.splitlines():
            m = re.match(r"\s+Function : (\S+)", line)
            if m:
                cur = kernels.setdefault(m.group(1), collections.Counter())
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and cur is not None:
                cur[m.group(1)] += 1
        total = collections.Counter()
        for name, c in kernels.items():
            total.update(c)
            n = sum(c.values())
            short = demangle(name)
            short = re.sub(r"\(.*", "", short)[:110]
            keys = " ".join(f"{k}={c[k]}" for k in KEY if c[k])
            print(f"{short:110s} {n:6d} instr  {keys}")
        print(f"TOTAL {' '.join(f'{k}={total[k]}' for k in KEY if total[k])}")
        print()


if __name__ == "__main__":
    sys.exit(main())

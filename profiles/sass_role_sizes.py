"""Instruction count of every role of field_pipe2_kernel (regions between USETMAXREG instructions) — the kernel's instruction-cache
footprint matters: 4 roles share each sub-partition's L0.  usage: python profiles/sass_role_sizes.py [object-or-library] [kernel-substring]"""
import re
import subprocess
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "nerffaceediting_b200/lib/libnfe_b200.so"
want = sys.argv[2] if len(sys.argv) > 2 else "field_pipe2_kernelILi1ELb1"
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, ins = None, []
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur and want in cur:
        ins.append(m.group(1))
cuts = [0] + [i for i, s in enumerate(ins) if "USETMAXREG" in s] + [len(ins)]
print(f"{want}: {len(ins)} instructions ({len(ins) * 16 / 1024:.0f} KB); regions {[b - a for a, b in zip(cuts[:-1], cuts[1:])]}")

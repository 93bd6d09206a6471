"""Cycles per phase of field_bwd_kernel's tile loop (debug build: NFE_NVCC_FLAGS=-DNFE_BWD_PROFILE).  Run on the GPU box:
   NFE_NVCC_FLAGS=-DNFE_BWD_PROFILE python -m nerffaceediting_b200.build --force && python profiles/bwd_phase_profile.py"""
import ctypes
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from nerffaceediting_b200 import _lib  # noqa: E402

lib = _lib.load()
fn = lib.nfe_debug_bwd_profile
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
import bench  # noqa: E402,F401  (same process: run the c4 workload through bench.main)

sys.argv = ["bench.py", "--workload", "c4", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"] + sys.argv[1:]
buf = (ctypes.c_ulonglong * 16)()
bench.main()
fn(buf, 1)
names = ["gather", "G1 wait", "epilogue 1", "G2+G4 wait", "epilogue 2", "G3+G5 wait", "epilogue 3", "scatter", "", "loop top"]
tot = sum(buf[i] for i in range(16))
for i, nm in enumerate(names):
    if nm:
        print(f"{nm:12s} {100.0 * buf[i] / max(tot, 1):5.1f}%   {buf[i] / 1e6:10.1f} Mcycles")

"""Where each role of field_pipe_kernel waits (debug build: NFE_NVCC_FLAGS=-DNFE_PIPE_PROFILE).  Run on the GPU box:
   NFE_NVCC_FLAGS="-DNFE_PIPE_PROFILE -DNFE_GATHER_WARPS=8" python -m nerffaceediting_b200.build --force && python profiles/pipe_role_profile.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from nerffaceediting_b200 import _lib, triplane  # noqa: E402
from nerffaceediting_b200.ray_sampler import RaySampler  # noqa: E402
from nerffaceediting_b200.renderer import DisentangledImportanceRenderer  # noqa: E402

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["c2"]
raw_host, dec, c2w, k, opts = bench.make_inputs(torch, wl, dev, 1000)
opts["nfe_precision"] = "bf16x3"
mods = {"sampler": RaySampler(), "normalize_plane": triplane.normalize_plane, "renderer": DisentangledImportanceRenderer()}
raw, dec, c2w, k = raw_host.to(dev), dec.to(dev), c2w.to(dev), k.to(dev)
lib = _lib.load()
fn = lib.nfe_debug_pipe_profile
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 48)()
with torch.no_grad():
    for _ in range(3):
        bench.hot_path_step(torch, mods, raw, dec, c2w, k, wl["res"], opts)
    fn(buf, 1)
    steps = 5
    for _ in range(steps):
        bench.hot_path_step(torch, mods, raw, dec, c2w, k, wl["res"], opts)
    fn(buf, 1)
names = ["empty", "tmem_free", "full", "a2_full", "d1_full", "d2a_full", "d2b_full"]
for r, role in enumerate(("gather", "mma", "epilogue")):
    tot = buf[r * 16 + 15]
    print(f"{role:9s} total warp-cycles {tot:.3e}: " + ", ".join(f"{n} {100.0 * buf[r * 16 + i] / max(tot, 1):.1f}%" for i, n in enumerate(names) if buf[r * 16 + i]))

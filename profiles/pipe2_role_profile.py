"""Where each role of field_pipe2_kernel waits (debug build: -DNFE_PIPE_PROFILE).  Run on the GPU box, e.g. from
profiles/run_r02_pipe2_variants.sh.  Prints, per role, the share of its warps' lifetime spent blocked on each barrier."""
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
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
raw_host, dec, c2w, k, opts = bench.make_inputs(torch, wl, dev, 1000)
opts["nfe_precision"] = "bf16x3"
mods = {"sampler": RaySampler(), "normalize_plane": triplane.normalize_plane, "renderer": DisentangledImportanceRenderer()}
raw, dec, c2w, k = raw_host.to(dev), dec.to(dev), c2w.to(dev), k.to(dev)
lib = _lib.load()
fn = lib.nfe_debug_pipe2_profile
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 64)()
with torch.no_grad():
    for _ in range(3):
        bench.hot_path_step(torch, mods, raw, dec, c2w, k, wl["res"], opts)
    fn(buf, 1)
    for _ in range(5):
        bench.hot_path_step(torch, mods, raw, dec, c2w, k, wl["res"], opts)
    fn(buf, 1)
names = ["taps_full", "empty", "taps_empty", "full", "a2a_full", "d2_free", "a2b_full", "d2x_full(own)", "d1_full", "d2a_full", "d2b_full"]
for r, role in enumerate(("gather", "tap", "mma", "epilogue")):
    tot = buf[r * 16 + 15]
    print(f"{role:9s} total warp-cycles {tot:.3e}: " + ", ".join(f"{n} {100.0 * buf[r * 16 + i] / max(tot, 1):.1f}%" for i, n in enumerate(names) if buf[r * 16 + i]))

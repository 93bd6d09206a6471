"""Where the roles of conv_gemm_kernel wait (debug build: NFE_NVCC_FLAGS=-DNFE_MC_PROFILE).  One SynthesisLayer at full size.
usage: python profiles/modconv_role_profile.py [in_ch out_ch res up fp16|fp32 batch]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import synth_inputs as synth  # noqa: E402
from nerffaceediting_b200 import _lib, networks as net  # noqa: E402

i, o, res, up = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (256, 256, 256, 1)))
dtype = torch.float32 if len(sys.argv) > 5 and sys.argv[5] == "fp32" else torch.float16
n = int(sys.argv[6]) if len(sys.argv) > 6 else 8
layer = synth.fill_module(net.SynthesisLayer(i, o, w_dim=512, resolution=res, up=up, conv_clamp=256), 11).cuda().eval()
x = torch.randn(n, i, res // up, res // up, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
w = torch.randn(n, 512, device="cuda")
lib = _lib.load()
try:
    fn = lib.nfe_debug_modconv_profile
    fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
except AttributeError:                       # production build: timing only
    def fn(buf, reset):
        return 0
buf = (ctypes.c_ulonglong * 16)()
with torch.no_grad():
    for _ in range(3):
        layer(x, w, noise_mode='const')
    fn(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        layer(x, w, noise_mode='const')
    e1.record()
    fn(buf, 1)
ctas = max(buf[9], 1)
print(f"{i}->{o} @ {res}^2 up={up} {dtype} batch {n}: {e0.elapsed_time(e1) / 5:.4f} ms per layer call, {ctas // 5} CTAs per call")
print(f"  per CTA (cycles): total {buf[8] / ctas:.0f}, setup {buf[10] / ctas:.0f}")
print(f"  mma thread:  loop {buf[2] / ctas:.0f}, waits a_full {buf[0] / ctas:.0f}, b_full {buf[1] / ctas:.0f}")
print(f"  loader t0:   loading {buf[4] / ctas:.0f}, waits a_empty {buf[3] / ctas:.0f}")
print(f"  epilogue t0: waits acc_full {buf[5] / ctas:.0f}, epilogue {buf[6] / ctas:.0f}")
print(f"  b producer:  waits b_empty {buf[7] / ctas:.0f}")
print(f"  epilogue t0, fast loop: tcgen05.wait::ld + next ld {buf[15] / ctas:.0f}, first group of a pair {buf[14] / ctas:.0f}, second (with its wait) {buf[13] / ctas:.0f}, "
      f"copy-out {buf[11] / ctas:.0f}")

"""SR pre-resize (f1): nfe_resize_bilinear vs torch F.interpolate on the c2 feature image [8,32,64,64] -> 128^2.  GPU box only."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from nerffaceediting_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for shape, size in (((8, 32, 64, 64), 128), ((8, 32, 256, 256), 128)):
    x = torch.randn(*shape, device=dev)

    def t(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    ours = t(lambda: ops.resize_bilinear(x, size, True))
    # the reference resizes x and rgb = x[:, :3] separately (superresolution.py:283-286)
    ref = t(lambda: (F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False, antialias=True),
                     F.interpolate(x[:, :3], size=(size, size), mode="bilinear", align_corners=False, antialias=True)))
    byt = (x.numel() + x.shape[0] * x.shape[1] * size * size) * 4
    print(f"{shape} -> {size}^2: nfe_resize_bilinear {ours:.1f} us ({byt / ours / 1e3:.0f} GB/s read+write), torch interpolate x2 {ref:.1f} us")

"""CUDA-graph capture of a fixed-shape hot-path call.

At small ray counts the render is launch-bound: BASELINE configs[0] (batch 1, 64^2 rays) spends 0.20 ms in
kernels out of a 0.32 ms step, the rest being the Python/ctypes issue cost of its ~13 launches.  Every entry
point of libnfe_b200.so only enqueues work on the caller's stream (no allocation, no synchronisation,
include/nfe_b200.h), so the whole call sequence  RaySampler -> normalize_plane -> renderer.forward  is
capturable as one CUDA graph and replayed with a single launch.

    step = graphs.capture(lambda: renderer(norm, raw, decoder, o, d, opts))
    rgb, seg, depth, wsum = step()          # replays; the outputs are the SAME tensors every time

Contract (the usual one for CUDA graphs):
  * the tensors the callable closes over are *static*: feed new inputs by writing into them in place
    (`raw.copy_(new_planes)`), never by rebinding; shapes, options and the decoder's parameter tensors are
    frozen at capture time (parameter *values* may change, their addresses may not);
  * outputs are overwritten by the next replay — copy what must survive;
  * deterministic sampling only (`rendering_options['nfe_deterministic']=True`, or cameras/jitter supplied
    from outside): the Philox seed/offset of stochastic mode is a by-value launch argument and would be
    frozen into the graph, repeating the same jitter on every replay.  `capture` refuses a stochastic render;
  * inference only (capture runs under torch.no_grad());
  * warm-up and capture run in a fresh epoch of the plane registries (plane_registry.new_epoch): a staging buffer or
    statistics recorded BEFORE the capture (e.g. by a normalize_plane outside the captured call, or kept by
    `nfe_cache_planes`) is not visible inside, so the captured step re-stages from its static inputs and a replay after
    `raw.copy_(new_planes)` renders the new planes.  Whatever the step registers itself (normalize_plane inside the
    captured call) still hits, and is captured with it.

The reference has no counterpart (eager ATen ops, training/volumetric_rendering/renderer.py:88-140).
"""
import torch

from . import _lib, ops
from . import plane_registry as registry


class GraphedCall:
    """A captured call: `self()` replays the graph on the current stream and returns the static outputs."""

    def __init__(self, graph, outputs, kernels):
        self.graph = graph
        self.outputs = outputs
        self.kernels = kernels          # launches of libnfe_b200.so kernels recorded in the graph (per replay)

    def __call__(self):
        self.graph.replay()
        return self.outputs


def capture(fn, warmup=2, pool=None):
    """Run `fn()` `warmup` times eagerly on a side stream (first-use work such as cudaFuncSetAttribute and torch's
    allocator growth must not happen inside the capture), then capture one more call into a CUDA graph."""
    if not torch.cuda.is_available():
        raise RuntimeError("graphs.capture: needs a CUDA device (this path has no CPU fallback)")
    philox0 = ops.philox_draws()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with registry.new_epoch():
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if ops.philox_draws() != philox0:
            raise RuntimeError("graphs.capture: the call draws random numbers (stochastic sampling or density noise); its Philox "
                               "seed/offset would be frozen into the graph — capture a deterministic render "
                               "(rendering_options['nfe_deterministic']=True, density_noise=0)")
    graph = torch.cuda.CUDAGraph()
    launches0 = _lib.launch_count()
    with registry.new_epoch():
        with torch.no_grad(), torch.cuda.graph(graph, pool=pool):
            outputs = fn()
    kernels = _lib.launch_count() - launches0
    if kernels <= 0:
        raise RuntimeError("graphs.capture: the call launched no kernel of libnfe_b200.so")
    return GraphedCall(graph, outputs, kernels)

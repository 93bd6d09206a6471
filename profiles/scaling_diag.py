"""Where does the N>1 step time go?  Run under torchrun (2+ ranks) on the GPU box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/scaling_diag.py
Prints, per variant, the max-over-ranks ms per step: render only / + NCCL all-gather in stream / + all-gather overlapped on
a side stream / the all-gather alone / host time per step (launch overhead)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from nerffaceediting_b200 import triplane  # noqa: E402
from nerffaceediting_b200.ray_sampler import RaySampler  # noqa: E402
from nerffaceediting_b200.renderer import DisentangledImportanceRenderer  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS["c2"]
raw_host, dec, c2w, k, opts = bench.make_inputs(torch, wl, dev, 1000 + rank)
opts["nfe_precision"] = "bf16x3"
mods = {"sampler": RaySampler(), "normalize_plane": triplane.normalize_plane, "renderer": DisentangledImportanceRenderer()}
raw, dec, c2w, k = raw_host.to(dev), dec.to(dev), c2w.to(dev), k.to(dev)
n, res = wl["batch"], wl["res"]
packed = [torch.empty((n, res * res, 49), device=dev) for _ in range(2)]
gathered = [torch.empty((world * n, res * res, 49), device=dev) for _ in range(2)]
comm = torch.cuda.Stream(device=dev)
ready = [torch.cuda.Event() for _ in range(2)]
comm_done = [torch.cuda.Event() for _ in range(2)]
state = {"i": 0}


def render(slot):
    rgb, seg, depth, wsum = bench.hot_path_step(torch, mods, raw, dec, c2w, k, res, opts)
    torch.cat([rgb, seg, depth, wsum], dim=-1, out=packed[slot])


def v_render():
    render(0)


def v_inline():
    render(0)
    if world > 1:
        dist.all_gather_into_tensor(gathered[0], packed[0])


def v_overlap():
    i = state["i"]; slot = i & 1; state["i"] = i + 1
    main = torch.cuda.current_stream()
    main.wait_event(comm_done[slot])
    render(slot)
    ready[slot].record(main)
    if world > 1:
        with torch.cuda.stream(comm):
            comm.wait_event(ready[slot])
            dist.all_gather_into_tensor(gathered[slot], packed[slot])
            comm_done[slot].record(comm)


def v_gather_only():
    if world > 1:
        dist.all_gather_into_tensor(gathered[0], packed[0])


def timed(fn, steps=30, warmup=5):
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        host = (time.perf_counter() - t0) / steps * 1e3
        torch.cuda.current_stream().wait_stream(comm)
        e1.record()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps, host], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


for name, fn in (("render only", v_render), ("render + inline NCCL all-gather", v_inline), ("render + overlapped all-gather", v_overlap),
                 ("all-gather alone", v_gather_only), ("render only (again)", v_render)):
    ms, host = timed(fn)
    if rank == 0:
        print(f"{name:34s} {ms:7.3f} ms/step   host {host:6.3f} ms/step   world {world}", flush=True)
if world > 1:
    dist.destroy_process_group()

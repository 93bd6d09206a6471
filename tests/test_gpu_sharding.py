"""sharding.render_sharded / ShardedRenderer with REAL kernels: the deferred depth clamp,
nfe_finish_depth, the MIN/MAX reduction of the depth range and the pack / pad / all-gather / unpack plumbing on the device.

* world of one: the sharded code path (force_sharded_path) == the plain render, bit for bit;
* world of two: two processes — NCCL on two GPUs when the box has them, otherwise gloo with both ranks on cuda:0 (the
  kernels and the sharding code are the same; only the transport differs) — batch split, ray-block split, a shared plane
  set, stochastic sampling (where the depth range really differs per shard) and the overlapped ShardedRenderer, each
  against the single-rank render of the same inputs computed in the same process.
"""
import os
import socket

import numpy as np
import pytest
import torch

import synth_inputs as synth

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene(dev, n, plane_batch, hw, res, seed=3):
    from nerffaceediting_b200 import triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.triplane import DisentangledOSGDecoder
    raw = torch.from_numpy(synth.hash_normal(seed, (plane_batch, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3)).to(dev)
    torch.manual_seed(seed)
    dec = DisentangledOSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'decoder_seg_dim': 15}).to(dev)
    with torch.no_grad():
        for p in dec.parameters():
            if p.dim() == 1:
                p.copy_(torch.randn_like(p) * 0.5)
    c2w, k = synth.camera_sweep(n)
    with torch.no_grad():
        o, d = RaySampler()(c2w.to(dev), k.to(dev), res)
        norm, _, _ = triplane.normalize_plane(raw)
    return norm.view(plane_batch, 3, 32, hw, hw), raw.view(plane_batch, 3, 32, hw, hw), dec, o, d


CASES = {
    # name: (batch, plane batch, image side, deterministic)
    "batch_split": (4, 4, 12, True),
    "batch_split_ragged": (3, 3, 10, True),
    "ray_blocks": (1, 1, 13, True),
    "shared_planes": (3, 1, 10, True),
    "ray_blocks_stochastic": (1, 1, 16, False),
    "batch_split_stochastic": (2, 2, 12, False),
}


def _opts(deterministic):
    return dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=16, depth_resolution_importance=16, nfe_deterministic=deterministic)


def _worker(rank, world, port, backend, results):
    import torch.distributed as dist
    from nerffaceediting_b200 import sharding
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    out = {}
    try:
        renderer = DisentangledImportanceRenderer()
        for name, (n, pb, res, det) in CASES.items():
            norm, raw, dec, o, d = _scene(dev, n, pb, 32, res)
            opts = _opts(det)
            with torch.no_grad():
                torch.manual_seed(100)
                got = sharding.render_sharded(renderer, norm, raw, dec, o, d, opts)
                if det:
                    want = renderer(norm, raw, dec, o, d, opts)
                    out[name] = max(float((a - b).abs().max()) for a, b in zip(got, want))
                else:
                    # stochastic shards draw other jitter than the whole image would: check the contract instead — every rank
                    # holds the same maps, and the depth map respects ONE global range (the all-reduced one)
                    flat = torch.cat([t.reshape(-1) for t in got])
                    ref = flat.clone()
                    dist.broadcast(ref, 0)
                    out[name] = float((flat - ref).abs().max())
                    assert torch.isfinite(flat).all()
                    assert float(got[2].min()) >= opts['ray_start'] - 1e-6 and float(got[2].max()) <= opts['ray_end'] + 0.1
        # weak-scaling form: each rank renders ITS OWN items, maps gathered on a communication stream
        n, res = 2, 12
        norm, raw, dec, o, d = _scene(dev, n, n, 32, res, seed=10 + rank)
        opts = _opts(True)
        sr = sharding.ShardedRenderer(renderer, overlap=True)
        with torch.no_grad():
            pend = [sr(norm, raw, dec, o, d, opts, local_batch=True) for _ in range(3)]       # ring of 2: the third reuses slot 0
            rgb, seg, depth, wsum = pend[-1].wait()
            sr.drain()
            mine = renderer(norm, raw, dec, o, d, opts)
        assert rgb.shape == (world * n, res * res, 32) and seg.shape == (world * n, res * res, 15)
        sl = slice(rank * n, (rank + 1) * n)
        out["local_batch"] = max(float((a[sl] - b).abs().max()) for a, b in zip((rgb, seg, depth, wsum), mine))
        # the other rank's block is that rank's own render
        other = torch.cat([t.reshape(-1) for t in (rgb, seg, depth, wsum)])
        ref = other.clone()
        dist.broadcast(ref, 0)
        out["local_batch_same_everywhere"] = float((other - ref).abs().max())
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_render_sharded_two_ranks_real_kernels():
    import torch.multiprocessing as mp
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, results)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(600) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for rank in range(2):
        for name, err in results[rank].items():
            assert err == 0.0, (backend, rank, name, err)


def test_render_sharded_world_of_one_is_the_plain_render():
    import torch.distributed as dist
    from nerffaceediting_b200 import sharding
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer, ImportanceRenderer
    from nerffaceediting_b200.triplane import OSGDecoder
    dev = torch.device("cuda:0")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        renderer = DisentangledImportanceRenderer()
        for det in (True, False):
            norm, raw, dec, o, d = _scene(dev, 2, 2, 32, 12)
            opts = _opts(det)
            with torch.no_grad():
                torch.manual_seed(7)
                want = renderer(norm, raw, dec, o, d, opts)
                torch.manual_seed(7)
                got = sharding.render_sharded(renderer, norm, raw, dec, o, d, opts, force_sharded_path=True)
            for a, b in zip(got, want):
                assert torch.equal(a, b)
        # OSG renderer (no seg map) through the same path
        torch.manual_seed(1)
        osg = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).to(dev)
        with torch.no_grad():
            want = ImportanceRenderer()(raw, osg, o, d, _opts(True))
            got = sharding.render_sharded(ImportanceRenderer(), None, raw, osg, o, d, _opts(True), force_sharded_path=True)
        assert got[1] is None and all(torch.equal(a, b) for a, b in zip((got[0], got[2], got[3]), want))
    finally:
        dist.destroy_process_group()

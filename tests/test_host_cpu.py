"""CPU-side tests: the C-ABI library loads and exports what include/nfe_b200.h declares, the host
logic (decoder recognition, option handling, shadow package) behaves, and nothing falls back to a CPU
compute path.  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from nerffaceediting_b200 import _lib, ops, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("NFE_REFERENCE", "/root/reference")


@pytest.fixture(scope="module")
def built():
    from nerffaceediting_b200 import build
    return build.build()


def header_functions():
    text = open(os.path.join(ROOT, "include", "nfe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nfe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    names = header_functions()
    assert len(names) >= 20
    lib = ctypes.CDLL(built)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/nfe_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().nfe_version() == 1
    assert isinstance(_lib.launch_count(), int)


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header():
    # nfe_mlp: 4 pointers, 3 ints, 4 floats; nfe_render_cfg ends with precision after two uint64
    assert ctypes.sizeof(_lib.NfeMlp) == 4 * 8 + 3 * 4 + 4 * 4 + 4   # trailing pad to 8
    assert _lib.NfeRenderCfg.seed.offset % 8 == 0 and _lib.NfeRenderCfg.precision.offset == _lib.NfeRenderCfg.offset.offset + 8


def test_no_cpu_fallback():
    from nerffaceediting_b200 import triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import ImportanceRenderer, sample_from_planes
    with pytest.raises(RuntimeError, match="CUDA"):
        triplane.normalize_plane(torch.randn(1, 96, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        RaySampler()(torch.eye(4)[None], torch.eye(3)[None], 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        sample_from_planes(None, torch.randn(1, 3, 32, 8, 8), torch.zeros(1, 2, 3), box_warp=1)
    dec = triplane.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    with pytest.raises(RuntimeError, match="CUDA"), torch.no_grad():
        ImportanceRenderer()(torch.randn(1, 3, 32, 8, 8), dec, torch.zeros(1, 2, 3), torch.ones(1, 2, 3), synth.FFHQ_RENDERING_OPTIONS)


def test_graph_capture_needs_cuda():
    from nerffaceediting_b200 import graphs
    if torch.cuda.is_available():
        pytest.skip("CPU-container check")
    with pytest.raises(RuntimeError, match="CUDA"):
        graphs.capture(lambda: None)


def test_bench_workloads_cover_every_baseline_config():
    """bench.py names one workload per BASELINE.json config (c1..c5) and the reference arm runs the oracle port on them."""
    import json
    import bench
    configs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert sorted(bench.WORKLOADS) == [f"c{i + 1}" for i in range(len(configs))]
    assert bench.WORKLOADS["c5"]["plane_batch"] == 1 and bench.WORKLOADS["c5"]["res"] == 256 and bench.WORKLOADS["c5"]["s_c"] == 96
    raw, dec, c2w, k, opts = bench.make_inputs(torch, dict(bench.WORKLOADS["c5"], batch=2), torch.device("cpu"), 0)
    assert raw.shape == (1, 96, 256, 256) and c2w.shape == (2, 4, 4) and opts["depth_resolution_importance"] == 96


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): ONE JSON line on stdout with the contract's keys,
    timed on the oracle port with the host's threads, no GPU needed."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "rendered rays/sec (48+48 samples)" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 of a torchrun launch exit without work or output
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_plane_registries_key_on_version_and_tolerate_inference_tensors():
    """The staging / provenance registries key on (address, shape, version): a 4-D tensor and its 5-D view share a key, an
    in-place write invalidates it, and tensors made under torch.inference_mode() (no version counter) simply never hit."""
    from nerffaceediting_b200 import ops
    norm, raw = torch.randn(2, 96, 4, 4), torch.randn(2, 96, 4, 4)
    assert ops._key5(norm) == ops._key5(norm.view(2, 3, 32, 4, 4))
    ops._provenance_put(raw, ops._key5(norm), torch.ones(2, 96), torch.zeros(2, 96))
    hit = ops.provenance(norm.view(2, 3, 32, 4, 4), raw.view(2, 3, 32, 4, 4))
    assert hit is not None and hit[0].shape == (2, 96)
    raw.add_(1.0)                                               # stale: the version moved on
    assert ops.provenance(norm, raw) is None
    with torch.inference_mode():
        n2, r2 = torch.randn(1, 96, 4, 4), torch.randn(1, 96, 4, 4)
        before = len(ops._PROVENANCE)
        ops._provenance_put(r2, ops._key5(n2), torch.ones(1, 96), torch.zeros(1, 96))
        assert len(ops._PROVENANCE) == before and ops.provenance(n2, r2) is None
        ops._cache_put(ops._key5(n2), n2, n2)
        assert ops._key5(n2) not in ops._CL_CACHE
    ops._PROVENANCE.clear()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnfe_b200.so")
    with pytest.raises(RuntimeError, match="not built"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nerffaceediting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the CPU reference is the oracle", "").replace("(SURVEY.md §7.5, oracle nfo_resample_ray)", "") \
                    .replace("the oracle does not have", ""), f"{f} mentions the oracle"


def test_decoder_recognition_by_structure():
    from nerffaceediting_b200 import triplane
    o = {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'decoder_seg_dim': 15}
    assert ops.describe_decoder(triplane.OSGDecoder(32, o))[0] == ops.DEC_OSG
    assert ops.describe_decoder(triplane.DisentangledOSGDecoder(32, o))[0] == ops.DEC_DISENTANGLED
    assert ops.describe_decoder(triplane.SegmentationOSGDecoder(32, o))[0] == ops.DEC_SEGMENTATION
    assert ops.describe_decoder(torch.nn.Linear(3, 3)) is None
    odd = triplane.OSGDecoder(32, dict(o, decoder_output_dim=16))
    assert ops.describe_decoder(odd) is None                                   # unsupported width -> staged path
    relu = triplane.OSGDecoder(32, o)
    relu.net[1] = torch.nn.ReLU()
    assert ops.describe_decoder(relu) is None
    # parameter names are the reference's (state dicts are interchangeable)
    assert sorted(triplane.DisentangledOSGDecoder(32, o).state_dict()) == sorted(
        f"{n}.{i}.{p}" for n in ("geo_net", "app_net") for i in (0, 2) for p in ("weight", "bias"))
    fc = triplane.FullyConnectedLayer(32, 64, lr_multiplier=0.5)
    assert abs(fc.weight_gain - 0.5 / np.sqrt(32)) < 1e-12 and fc.bias_gain == 0.5


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "training")), reason="reference checkout not present")
def test_reference_decoders_are_recognised_and_shadow_package_resolves():
    code = r"""
import sys, pickle
import training.triplane as t
from training.volumetric_rendering import renderer, ray_marcher, ray_sampler, math_utils
from nerffaceediting_b200 import ops
assert renderer.__file__.startswith(sys.argv[1]), renderer.__file__
assert t.__file__.startswith(sys.argv[2]), t.__file__
assert t.DisentangledImportanceRenderer is renderer.DisentangledImportanceRenderer
assert t.RaySampler is ray_sampler.RaySampler
o = {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'decoder_seg_dim': 15}
assert ops.describe_decoder(t.OSGDecoder(32, o))[0] == ops.DEC_OSG
assert ops.describe_decoder(t.DisentangledOSGDecoder(32, o))[0] == ops.DEC_DISENTANGLED
assert ops.describe_decoder(t.SegmentationOSGDecoder(32, o))[0] == ops.DEC_SEGMENTATION
r = pickle.loads(pickle.dumps(renderer.DisentangledImportanceRenderer()))
assert type(r).__module__ == 'training.volumetric_rendering.renderer'
assert type(r.ray_marcher).__module__ == 'training.volumetric_rendering.ray_marcher'
assert hasattr(math_utils, 'get_ray_limits_box') and hasattr(renderer, 'sample_from_3dgrid')
print('ok')
"""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "shadow"), REFERENCE]))
    r = subprocess.run([sys.executable, "-c", code, os.path.join(ROOT, "shadow"), REFERENCE], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_synthetic_inputs():
    a = synth.hash_normal(3, (4, 5))
    assert a.dtype == np.float32 and np.array_equal(a, synth.hash_normal(3, (4, 5))) and not np.array_equal(a, synth.hash_normal(4, (4, 5)))
    big = synth.hash_normal(1, (200000,))
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1) < 0.01
    c2w, k = synth.camera_sweep(8)
    assert c2w.shape == (8, 4, 4) and k.shape == (8, 3, 3)
    assert torch.allclose(c2w[:, :3, 3].norm(dim=-1), torch.full((8,), 2.7), atol=1e-5)
    assert abs(float(k[0, 0, 0]) - 4.2634) < 1e-3          # FOV 18.837 deg with the reference's 3.14159 / 1.414


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "training")), reason="reference checkout not present")
def test_synth_cameras_match_reference_camera_utils():
    code = r"""
import sys, math, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import camera_utils
from nerffaceediting_b200 import synth
for h, v in ((math.pi/2, math.pi/2), (math.pi/2 - 0.4, math.pi/2 + 0.25)):
    ref = camera_utils.LookAtPoseSampler.sample(h, v, torch.tensor([0, 0, 0.2]), radius=2.7)
    assert torch.equal(ref, synth.look_at_cam2world([h], [v]))
assert torch.equal(camera_utils.FOV_to_intrinsics(18.837), synth.fov_to_intrinsics(18.837))
print('ok')
"""
    r = subprocess.run([sys.executable, "-c", code, ROOT, REFERENCE], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]

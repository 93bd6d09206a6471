"""Synthetic inputs for the hot path: planes, cameras and rendering options.

Checkpoints and datasets are unavailable offline, so benches and parity tests run on
random-init planes/decoders and on the camera schedule the reference's scripts use
(gen_samples.py:156-172, gen_videos.py:126-130).  Nothing here touches the GPU kernels; it is test and bench infrastructure, not part of the product package.
"""
import math

import numpy as np
import torch

# ffhq block of the reference's rendering options (train.py:305-313) + the keys the renderer reads
# (renderer.py:91-100,116,143,146; ray_marcher.py:32,52).
FFHQ_RENDERING_OPTIONS = {
    'depth_resolution': 48,
    'depth_resolution_importance': 48,
    'ray_start': 2.25,
    'ray_end': 3.3,
    'box_warp': 1,
    'disparity_space_sampling': False,
    'clamp_mode': 'softplus',
    'white_back': False,
}


def hash_normal(seed, shape, offset=0):
    """Bit-reproducible ~N(0,1) floats: splitmix64 of the element index, four 16-bit uniforms summed
    (Irwin-Hall) and rescaled.  Integer arithmetic plus exact double ops only, so the golden-vector
    generator (run beside the reference) and the tests (run anywhere) regenerate identical planes
    instead of committing 25 MB fixtures.  `offset` is the flat index of the first element, so that a large
    tensor can be regenerated slice by slice: hash_normal(s, (a+b, ...))[a:] == hash_normal(s, (b, ...), offset=a*...)."""
    n = int(np.prod(shape))
    with np.errstate(over='ignore'):
        z = np.arange(offset, offset + n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    m = np.uint64(0xFFFF)
    s = (z & m) + ((z >> np.uint64(16)) & m) + ((z >> np.uint64(32)) & m) + (z >> np.uint64(48))
    x = (s.astype(np.float64) / 65536.0 - 2.0) * math.sqrt(3.0)
    return x.astype(np.float32).reshape(shape)


def _normalize(v):
    return v / torch.norm(v, dim=-1, keepdim=True)


def look_at_cam2world(horizontal, vertical, lookat=(0.0, 0.0, 0.2), radius=2.7, device='cpu'):
    """cam2world [B,4,4] of a camera on a sphere looking at `lookat`; the zero-stddev case of
    LookAtPoseSampler.sample + create_cam2world_matrix (camera_utils.py:69-86,118-137).
    `horizontal`/`vertical` are sequences of angles in radians (pi/2, pi/2 = frontal)."""
    h = torch.as_tensor(horizontal, dtype=torch.float32, device=device).reshape(-1, 1)
    v = torch.as_tensor(vertical, dtype=torch.float32, device=device).reshape(-1, 1)
    v = torch.clamp(v, 1e-5, math.pi - 1e-5)
    theta = h
    phi = torch.arccos(1 - 2 * (v / math.pi))
    b = h.shape[0]
    origin = torch.zeros((b, 3), device=device)
    origin[:, 0:1] = radius * torch.sin(phi) * torch.cos(math.pi - theta)
    origin[:, 2:3] = radius * torch.sin(phi) * torch.sin(math.pi - theta)
    origin[:, 1:2] = radius * torch.cos(phi)
    forward = _normalize(torch.as_tensor(lookat, dtype=torch.float32, device=device) - origin)
    up = torch.tensor([0.0, 1.0, 0.0], device=device).expand_as(forward)
    right = -_normalize(torch.cross(up, forward, dim=-1))
    up = _normalize(torch.cross(forward, right, dim=-1))
    rot = torch.eye(4, device=device).unsqueeze(0).repeat(b, 1, 1)
    rot[:, :3, :3] = torch.stack((right, up, forward), dim=-1)
    trans = torch.eye(4, device=device).unsqueeze(0).repeat(b, 1, 1)
    trans[:, :3, 3] = origin
    return trans @ rot


def fov_to_intrinsics(fov_degrees=18.837, device='cpu'):
    """Normalised 3x3 intrinsics (camera_utils.py:140-149; note the reference's 3.14159 and 1.414)."""
    focal = float(1 / (math.tan(fov_degrees * 3.14159 / 360) * 1.414))
    return torch.tensor([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], device=device)


def camera_sweep(batch, yaw_range=0.4, pitch_range=0.25, device='cpu'):
    """(cam2world [B,4,4], intrinsics [B,3,3]): yaw/pitch evenly spaced over the batch around the
    frontal pose, as SURVEY.md §8(d) specifies for the bench workloads."""
    if batch == 1:
        yaw, pitch = [0.0], [0.0]
    else:
        yaw = [-yaw_range + 2 * yaw_range * i / (batch - 1) for i in range(batch)]
        pitch = [-pitch_range + 2 * pitch_range * i / (batch - 1) for i in range(batch)]
    c2w = look_at_cam2world([math.pi / 2 + y for y in yaw], [math.pi / 2 + p for p in pitch], device=device)
    k = fov_to_intrinsics(device=device).unsqueeze(0).repeat(batch, 1, 1)
    return c2w, k


def fill_module(module, seed):
    """Deterministic parameters / buffers for a backbone or super-resolution module (the reference's or this package's: the names are
    the same), so that goldens and tests build identical networks without storing weights: weights, constants and noise maps
    ~N(0,1) as at initialisation; style-affine biases around 1, other biases and the noise strengths small but non-zero."""
    import torch
    with torch.no_grad():
        for i, (name, t) in enumerate(sorted(module.state_dict().items())):
            if name.endswith('resample_filter'):
                continue
            h = torch.from_numpy(hash_normal(seed + 7 * i, tuple(t.shape) if t.ndim else (1,))).reshape(t.shape).to(t.dtype)
            if name.endswith('noise_strength'):
                h = 0.3 * h
            elif name.endswith('affine.bias'):
                h = 1.0 + 0.2 * h
            elif name.endswith('bias') or name.endswith('w_avg'):
                h = 0.2 * h
            t.copy_(h)
    return module

"""Shared helpers for the parity tests (golden loading, error metrics, decoder rebuilds)."""
import os

import numpy as np

from oracle import nfe_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b):
    """max |a-b| / max|b|: the 'relative' of BASELINE.json's tolerances, scaled by the reference
    tensor's magnitude so that exact zeros do not blow the ratio up."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def elem_err(a, b, floor=1e-2):
    """Element-wise relative error with an absolute floor: max over elements of |a-b| / max(|b|, floor * max|b|).
    `rel_err` (max-norm) is the loosest reading of BASELINE.json's "1e-4 relative"; this is the strict one, and the floor
    says from which magnitude on an element is held to it (elements below floor * max|b| are held to floor * max|b| * tol
    absolute).  VERDICT r01 weak #1a."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.max(np.abs(b)), 1e-30)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * scale)))


def oracle_decoder(arrays, prefix, kind, lr=1.0):
    """(kind_id, net_a, net_b, color_dim, seg_dim) from saved state_dict arrays `<prefix>.<net>.<i>.<p>`."""
    def mlp(net):
        w1, b1 = arrays[f"{prefix}.{net}.0.weight"], arrays[f"{prefix}.{net}.0.bias"]
        w2, b2 = arrays[f"{prefix}.{net}.2.weight"], arrays[f"{prefix}.{net}.2.bias"]
        return orc.Mlp(w1, b1, w2, b2, lr / np.sqrt(w1.shape[1]), lr, lr / np.sqrt(w2.shape[1]), lr)
    if kind == "osg":
        a = mlp("net")
        return orc.DEC_OSG, a, None, a.out_dim - 1, 0
    if kind == "dis":
        a, b = mlp("geo_net"), mlp("app_net")
        return orc.DEC_DISENTANGLED, a, b, b.out_dim, a.out_dim - 1
    a, b = mlp("net"), mlp("seg_net")
    return orc.DEC_SEGMENTATION, a, b, a.out_dim - 1, b.out_dim

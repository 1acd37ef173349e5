"""Voxel-grid mean downsampling: CPU oracle (test infrastructure).

Restates Open3D >= 0.18 ``t.PointCloud.voxel_down_sample(voxel_size)`` with
``reduction="mean"`` (called at ``pp.py:509-512``) - PARITY UNPINNED
(Open3D is not vendored/installable; <= 0.17 returned voxel corners instead of centroids).

Choices the reference leaves open and this oracle fixes:
  * voxel index   = ``floor(float32(x) / float32(voxel_size))`` (a true IEEE divide), int64;
  * output order  = first-occurrence order: voxel v is the v-th distinct voxel met when the
                    points are read in input order (Open3D's order is hash-buffer order,
                    i.e. arbitrary);
  * membership    = ``p2v[i]`` = output row of the voxel that holds input point i;
  * centroid      = two definitions, both provided:
      - ``centroids_o3d``   : float32 serial ``index_add`` in input order then ``sum/count``
                              in float32 - Open3D's CPU arithmetic.  The GPU result must be
                              within 1e-5 (relative) of this.
      - ``centroids_fixed`` : order-independent fixed-point mean - every coordinate is
                              rounded to a multiple of 2^-24 m (``rint(x * 2^24)``, int64),
                              summed exactly, divided once in float64 and rounded to
                              float32.  Deterministic for any accumulation order, which is
                              what the CUDA kernel's integer atomics compute; the GPU result
                              must match this bit for bit.  (Attributes use 2^-20.)
"""
from __future__ import annotations

import numpy as np

POS_SCALE = float(1 << 24)
ATTR_SCALE = float(1 << 20)
KEY_HALF = 1 << 20            # voxel indices must lie in [-2^20, 2^20)
POS_LIMIT = 65536.0           # |coordinate| must be < 2^16 m for the fixed-point sums
ATTR_LIMIT = float(1 << 20)


def voxel_index(pos: np.ndarray, voxel_size: float) -> np.ndarray:
    vs = np.float32(voxel_size)
    return np.floor(pos.astype(np.float32) / vs).astype(np.int64)


def pack_key(idx: np.ndarray) -> np.ndarray:
    """63-bit packed key ``((ix+2^20)<<42) | ((iy+2^20)<<21) | (iz+2^20)``."""
    u = (idx + KEY_HALF).astype(np.uint64)
    return (u[:, 0] << np.uint64(42)) | (u[:, 1] << np.uint64(21)) | u[:, 2]


def in_domain(pos: np.ndarray, voxel_size: float) -> np.ndarray:
    """Points the voxel stage accepts: finite, |x| < 2^16 m, voxel index within +-2^20."""
    with np.errstate(invalid="ignore", over="ignore"):
        ok = np.isfinite(pos).all(axis=1) & (np.abs(pos) < POS_LIMIT).all(axis=1)
        q = np.floor(pos.astype(np.float32) / np.float32(voxel_size))
        ok &= ((q >= -KEY_HALF) & (q < KEY_HALF)).all(axis=1)
    return ok


def membership(pos: np.ndarray, voxel_size: float):
    """``(p2v, first_idx)``: voxel row of every point and the first point of every voxel."""
    keys = pack_key(voxel_index(pos, voxel_size))
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")            # sorted-key rank -> first-occurrence rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[np.asarray(inv).reshape(-1)].astype(np.int32), first[order].astype(np.int64)


def centroids_o3d(vals: np.ndarray, p2v: np.ndarray, n_vox: int) -> np.ndarray:
    """float32 serial index_add in input order, then sum / count (Open3D CPU arithmetic)."""
    vals = vals.astype(np.float32)
    vals2 = vals.reshape(vals.shape[0], -1)
    acc = np.zeros((n_vox, vals2.shape[1]), dtype=np.float32)
    np.add.at(acc, p2v, vals2)                          # unbuffered, in input order
    cnt = np.zeros(n_vox, dtype=np.float32)
    np.add.at(cnt, p2v, np.float32(1.0))
    out = acc / cnt[:, None]
    return out.reshape((n_vox,) + vals.shape[1:])


def centroids_fixed(vals: np.ndarray, p2v: np.ndarray, n_vox: int, scale: float = POS_SCALE) -> np.ndarray:
    """Order-independent fixed-point mean (the definition the CUDA kernel implements)."""
    vals2 = vals.astype(np.float32).reshape(vals.shape[0], -1)
    q = np.rint(vals2.astype(np.float64) * scale).astype(np.int64)
    acc = np.zeros((n_vox, vals2.shape[1]), dtype=np.int64)
    np.add.at(acc, p2v, q)
    cnt = np.bincount(p2v, minlength=n_vox).astype(np.float64)
    out = ((acc.astype(np.float64) / cnt[:, None]) * (1.0 / scale)).astype(np.float32)
    return out.reshape((n_vox,) + vals.shape[1:])


def voxel_down_sample(pos: np.ndarray, voxel_size: float, intensity=None, fixed=True):
    """Returns ``dict(positions, intensity, p2v, counts, first_idx)``."""
    p2v, first = membership(pos, voxel_size)
    nv = first.size
    mean = centroids_fixed if fixed else centroids_o3d
    out = {"positions": mean(pos, p2v, nv) if not fixed else centroids_fixed(pos, p2v, nv, POS_SCALE),
           "p2v": p2v, "counts": np.bincount(p2v, minlength=nv).astype(np.uint32), "first_idx": first}
    if intensity is not None:
        out["intensity"] = (centroids_fixed(intensity, p2v, nv, ATTR_SCALE) if fixed
                            else centroids_o3d(intensity, p2v, nv))
    return out

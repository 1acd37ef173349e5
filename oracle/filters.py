"""Non-finite filter, rigid transform, ROI crop: CPU oracle (test infrastructure).

Open3D (``isl-org/Open3D``, version unpinned by the reference; v0.18/0.19 semantics restated)
is the dependency behind ``pp.py:469`` / ``:482,487,490`` / ``utils.py:299`` - PARITY
UNPINNED for those; the numpy and torch crop expressions (``utils.py:266-281``) live in the
reference itself and are pinned by ``tests/golden/crop_*.npz``.
"""
from __future__ import annotations

import numpy as np

CROP_NUMPY, CROP_TORCH, CROP_OPEN3D = 0, 1, 2


def crop_mode_from_backend(backend: str) -> int:
    """Back-end string dispatch of utils.py:254,272,298."""
    b = backend.lower()
    if b in ("np", "numpy"):
        return CROP_NUMPY
    if b in ("torch", "pytorch"):
        return CROP_TORCH
    return CROP_OPEN3D


def non_finite_mask(pos: np.ndarray, remove_nan=True, remove_infinite=True) -> np.ndarray:
    """Open3D ``remove_non_finite_points`` mask (pp.py:469-471).

    Only positions are inspected; True = keep.
    """
    mask = np.ones(pos.shape[0], dtype=bool)
    if remove_nan:
        mask &= ~np.isnan(pos).any(axis=1)
    if remove_infinite:
        mask &= ~np.isinf(pos).any(axis=1)
    return mask


def transform(pos: np.ndarray, T) -> np.ndarray:
    """Open3D ``t.PointCloud.transform`` (pp.py:482,487,490).

    float32, unfused, left to right: ``((t0*x + t1*y) + t2*z) + t3`` per row, then an IEEE
    divide by ``w`` (row 3).  Whether Open3D's build fuses the multiply-adds is
    compiler-dependent; the oracle fixes "unfused".
    """
    T = np.asarray(T, dtype=np.float32).reshape(4, 4)
    x, y, z = pos[:, 0].astype(np.float32), pos[:, 1].astype(np.float32), pos[:, 2].astype(np.float32)
    with np.errstate(all="ignore"):
        rows = [((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] for r in range(4)]
        out = np.stack([rows[0] / rows[3], rows[1] / rows[3], rows[2] / rows[3]], axis=1)
    return out.astype(np.float32)


def crop_mask(pos: np.ndarray, min_bound, max_bound, invert=False, mode=CROP_OPEN3D) -> np.ndarray:
    """ROI mask of ``crop_pointcloud`` (utils.py:240-301).

    * numpy back end (utils.py:266-269): bounds are Python lists -> float64 arrays, the
      float32 points are promoted, comparison in float64.
    * torch back end (utils.py:275-281): ``torch.as_tensor(list)`` -> float32 bounds.
    * open3d back end (utils.py:299, AABB built float32 at pp.py:311-313): inclusive
      float32 test, ``invert`` is the logical NOT of the mask.
    For numpy/torch ``invert`` is ``any((p <= min) | (p >= max))`` - not the complement:
    boundary points pass both, NaN rows pass neither.
    """
    if mode == CROP_NUMPY:
        p = pos.astype(np.float64)
        lo = np.asarray(min_bound, dtype=np.float64)
        hi = np.asarray(max_bound, dtype=np.float64)
    else:
        p = pos.astype(np.float32)
        lo = np.asarray(min_bound, dtype=np.float64).astype(np.float32)
        hi = np.asarray(max_bound, dtype=np.float64).astype(np.float32)
    with np.errstate(invalid="ignore"):
        if mode == CROP_OPEN3D:
            m = np.all((p >= lo) & (p <= hi), axis=1)
            return ~m if invert else m
        if invert:
            return np.any((p <= lo) | (p >= hi), axis=1)
        return np.all((p >= lo) & (p <= hi), axis=1)


def frontend(pos, *, nanskip_mask=None, dedup_mask_fn=None, remove_nan=True, remove_infinite=True,
             transforms=(), crop=None):
    """The fused front end in reference order (pp.py:450-506): [read_points NaN skip] ->
    [dedup] -> non-finite -> transform(s) -> crop.

    Returns ``(positions_out, src_idx, stage_mask)``.  ``stage_mask`` (uint8, length N) has
    bit0 = passed NaN skip, bit1 = passed dedup, bit2 = passed non-finite, bit3 = passed crop;
    a point that fails a stage has no later bits set.  ``crop`` is ``None`` or a dict
    ``{min, max, invert, mode}``.
    """
    n = pos.shape[0]
    stage = np.zeros(n, dtype=np.uint8)
    alive = np.ones(n, dtype=bool) if nanskip_mask is None else nanskip_mask.copy()
    stage[alive] |= 1
    if dedup_mask_fn is not None:
        idx = np.flatnonzero(alive)
        keep = dedup_mask_fn(pos[idx])
        alive = np.zeros(n, dtype=bool)
        alive[idx[keep]] = True
    stage[alive] |= 2
    if remove_nan or remove_infinite:
        alive &= non_finite_mask(pos, remove_nan, remove_infinite)
    stage[alive] |= 4
    p = pos.astype(np.float32).copy()
    for T in transforms:
        p = transform(p, T)
    if crop is not None:
        alive &= crop_mask(p, crop["min"], crop["max"], crop.get("invert", False),
                           crop.get("mode", CROP_OPEN3D))
    stage[alive] |= 8
    src = np.flatnonzero(alive).astype(np.uint32)
    return p[src], src, stage


def frontend_sorted(pos, index_fn, *, nanskip_mask=None, remove_nan=True, remove_infinite=True,
                    transforms=(), crop=None):
    """The front end with the numpy / torch duplicate-removal back ends (utils.py:520-542), which
    *reorder* the cloud: ``index_fn(points)`` is ``oracle.dedup.numpy_index`` (sorted first
    occurrences) or ``torch_compat_index`` (the inverse map, N rows); the rows it selects then go
    through non-finite -> transform(s) -> crop in that order (pp.py:466-506).

    Returns ``(positions_out, src_idx)``; ``src_idx`` indexes the input rows (with repeats in the
    torch mode).
    """
    n = pos.shape[0]
    alive = np.ones(n, dtype=bool) if nanskip_mask is None else nanskip_mask
    idx0 = np.flatnonzero(alive)
    sel = np.asarray(index_fn(pos[idx0])).reshape(-1).astype(np.int64)
    order = idx0[sel]
    p = pos[order].astype(np.float32).copy()
    keep = np.ones(order.shape[0], dtype=bool)
    if remove_nan or remove_infinite:
        keep &= non_finite_mask(p, remove_nan, remove_infinite)
    for T in transforms:
        p = transform(p, T)
    if crop is not None:
        keep &= crop_mask(p, crop["min"], crop["max"], crop.get("invert", False), crop.get("mode", CROP_OPEN3D))
    return p[keep], order[keep].astype(np.uint32)

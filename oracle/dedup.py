"""Duplicate removal: CPU oracle (test infrastructure).

Follows ``utils.py:509-546`` ``remove_duplicates``.  The numpy and torch branches live in
the reference and are pinned by ``tests/golden/dedup_*.npz``; the open3d branch
(``remove_duplicated_points``, utils.py:544) is PARITY UNPINNED (Open3D not installable).
"""
from __future__ import annotations

import numpy as np

DEDUP_OFF, DEDUP_OPEN3D, DEDUP_NUMPY, DEDUP_TORCH_COMPAT = 0, 1, 2, 3


def dedup_mode_from_backend(backend: str) -> int:
    b = backend.lower()
    if b in ("np", "numpy"):
        return DEDUP_NUMPY
    if b in ("torch", "pytorch"):
        return DEDUP_TORCH_COMPAT
    return DEDUP_OPEN3D


def open3d_mask(pos: np.ndarray) -> np.ndarray:
    """Open3D ``remove_duplicated_points`` (utils.py:544).

    Rows are hashed on the *bit pattern* of xyz (float32 reinterpret as int32, so -0.0 and
    +0.0 differ and identical NaN payloads coincide).  Which duplicate survives is a race in
    Open3D; the oracle fixes "lowest input index wins".  True = keep, order preserving.
    """
    bits = np.ascontiguousarray(pos, dtype=np.float32).view(np.int32).reshape(-1, 3)
    _, first = np.unique(bits, axis=0, return_index=True)
    mask = np.zeros(pos.shape[0], dtype=bool)
    mask[first] = True
    return mask


def numpy_index(pos: np.ndarray) -> np.ndarray:
    """utils.py:532-534: ``np.unique(points, axis=0, return_index=True, sorted=False)``.

    The survivors come out in lexicographic (x-major) order of the unique rows, each
    represented by its lowest input index; -0.0 == +0.0 merge, NaN rows never merge.
    """
    _, first = np.unique(pos, axis=0, return_index=True)
    return first


def torch_compat_index(pos: np.ndarray) -> np.ndarray:
    """utils.py:538-542 as written: ``select_by_index(inverse)`` - the reference passes the
    *inverse* map of ``torch.unique`` as if it were an index list, so the result has N rows
    ``points[inverse]``.  Reproduced only behind the explicit compat mode.
    """
    import torch  # the reference's own call; present on CPU in this image
    _, inverse = torch.unique(torch.from_numpy(np.ascontiguousarray(pos)), dim=0, return_inverse=True,
                              sorted=False)
    return inverse.numpy().reshape(-1)


def torch_compat_index_numpy(pos: np.ndarray) -> np.ndarray:
    """Same map through ``np.unique`` - identical to torch's on NaN-free input (torch orders
    NaN rows differently; the CUDA compat mode is defined for NaN-free input only, which is
    what reaches this stage when ``remove_nans`` is set: read_points drops NaN rows first)."""
    _, inverse = np.unique(pos, axis=0, return_inverse=True)
    return np.asarray(inverse).reshape(-1)


def sort_keys(pos: np.ndarray) -> np.ndarray:
    """Order-preserving float32 -> uint32 map the CUDA sort uses (csrc/sort.cu:sort_key_f32),
    restated: NaN -> 0xffffffff (after +inf, all NaNs tie), +-0 -> 0x80000000, negative values
    bit-inverted, positive values with the sign bit set."""
    b = np.ascontiguousarray(pos, dtype=np.float32).view(np.uint32).reshape(-1, 3).copy()
    mag = b & np.uint32(0x7fffffff)
    k = np.where((b & np.uint32(0x80000000)) != 0, ~b, b | np.uint32(0x80000000)).astype(np.uint32)
    k[mag == 0] = 0x80000000
    k[mag > 0x7f800000] = 0xffffffff
    return k


def unique_rows_model(pos: np.ndarray):
    """``(first_index, inverse)`` by the algorithm of the CUDA path: stable sort on the 96-bit
    key, a row starts a new group when its key differs from the previous row's or holds a NaN.
    ``tests/test_oracle_golden.py`` holds it against ``np.unique`` and the reference goldens, so
    the key rules are pinned independently of the GPU."""
    k = sort_keys(pos)
    order = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
    ks = k[order]
    head = np.ones(k.shape[0], dtype=bool)
    head[1:] = (ks[1:] != ks[:-1]).any(axis=1) | (ks[1:] == 0xffffffff).any(axis=1)
    inverse = np.empty(k.shape[0], dtype=np.int64)
    inverse[order] = np.cumsum(head) - 1
    return order[head], inverse

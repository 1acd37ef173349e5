"""Statistical and radius outlier removal: CPU oracle (test infrastructure).

Restates Open3D ``t.PointCloud.remove_statistical_outliers`` (called at ``pp.py:514-519``;
SURVEY appendix B8) and ``remove_radius_outliers`` (absent from the reference - only the
TODO at ``pp.py:37``; SURVEY appendix B9 adopts Open3D's) - PARITY UNPINNED.

Choices fixed here (the reference/Open3D leave them to the build or to compiler whim):
  * squared distance = ``(dx*dx + dy*dy) + dz*dz`` in float32, unfused;
  * KNN: the k nearest *including the query itself* (d = 0); ``k_eff = min(k, P)``;
    ``avg_i = (sum_j sqrtf(d2_ij) in ascending-distance order, sequential float32) / k_eff``;
  * global mean / std in float64 with the "adjacent pairwise tree" reduction
    (``tree_sum``: zero-pad to a power of two, repeatedly add neighbours) so that the CUDA
    reduction can mirror it bit for bit; ``sigma = sqrt(sum((avg-mu)^2) / (P-1))``;
    keep iff ``float64(avg_i) <= mu + std_ratio * sigma`` (tensor-API ``<=``); ``P < 2`` keeps all;
  * radius: ``r2 = float32(r) * float32(r)``; neighbour iff ``d2 <= r2`` (self included);
    keep iff ``count >= nb_points`` (tensor-API ``>=``).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def d2_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """(dx*dx + dy*dy) + dz*dz in float32 for broadcastable (...,3) arrays."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    dx = a[..., 0] - b[..., 0]
    dy = a[..., 1] - b[..., 1]
    dz = a[..., 2] - b[..., 2]
    return (dx * dx + dy * dy) + dz * dz


def tree_sum(v: np.ndarray) -> np.float64:
    """Adjacent pairwise binary-tree sum over a zero-padded power-of-two length (float64)."""
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    if v.size == 0:
        return np.float64(0.0)
    n = 1 << int(np.ceil(np.log2(max(v.size, 1))))
    s = np.zeros(n, dtype=np.float64)
    s[:v.size] = v
    while s.size > 1:
        s = s[0::2] + s[1::2]
    return s[0]


def knn_avg_distance(pos: np.ndarray, k: int, margin: int = 16, workers: int = -1) -> np.ndarray:
    """``avg_i`` (float32) as defined above.  cKDTree proposes candidates in float64; the
    float32 distances are recomputed with the defined formula and re-ranked."""
    P = pos.shape[0]
    k_eff = min(k, P)
    kq = min(P, k_eff + margin)
    tree = cKDTree(pos.astype(np.float64))
    _, nn = tree.query(pos.astype(np.float64), k=kq, workers=workers)
    nn = nn.reshape(P, kq)
    d2 = d2_f32(pos[:, None, :], pos[nn])
    d2.sort(axis=1)
    d = np.sqrt(d2[:, :k_eff]).astype(np.float32)
    s = d[:, 0].copy()
    for j in range(1, k_eff):
        s = s + d[:, j]                                  # sequential float32, ascending order
    return (s / np.float32(k_eff)).astype(np.float32)


def knn_avg_distance_brute(pos: np.ndarray, k: int) -> np.ndarray:
    """Exhaustive version for small P (no kd-tree; float32 throughout)."""
    P = pos.shape[0]
    k_eff = min(k, P)
    d2 = d2_f32(pos[:, None, :], pos[None, :, :])
    d2.sort(axis=1)
    d = np.sqrt(d2[:, :k_eff]).astype(np.float32)
    s = d[:, 0].copy()
    for j in range(1, k_eff):
        s = s + d[:, j]
    return (s / np.float32(k_eff)).astype(np.float32)


def statistical_threshold(avg: np.ndarray, std_ratio: float):
    """``(mu, sigma, thr)`` in float64 with the tree reduction."""
    P = avg.size
    a = avg.astype(np.float64)
    mu = tree_sum(a) / np.float64(P)
    dev = a - mu
    sigma = np.sqrt(tree_sum(dev * dev) / np.float64(P - 1))
    return mu, sigma, mu + np.float64(std_ratio) * sigma


def statistical_mask(pos: np.ndarray, nb_neighbors: int = 20, std_ratio: float = 2.0, brute=False):
    """Returns ``(mask, avg)``; True = keep (pp.py:516-518 defaults pp.py:174-175)."""
    if nb_neighbors < 1 or std_ratio <= 0:
        raise ValueError("nb_neighbors must be >= 1 and std_ratio > 0")
    P = pos.shape[0]
    if P == 0:
        return np.zeros(0, dtype=bool), np.zeros(0, dtype=np.float32)
    avg = knn_avg_distance_brute(pos, nb_neighbors) if brute else knn_avg_distance(pos, nb_neighbors)
    if P < 2:
        return np.ones(P, dtype=bool), avg
    _, _, thr = statistical_threshold(avg, std_ratio)
    return avg.astype(np.float64) <= thr, avg


def radius_counts_brute(pos: np.ndarray, radius: float) -> np.ndarray:
    r32 = np.float32(radius)
    r2 = r32 * r32
    d2 = d2_f32(pos[:, None, :], pos[None, :, :])
    return (d2 <= r2).sum(axis=1).astype(np.uint32)


def radius_mask(pos: np.ndarray, nb_points: int, radius: float, margin: int = 16, workers: int = -1,
                brute=False) -> np.ndarray:
    """Keep iff at least ``nb_points`` points (self included) lie within ``radius``."""
    P = pos.shape[0]
    if P == 0:
        return np.zeros(0, dtype=bool)
    if brute:
        return radius_counts_brute(pos, radius) >= nb_points
    r32 = np.float32(radius)
    r2 = r32 * r32
    kq = min(P, nb_points + margin)
    tree = cKDTree(pos.astype(np.float64))
    _, nn = tree.query(pos.astype(np.float64), k=kq, workers=workers)
    nn = nn.reshape(P, kq)
    d2 = d2_f32(pos[:, None, :], pos[nn])
    return (d2 <= r2).sum(axis=1) >= nb_points

"""Normal estimation: CPU oracle (test infrastructure; never imported by the product path).

Follows the call ``pp.py:521-530`` ``estimate_normals(radius=..., max_nn=...)``.  The arithmetic
lives in Open3D (``t::geometry::PointCloud::EstimateNormals`` -> hybrid-search covariances ->
``EstimatePointWiseNormalsWithFastEigen3x3``), which is not vendored and not installable here:
PARITY UNPINNED.  Restated from Open3D v0.18/0.19 as recalled in SURVEY.md appendix B11:

* neighbourhood = the ``max_nn`` nearest of the points within ``radius`` (query included);
  here: float32 ``d2 = (dx*dx + dy*dy) + dz*dz <= float32(radius)**2`` like the outlier stages,
  nearest by the total order ``(d2, index)``;
* fewer than 3 neighbours -> identity covariance -> normal (0, 0, 1);
* covariance = second moments / count - mean mean^T (float64 here, taken about the query point);
* normal = eigenvector of the smallest eigenvalue by the analytic symmetric 3x3 solver of
  D. Eberly, "A Robust Eigensolver for 3x3 Symmetric Matrices" (the one Open3D uses); no
  orientation step, so the sign is whatever the solver produces.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.spatial import cKDTree


def neighbourhoods(pos: np.ndarray, radius: float, max_nn: int):
    """List of index arrays: for every point its selected neighbours, ordered by (d2, index)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    r32 = np.float32(radius)
    r2 = np.float32(r32 * r32)
    tree = cKDTree(pos.astype(np.float64))
    cand = tree.query_ball_point(pos.astype(np.float64), float(r32) * (1.0 + 1e-5) + 1e-7, workers=-1)
    out = []
    for i, c in enumerate(cand):
        c = np.asarray(c, dtype=np.int64)
        d = pos[c] - pos[i]                                   # float32
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        keep = d2 <= r2
        c, d2 = c[keep], d2[keep]
        order = np.lexsort((c, d2))
        out.append(c[order][:max_nn])
    return out


def covariance(pos: np.ndarray, i: int, nb: np.ndarray) -> np.ndarray:
    if nb.size < 3:
        return np.eye(3)
    d = pos[nb].astype(np.float64) - pos[i].astype(np.float64)
    mean = d.sum(axis=0) / nb.size
    return (d.T @ d) / nb.size - np.outer(mean, mean)


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _eig_vector0(A, ev):
    r0 = [A[0] - ev, A[1], A[2]]
    r1 = [A[1], A[4] - ev, A[5]]
    r2 = [A[2], A[5], A[8] - ev]
    cs = [_cross(r0, r1), _cross(r0, r2), _cross(r1, r2)]
    ds = [_dot(c, c) for c in cs]
    best, dmax = cs[0], ds[0]
    if ds[1] > dmax:
        best, dmax = cs[1], ds[1]
    if ds[2] > dmax:
        best, dmax = cs[2], ds[2]
    inv = 1.0 / math.sqrt(dmax)
    return [best[0] * inv, best[1] * inv, best[2] * inv]


def _eig_vector1(A, w, ev1):
    if abs(w[0]) > abs(w[1]):
        inv = 1.0 / math.sqrt(w[0] * w[0] + w[2] * w[2])
        U = [-w[2] * inv, 0.0, w[0] * inv]
    else:
        inv = 1.0 / math.sqrt(w[1] * w[1] + w[2] * w[2])
        U = [0.0, w[2] * inv, -w[1] * inv]
    V = _cross(w, U)
    AU = [A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[4] * U[1] + A[5] * U[2],
          A[2] * U[0] + A[5] * U[1] + A[8] * U[2]]
    AV = [A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[4] * V[1] + A[5] * V[2],
          A[2] * V[0] + A[5] * V[1] + A[8] * V[2]]
    m00, m01, m11 = _dot(U, AU) - ev1, _dot(U, AV), _dot(V, AV) - ev1
    a00, a01, a11 = abs(m00), abs(m01), abs(m11)
    if a00 >= a11:
        if max(a00, a01) > 0.0:
            if a00 >= a01:
                m01 /= m00
                m00 = 1.0 / math.sqrt(1.0 + m01 * m01)
                m01 *= m00
            else:
                m00 /= m01
                m01 = 1.0 / math.sqrt(1.0 + m00 * m00)
                m00 *= m01
            return [m01 * U[k] - m00 * V[k] for k in range(3)]
        return U
    if max(a11, a01) > 0.0:
        if a11 >= a01:
            m01 /= m11
            m11 = 1.0 / math.sqrt(1.0 + m01 * m01)
            m01 *= m11
        else:
            m11 /= m01
            m01 = 1.0 / math.sqrt(1.0 + m11 * m11)
            m11 *= m01
        return [m11 * U[k] - m01 * V[k] for k in range(3)]
    return U


def normal_from_covariance(C: np.ndarray):
    """Eigenvector of the smallest eigenvalue of the symmetric 3x3 ``C`` (analytic solver)."""
    Cf = [float(v) for v in np.asarray(C, dtype=np.float64).reshape(9)]
    mx = max(Cf)
    if mx == 0.0:
        return [0.0, 0.0, 0.0]
    A = [v / mx for v in Cf]
    norm = A[1] * A[1] + A[2] * A[2] + A[5] * A[5]
    if not norm > 0.0:
        if Cf[0] < Cf[4] and Cf[0] < Cf[8]:
            return [1.0, 0.0, 0.0]
        if Cf[4] < Cf[0] and Cf[4] < Cf[8]:
            return [0.0, 1.0, 0.0]
        return [0.0, 0.0, 1.0]
    q = (A[0] + A[4] + A[8]) / 3.0
    b00, b11, b22 = A[0] - q, A[4] - q, A[8] - q
    p = math.sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0)
    c00 = b11 * b22 - A[5] * A[5]
    c01 = A[1] * b22 - A[5] * A[2]
    c02 = A[1] * A[5] - b11 * A[2]
    det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p)
    half_det = min(max(det * 0.5, -1.0), 1.0)
    angle = math.acos(half_det) / 3.0
    beta2 = math.cos(angle) * 2.0
    beta0 = math.cos(angle + 2.09439510239319549) * 2.0
    beta1 = -(beta0 + beta2)
    e0, e1, e2 = q + p * beta0, q + p * beta1, q + p * beta2
    if half_det >= 0.0:
        v2 = _eig_vector0(A, e2)
        if e2 < e0 and e2 < e1:
            return v2
        v1 = _eig_vector1(A, v2, e1)
        if e1 < e0 and e1 < e2:
            return v1
        return _cross(v1, v2)
    v0 = _eig_vector0(A, e0)
    if e0 < e1 and e0 < e2:
        return v0
    v1 = _eig_vector1(A, v0, e1)
    if e1 < e0 and e1 < e2:
        return v1
    return _cross(v0, v1)


def estimate_normals(pos: np.ndarray, radius: float, max_nn: int):
    """Returns ``(normals float32[N,3], counts int32[N], covariances float64[N,3,3])``."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    nbs = neighbourhoods(pos, radius, max_nn)
    n = pos.shape[0]
    normals = np.zeros((n, 3), dtype=np.float32)
    counts = np.zeros(n, dtype=np.int32)
    covs = np.zeros((n, 3, 3), dtype=np.float64)
    for i, nb in enumerate(nbs):
        C = covariance(pos, i, nb)
        covs[i] = C
        counts[i] = nb.size
        normals[i] = np.asarray(normal_from_covariance(C), dtype=np.float64).astype(np.float32)
    return normals, counts, covs

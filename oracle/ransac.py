"""RANSAC ground-plane segmentation: CPU oracle (test infrastructure).

Restates Open3D's legacy ``PointCloud::SegmentPlane`` which the tensor API
``segment_plane`` (called at ``pp.py:533-543``) converts to; SURVEY appendix B10 -
PARITY UNPINNED (Open3D not installable; its loop is OpenMP-parallel and unseeded in the
reference, so the reference itself is not reproducible run to run).

Definitions fixed here so the CUDA path can match bit for bit:
  * all arithmetic in float64, unfused, in exactly the operation order written below;
  * hypothesis generator: a counter-based stream.  Draw c of iteration it is
    ``z = splitmix64(seed + (it << 32) + c)``; candidate index ``((z >> 32) * P) >> 32``;
    candidates equal to an already accepted index are rejected; the first ``ransac_n``
    accepted indices form the sample.  An explicit ``int32[iters, ransac_n]`` table may be
    supplied instead;
  * score: ``dist = |((a*x + b*y) + c*z) + d|``; inlier iff ``dist < thr`` (strict);
    the error term (a tie-breaker only) is accumulated as an integer so it is independent of
    summation order: ``err += rint(d32*d32 * float32(2^16 / thr^2))`` over inliers, with
    ``d32`` the distance evaluated with float32 fused multiply-adds (see ``score``);
  * selection has *sequential* semantics: iterations are visited in order; a hypothesis
    replaces the best when it has more inliers, or the same number and a smaller ``err``
    (equal inlier count => rmse order == err order); after each improvement
    ``break_it = 0 if inl == P else min(log(1-p) / log(1 - fitness^n), num_iterations)``
    with ``fitness^n`` by repeated multiplication, and iterations with ``it > break_it`` are
    skipped;
  * final inliers: ``dist < thr`` against the best hypothesis (none when no valid
    hypothesis was found), ascending index; the returned plane is the least-squares refit
    on the final inliers (compared within 1e-5, so its summation order is free).
"""
from __future__ import annotations

import math

import numpy as np

M64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def sample_table(seed: int, num_iterations: int, ransac_n: int, P: int) -> np.ndarray:
    """``int32[num_iterations, ransac_n]`` of distinct indices per row."""
    out = np.zeros((num_iterations, ransac_n), dtype=np.int32)
    for it in range(num_iterations):
        got = []
        c = 0
        while len(got) < ransac_n:
            z = splitmix64((seed + (it << 32) + c) & M64)
            idx = ((z >> 32) * P) >> 32
            c += 1
            if idx not in got:
                got.append(idx)
        out[it] = got
    return out


def _f(x):
    return np.float64(x)


def fit_plane(Q: np.ndarray) -> np.ndarray:
    """Open3D ``GetPlaneFromPoints`` ("fast plane fit"); Q is (n,3) float64.  Sums are
    sequential in row order."""
    n = Q.shape[0]
    cx = cy = cz = _f(0.0)
    for p in Q:
        cx = cx + p[0]
        cy = cy + p[1]
        cz = cz + p[2]
    cx, cy, cz = cx / _f(n), cy / _f(n), cz / _f(n)
    xx = xy = xz = yy = yz = zz = _f(0.0)
    for p in Q:
        rx, ry, rz = p[0] - cx, p[1] - cy, p[2] - cz
        xx = xx + rx * rx
        xy = xy + rx * ry
        xz = xz + rx * rz
        yy = yy + ry * ry
        yz = yz + ry * rz
        zz = zz + rz * rz
    return _plane_from_moments(cx, cy, cz, xx, xy, xz, yy, yz, zz)


def _plane_from_moments(cx, cy, cz, xx, xy, xz, yy, yz, zz) -> np.ndarray:
    det_x = yy * zz - yz * yz
    det_y = xx * zz - xz * xz
    det_z = xx * yy - xy * xy
    if det_x >= det_y and det_x >= det_z:
        nx, ny, nz = det_x, xz * yz - xy * zz, xy * yz - xz * yy
    elif det_y >= det_z:                 # det_y is the largest (det_x was not)
        nx, ny, nz = xz * yz - xy * zz, det_y, xy * xz - yz * xx
    else:
        nx, ny, nz = xy * yz - xz * yy, xy * xz - yz * xx, det_z
    norm = np.sqrt((nx * nx + ny * ny) + nz * nz)
    if not (norm > 0.0):
        return np.zeros(4, dtype=np.float64)
    nx, ny, nz = nx / norm, ny / norm, nz / norm
    d = -((nx * cx + ny * cy) + nz * cz)
    return np.array([nx, ny, nz, d], dtype=np.float64)


def triangle_plane(p0, p1, p2) -> np.ndarray:
    e1 = p1 - p0
    e2 = p2 - p0
    nx = e1[1] * e2[2] - e1[2] * e2[1]
    ny = e1[2] * e2[0] - e1[0] * e2[2]
    nz = e1[0] * e2[1] - e1[1] * e2[0]
    norm = np.sqrt((nx * nx + ny * ny) + nz * nz)
    if not (norm > 0.0):
        return np.zeros(4, dtype=np.float64)
    nx, ny, nz = nx / norm, ny / norm, nz / norm
    d = -((nx * p0[0] + ny * p0[1]) + nz * p0[2])
    return np.array([nx, ny, nz, d], dtype=np.float64)


def hypothesis(P64: np.ndarray, sample: np.ndarray) -> np.ndarray:
    Q = P64[np.asarray(sample, dtype=np.int64)]
    if len(sample) == 3:
        return triangle_plane(Q[0], Q[1], Q[2])
    return fit_plane(Q)


def plane_distance(P64: np.ndarray, plane: np.ndarray) -> np.ndarray:
    a, b, c, d = plane
    return np.abs(((a * P64[:, 0] + b * P64[:, 1]) + c * P64[:, 2]) + d)


ERR_BITS = 16   # error quantum = thr^2 / 2^16


def err_scale(thr: float) -> np.float32:
    """float32(2^16 / thr^2), the divide and the square in float64."""
    return np.float32(np.float64(1 << ERR_BITS) / (np.float64(thr) * np.float64(thr)))


def fma32(a, b, c) -> np.ndarray:
    """Exact emulation of the float32 fused multiply-add ``round32(a*b + c)`` (one rounding).

    ``a*b`` is exact in float64 (24+24 significand bits); the float64 sum ``s`` is rounded once,
    and rounding ``s`` again to float32 can only go wrong when ``s`` lands exactly on a float32
    midpoint while the true sum does not - detected with the TwoSum residual and nudged one
    float64 ulp towards the true value first."""
    p = np.asarray(a, dtype=np.float32).astype(np.float64) * np.asarray(b, dtype=np.float32).astype(np.float64)
    c = np.broadcast_to(np.asarray(c, dtype=np.float32).astype(np.float64), p.shape)
    s = p + c
    bb = s - p
    e = (p - (s - bb)) + (c - bb)                     # s + e == p + c exactly
    mid = (s.view(np.uint64) & np.uint64((1 << 29) - 1)) == np.uint64(1 << 28)
    fix = mid & (e != 0.0) & np.isfinite(s)
    s = np.where(fix, np.nextafter(s, np.where(e > 0.0, np.inf, -np.inf)), s)
    return s.astype(np.float32)


def plane_distance_f32(P32: np.ndarray, plane: np.ndarray) -> np.ndarray:
    """SIGNED float32 distance ``fma(c, z, fma(b, y, fma(a, x, d)))`` with the coefficients
    rounded to float32 (the scoring kernel's FFMA2 chain)."""
    a, b, c, d = (np.float32(v) for v in plane)
    P32 = np.ascontiguousarray(P32, dtype=np.float32)
    return fma32(c, P32[:, 2], fma32(b, P32[:, 1], fma32(a, P32[:, 0], d)))


def score(P64: np.ndarray, plane: np.ndarray, thr: float):
    """``(inlier_count, err_q)``.  Inlier-ness is decided in float64 (``dist < thr``); the error
    term - only a tie-breaker between hypotheses with equal inlier counts - is an integer sum
    (independent of summation order) built from the float32 distance ``d32``:
    ``q = rint(float32(d32*d32) * float32(2^16/thr^2))``, the product exact and the rounding
    half-to-even (what ``fma(t, scale, 2^23)`` leaves in the float32 mantissa)."""
    dist = plane_distance(P64, plane)
    inl = dist < np.float64(thr)
    d32 = plane_distance_f32(P64[inl].astype(np.float32), plane)
    t = (d32 * d32).astype(np.float32)
    q = np.rint(t.astype(np.float64) * np.float64(err_scale(thr))).astype(np.uint64)
    return int(inl.sum()), int(q.sum(dtype=np.uint64))


def select(scores, valid, P: int, ransac_n: int, num_iterations: int, probability: float) -> int:
    """Sequential-semantics selection; returns the best iteration or -1."""
    best_inl, best_err, best_it = 0, 0, -1
    break_it = float(num_iterations)
    log1mp = math.log(1.0 - probability) if probability < 1.0 else -math.inf
    for it in range(num_iterations):
        if float(it) > break_it:
            continue
        if not valid[it]:
            continue
        inl, err = scores[it]
        if inl > best_inl or (inl == best_inl and inl > 0 and err < best_err):
            best_inl, best_err, best_it = inl, err, it
            if inl >= P:
                break_it = 0.0
            else:
                fitness = np.float64(inl) / np.float64(P)
                fn = fitness
                for _ in range(ransac_n - 1):
                    fn = fn * fitness
                denom = math.log(1.0 - float(fn))
                # 1 - fitness^n rounds to 1 for tiny fitness: treat as "no bound" instead of
                # Open3D's IEEE accident (negative / +0 = -inf, which would stop the search)
                cand = math.inf if denom == 0.0 else log1mp / denom
                break_it = cand if cand < float(num_iterations) else float(num_iterations)
    return best_it


def segment_plane(pos: np.ndarray, distance_threshold=0.2, ransac_n=5, num_iterations=100,
                  probability=0.99, seed: int = 0, samples: np.ndarray | None = None):
    """Returns ``(plane[4] float64, inlier_idx int64 ascending, info dict)``."""
    P = pos.shape[0]
    if not (0.0 < probability <= 1.0):
        raise ValueError("probability must be in (0, 1]")
    if ransac_n < 3:
        raise ValueError("ransac_n must be >= 3")
    if P < ransac_n:
        raise ValueError("not enough points")
    P64 = pos.astype(np.float64)
    if samples is None:
        samples = sample_table(seed, num_iterations, ransac_n, P)
    planes = np.stack([hypothesis(P64, samples[it]) for it in range(num_iterations)])
    valid = np.any(planes != 0.0, axis=1)
    scores = [score(P64, planes[it], distance_threshold) if valid[it] else (0, 0)
              for it in range(num_iterations)]
    best_it = select(scores, valid, P, ransac_n, num_iterations, probability)
    if best_it < 0:
        return np.zeros(4), np.zeros(0, dtype=np.int64), {"best_it": -1, "planes": planes, "scores": scores}
    best = planes[best_it]
    inl = np.flatnonzero(plane_distance(P64, best) < np.float64(distance_threshold)).astype(np.int64)
    refit = fit_plane_fast(P64[inl]) if inl.size else np.zeros(4)
    return refit, inl, {"best_it": best_it, "best_plane": best, "planes": planes, "scores": scores,
                        "samples": samples}


def fit_plane_fast(Q: np.ndarray) -> np.ndarray:
    """Vectorised refit (same formula as ``fit_plane``; summation order free, 1e-5 tolerance)."""
    c = Q.mean(axis=0)
    r = Q - c
    xx, xy, xz = (r[:, 0] * r[:, 0]).sum(), (r[:, 0] * r[:, 1]).sum(), (r[:, 0] * r[:, 2]).sum()
    yy, yz, zz = (r[:, 1] * r[:, 1]).sum(), (r[:, 1] * r[:, 2]).sum(), (r[:, 2] * r[:, 2]).sum()
    return _plane_from_moments(c[0], c[1], c[2], xx, xy, xz, yy, yz, zz)

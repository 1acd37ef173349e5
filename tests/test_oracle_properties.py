"""Size-independent properties of the CPU oracle (test infrastructure): the checker must be consistent with
itself before the CUDA path is held against it.  Seeded numpy inputs, a few thousand points each."""
import numpy as np
import pytest

from oracle import dedup, filters, outliers, ransac, voxel


def cloud(seed, n=4000, spread=30.0):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(n, 3)) * np.array([spread, spread, 2.0])).astype(np.float32)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_voxel_membership_and_means(seed):
    pos = cloud(seed)
    pos = np.concatenate([pos, pos[:500]])                       # repeated points join their voxel
    vs = 0.5
    ref = voxel.voxel_down_sample(pos, vs, np.ones(len(pos), np.float32), fixed=True)
    p2v, counts = ref["p2v"], ref["counts"]
    assert counts.sum() == len(pos) and p2v.min() == 0 and p2v.max() == len(counts) - 1
    # first-occurrence order: the first point of voxel v comes before the first point of voxel v + 1
    first = np.full(len(counts), len(pos), dtype=np.int64)
    np.minimum.at(first, p2v, np.arange(len(pos)))
    assert np.all(np.diff(first) > 0)
    # every point lies in the cell of its voxel, every centroid inside the cell's closed box
    idx = voxel.voxel_index(pos, vs)
    assert np.array_equal(idx, idx[first][p2v])
    cen = ref["positions"]
    lo = idx[first].astype(np.float64) * vs
    assert np.all(cen >= lo - 1e-4) and np.all(cen <= lo + vs + 1e-4)
    # order-independent fixed-point mean == float64 mean within the contract's 1e-5, and == Open3D-style float32 sums
    exact = np.zeros((len(counts), 3))
    np.add.at(exact, p2v, pos.astype(np.float64))
    exact /= counts[:, None]
    assert np.abs(cen - exact).max() < 1e-5
    assert np.abs(voxel.centroids_o3d(pos, p2v, len(counts)) - cen).max() < 1e-4
    # permuting the points inside the sums does not change a bit (that is what "fixed" buys)
    perm = np.random.default_rng(seed).permutation(len(pos))
    again = voxel.centroids_fixed(pos[perm], p2v[perm], len(counts))
    assert np.array_equal(again.view(np.uint32), cen.view(np.uint32))
    # idempotence at the centroid level: down-sampling the centroids with the same grid keeps every one of them
    twice = voxel.voxel_down_sample(cen, vs, None, fixed=True)
    assert twice["positions"].shape[0] <= cen.shape[0]


@pytest.mark.parametrize("seed", [4, 5])
def test_dedup_back_ends_agree_on_the_kept_set(seed):
    pos = cloud(seed, 3000)
    pos = np.concatenate([pos, pos[::7], pos[::11]])
    pos[5] = -0.0 * pos[5]
    keep = dedup.open3d_mask(pos)                                 # lowest index of every distinct row, order kept
    assert keep.sum() == len(np.unique(pos.view(np.uint32).reshape(len(pos), 3), axis=0))
    assert dedup.open3d_mask(pos[keep]).all()                     # idempotent
    first, inv = dedup.unique_rows_model(pos)                     # numpy semantics: sorted unique rows, first indices
    ref_u, ref_first, ref_inv = np.unique(pos, axis=0, return_index=True, return_inverse=True)
    assert np.array_equal(first, ref_first) and np.array_equal(inv, ref_inv.reshape(-1))
    assert np.array_equal(pos[first][inv].view(np.uint32), pos.view(np.uint32)) or np.array_equal(pos[first][inv], pos)
    assert np.array_equal(dedup.numpy_index(pos), ref_first)


def test_crop_semantics_differ_only_where_they_must():
    pos = cloud(6, 5000, 50.0)
    pos[::97] = np.nan
    lo, hi = [-20.0, -20.0, -1.0], [20.0, 20.0, 1.0]
    m = {mode: filters.crop_mask(pos, lo, hi, False, mode) for mode in (filters.CROP_NUMPY, filters.CROP_TORCH, filters.CROP_OPEN3D)}
    assert np.array_equal(m[filters.CROP_TORCH], m[filters.CROP_OPEN3D])        # both compare in float32
    nan_rows = np.isnan(pos).any(1)
    assert not m[filters.CROP_NUMPY][nan_rows].any()
    inv_o3d = filters.crop_mask(pos, lo, hi, True, filters.CROP_OPEN3D)
    inv_np = filters.crop_mask(pos, lo, hi, True, filters.CROP_NUMPY)
    assert np.array_equal(inv_o3d, ~m[filters.CROP_OPEN3D])                      # Open3D: logical NOT (NaN rows kept)
    assert not inv_np[nan_rows].any() and inv_o3d[nan_rows].all()                # numpy: "outside on some axis" (NaN rows dropped)


def test_transform_is_a_float32_left_to_right_product():
    pos = cloud(7, 2000)
    T = np.eye(4)
    T[:3, :3] = [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]
    T[:3, 3] = [1.5, -2.0, 0.25]
    out = filters.transform(pos, T)
    assert out.dtype == np.float32
    assert np.array_equal(out[:, 0], (-pos[:, 1] + np.float32(1.5)).astype(np.float32))
    back = filters.transform(out, np.linalg.inv(T))
    assert np.abs(back - pos).max() < 1e-4


def test_outlier_oracles_brute_force_cross_check():
    pos = cloud(8, 1500, 3.0)
    for nb, r in ((5, 0.5), (12, 1.0)):
        assert np.array_equal(outliers.radius_mask(pos, nb, r), outliers.radius_counts_brute(pos, r) >= nb)
    a = outliers.knn_avg_distance(pos, 10)
    b = outliers.knn_avg_distance_brute(pos, 10)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # adjacent-pairwise tree == exact sum for values that add exactly, and within rounding otherwise
    v = np.arange(1, 1001, dtype=np.float64)
    assert outliers.tree_sum(v) == 500500.0
    w = np.random.default_rng(9).random(1000)
    assert abs(outliers.tree_sum(w) - np.sum(w)) < 1e-9


def test_ransac_oracle_selects_the_planted_plane():
    rng = np.random.default_rng(10)
    ground = np.c_[rng.uniform(-30, 30, (3000, 2)), rng.normal(0.0, 0.02, 3000) - 1.8]
    clutter = rng.uniform([-30, -30, -1.5], [30, 30, 4.0], (1500, 3))
    pos = np.concatenate([ground, clutter]).astype(np.float32)
    plane, inl, info = ransac.segment_plane(pos, 0.2, 5, 100, 0.99, seed=3)
    assert info["best_it"] >= 0
    assert abs(abs(plane[2]) - 1.0) < 1e-2 and abs(abs(plane[3]) - 1.8) < 5e-2
    assert (inl < 3000).sum() > 2900                              # nearly all ground points are inliers
    plane2, inl2, _ = ransac.segment_plane(pos, 0.2, 5, 100, 0.99, seed=3)
    assert np.array_equal(inl, inl2) and np.array_equal(plane, plane2)   # seeded: reproducible

"""The oracle against the golden vectors minted from the reference's own utils.py
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np

from oracle import dedup as odedup
from oracle import filters, pc2


def test_crop_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "crop.npz"))
    p = g["points"]
    for case in ("roi", "frac"):
        lo, hi = g[f"{case}_min"].tolist(), g[f"{case}_max"].tolist()
        for backend, mode in (("numpy", filters.CROP_NUMPY), ("torch", filters.CROP_TORCH)):
            for invert in (False, True):
                want = g[f"{case}_{backend}_{'inv' if invert else 'fwd'}"]
                got = filters.crop_mask(p, lo, hi, invert=invert, mode=mode)
                assert np.array_equal(got, want), (case, backend, invert)
    # the fractional bounds must actually separate the f64 and f32 comparisons
    lo, hi = g["frac_min"].tolist(), g["frac_max"].tolist()
    assert not np.array_equal(filters.crop_mask(p, lo, hi, mode=filters.CROP_NUMPY),
                              filters.crop_mask(p, lo, hi, mode=filters.CROP_TORCH))


def test_crop_invert_is_not_complement(golden_dir):
    g = np.load(os.path.join(golden_dir, "crop.npz"))
    fwd, inv = g["roi_numpy_fwd"], g["roi_numpy_inv"]
    assert (fwd & inv).any()            # boundary points pass both
    assert (~fwd & ~inv).any()          # NaN rows pass neither
    o3d_inv = filters.crop_mask(g["points"], g["roi_min"].tolist(), g["roi_max"].tolist(), invert=True,
                                mode=filters.CROP_OPEN3D)
    assert np.array_equal(o3d_inv, ~filters.crop_mask(g["points"], g["roi_min"].tolist(),
                                                      g["roi_max"].tolist(), mode=filters.CROP_OPEN3D))


def test_dedup_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "dedup.npz"))
    p = g["points"]
    assert np.array_equal(odedup.numpy_index(p), g["numpy_index"])
    # torch compat mode: N rows, points[inverse]
    want = g["torch_index"]
    got = odedup.torch_compat_index(p)
    assert got.shape == want.shape == (p.shape[0],)
    assert np.array_equal(got, want)
    finite = np.isfinite(p).all(axis=1)
    assert np.array_equal(odedup.torch_compat_index_numpy(p[finite]), odedup.torch_compat_index(p[finite]))
    # open3d mode keeps bitwise-distinct -0.0 / +0.0 rows and merges identical NaN rows
    m = odedup.open3d_mask(p)
    assert m[10] and m[20] and m[30] and not m[40]


def test_sort_key_model(golden_dir):
    """The key rules of the CUDA sorted-unique path (csrc/sort.cu), restated in
    oracle.dedup.unique_rows_model, against the reference golden and np.unique on adversarial rows."""
    g = np.load(os.path.join(golden_dir, "dedup.npz"))
    first, inverse = odedup.unique_rows_model(g["points"])
    assert np.array_equal(first, g["numpy_index"])                       # reference numpy back end
    finite = np.isfinite(g["points"]).all(axis=1)
    _, inv_f = odedup.unique_rows_model(g["points"][finite])
    assert np.array_equal(inv_f, odedup.torch_compat_index(g["points"][finite]))   # reference torch back end
    rng = np.random.default_rng(5)
    vals = np.array([0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf, 1.0, -1.0, 1e-45, -1e-45, 3.5, 2.0,
                     np.float32(3.4e38), np.float32(-3.4e38)], np.float32)
    for _ in range(100):
        n = int(rng.integers(1, 300))
        q = vals[rng.integers(0, len(vals), size=(n, 3))]
        q.view(np.uint32)[np.isnan(q) & (rng.random(q.shape) < 0.5)] |= np.uint32(0x1234)   # NaN payloads
        first, inverse = odedup.unique_rows_model(q)
        _, fi, ii = np.unique(q, axis=0, return_index=True, return_inverse=True)
        assert np.array_equal(first, fi) and np.array_equal(inverse, np.asarray(ii).reshape(-1))


def test_convert_and_metadata_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "convert.npz"))
    meta = json.load(open(os.path.join(golden_dir, "metadata.json")))
    for name in ("velodyne", "autoware", "livox", "f64xyz", "rgb"):
        spec = meta[name]
        dt = np.dtype({"names": [f[0] for f in spec["fields"]], "formats": [f[1] for f in spec["fields"]],
                       "offsets": [f[2] for f in spec["fields"]], "itemsize": spec["itemsize"]})
        arr = np.frombuffer(g[f"{name}__bytes"].tobytes(), dtype=dt)
        m = pc2.get_pointcloud_metadata(dt.names)
        assert m == spec["metadata"], name
        d = pc2.convert_pointcloud_to_numpy(arr, dict(m, field_names=dt.names))
        keys = {k.split("__")[1] for k in g.files if k.startswith(name + "__")} - {"bytes"}
        assert set(d) == keys
        for k in keys:
            want = g[f"{name}__{k}"]
            assert d[k].dtype == want.dtype and np.array_equal(d[k], want, equal_nan=True), (name, k)
    for entry in meta["mappings"]:
        assert pc2.get_pointcloud_metadata(entry["names"]) == entry["metadata"]
    for entry in meta["packed"]:
        fields, step = pc2.packed_fields(entry["names"], entry["datatypes"])
        assert step == entry["point_step"]
        assert [[n, o, d, 1] for (n, o, d) in fields] == entry["fields"]


def test_rgb_helpers_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "rgb.npz"))
    assert np.array_equal(pc2.extract_rgb_from_pointcloud(g["merged_float"]), g["extracted"])
    assert np.array_equal(g["extracted"], np.stack([g["r"], g["g"], g["b"]], 1))


def test_normals_oracle_solver_against_eigh():
    """oracle.normals.normal_from_covariance (the analytic symmetric 3x3 solver the CUDA path mirrors)
    against numpy's eigh on generic, near-planar, exactly planar and needle-shaped neighbourhoods, and
    the fixed fall-backs (identity covariance -> +z, diagonal -> axis of the smallest entry)."""
    from oracle import normals as onrm
    rng = np.random.default_rng(0)
    for t in range(800):
        kind = t % 4
        P = rng.normal(size=(20, 3))
        if kind == 1:
            P *= [1.0, 1.0, 1e-3]
        elif kind == 2:
            P[:, 2] = 0.3 * P[:, 0] - 0.2 * P[:, 1]
        elif kind == 3:
            P = rng.normal(size=(5, 3)) * [1.0, 1e-2, 1e-4]
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        P = P @ R.T + rng.uniform(-50, 50, 3)
        C = np.cov(P.T, bias=True)
        n = np.array(onrm.normal_from_covariance(C))
        w, v = np.linalg.eigh(C)
        if (w[1] - w[0]) / w[2] > 1e-6:
            assert abs(abs(n @ v[:, 0]) - 1.0) < 1e-6 and abs(np.linalg.norm(n) - 1.0) < 1e-9
    assert onrm.normal_from_covariance(np.eye(3)) == [0.0, 0.0, 1.0]
    assert onrm.normal_from_covariance(np.diag([3.0, 1.0, 2.0])) == [0.0, 1.0, 0.0]
    pos = rng.uniform(-1, 1, size=(400, 3)).astype(np.float32)
    pos[:, 2] = 0.25
    normals, counts, cov = onrm.estimate_normals(pos, 0.3, 30)
    assert counts.max() <= 30 and np.allclose(np.abs(normals[counts >= 3][:, 2]), 1.0, atol=1e-6)

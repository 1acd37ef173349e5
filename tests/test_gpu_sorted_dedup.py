"""GPU parity tests of the numpy / torch back ends of remove_duplicates (utils.py:520-542): the
device radix sort + head flags (csrc/sort.cu) against the reference-minted golden
(tests/golden/dedup.npz), np.unique / torch.unique on seeded inputs, and the oracle pipeline.
Everything here is index work: compared bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

T_A = np.array([[0.9986295, -0.0523360, 0.0, 1.5], [0.0523360, 0.9986295, 0.0, -0.25], [0.0, 0.0, 1.0, 1.8],
                [0.0, 0.0, 0.0, 1.0]])


@pytest.fixture(scope="module")
def env():
    from autodriver_pointcloud_preprocessor_b200 import _capi, engine, synth
    ctx = engine.Context(max_points=1_600_000)
    yield dict(ctx=ctx, engine=engine, capi=_capi, synth=synth)
    ctx.close()


def to_xyzi(p):
    return torch.from_numpy(np.concatenate([p, np.zeros((p.shape[0], 1), np.float32)], 1)).cuda()


def gpu_unique(ctx, p):
    first, inverse, cnt = ctx.unique_rows(to_xyzi(p), want_first=True, want_inverse=True)
    ctx.check()
    k = int(cnt.item())
    return first[:k].cpu().numpy().astype(np.int64), inverse[:p.shape[0]].cpu().numpy().astype(np.int64)


def test_unique_rows_reference_golden(env, golden_dir):
    """first index == what the reference's numpy back end selected; inverse == the reference's torch
    back end on the NaN-free rows (torch orders NaN rows by an inconsistent comparator)."""
    from oracle import dedup as odedup
    g = np.load(os.path.join(golden_dir, "dedup.npz"))
    p = g["points"]
    for _ in range(2):                                   # scratch and look-back words are reused
        first, inverse = gpu_unique(env["ctx"], p)
        assert np.array_equal(first, g["numpy_index"])
        assert np.array_equal(inverse, odedup.torch_compat_index_numpy(p))
    finite = np.isfinite(p).all(axis=1)
    first_f, inverse_f = gpu_unique(env["ctx"], np.ascontiguousarray(p[finite]))
    assert np.array_equal(inverse_f, odedup.torch_compat_index(p[finite]))       # torch.unique itself


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 2047, 2048, 2049, 5000, 70001])
def test_unique_rows_adversarial(env, n):
    """+-0, NaN payloads, infinities, denormals, extreme magnitudes and heavy duplication."""
    rng = np.random.default_rng(n)
    vals = np.array([0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf, 1.0, -1.0, 1e-45, -1e-45, 3.5, 2.0,
                     np.float32(3.4e38), np.float32(-3.4e38), 1e-38, 255.0, 256.0, -65536.0], np.float32)
    q = vals[rng.integers(0, len(vals), size=(n, 3))]
    q.view(np.uint32)[np.isnan(q) & (rng.random(q.shape) < 0.5)] |= np.uint32(0x1234)
    mix = rng.random(n) < 0.5
    q[mix] = rng.uniform(-100, 100, size=(int(mix.sum()), 3)).astype(np.float32)
    first, inverse = gpu_unique(env["ctx"], q)
    _, fi, ii = np.unique(q, axis=0, return_index=True, return_inverse=True)
    assert np.array_equal(first, fi)
    assert np.array_equal(inverse, np.asarray(ii).reshape(-1))


def test_unique_rows_empty_and_constant(env):
    ctx = env["ctx"]
    first, inverse, cnt = ctx.unique_rows(torch.zeros((0, 4), device="cuda"), want_inverse=True)
    ctx.check()
    assert int(cnt.item()) == 0
    p = np.tile(np.array([[1.5, -2.25, 3.0]], np.float32), (5000, 1))          # every pass is the identity
    first, inverse = gpu_unique(ctx, p)
    assert np.array_equal(first, [0]) and not inverse.any()


@pytest.mark.parametrize("shape", [(128, 2048), (128, 11719)])
def test_unique_rows_full_size(env, shape):
    """BASELINE sizes (C2 262 144 points, C4 1.5 M): against np.unique, plus the size-independent
    properties: sortedness of the selected rows, first-occurrence, inverse consistency."""
    synth = env["synth"]
    scan = synth.lidar_scan(seed=17, n_beams=shape[0], n_az=shape[1], nan_frac=0.0)
    p = np.ascontiguousarray(scan["positions"][:1_500_000])
    first, inverse = gpu_unique(env["ctx"], p)
    rows = p[first]
    assert np.array_equal(rows[inverse], p)                                         # inverse rebuilds the cloud
    k = np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))
    assert np.array_equal(k, np.arange(len(rows)))                                  # strictly sorted rows
    assert len(np.unique(inverse)) == len(first)
    seen = np.full(len(first), p.shape[0], dtype=np.int64)
    np.minimum.at(seen, inverse, np.arange(p.shape[0]))
    assert np.array_equal(seen, first)                                              # lowest index represents
    _, fi = np.unique(p, axis=0, return_index=True)
    assert np.array_equal(first, fi)


def test_remove_duplicates_backends_on_carrier(env, golden_dir):
    """utils.remove_duplicates(backend=...) on the Open3D-shaped carrier == the reference's numpy
    and torch branches run on the same points (golden indices), attributes carried along."""
    from autodriver_pointcloud_preprocessor_b200 import geometry as o3d
    from autodriver_pointcloud_preprocessor_b200 import utils
    g = np.load(os.path.join(golden_dir, "dedup.npz"))
    p = g["points"]
    finite = np.isfinite(p).all(axis=1)
    for pts, want_np, want_torch in ((p, g["numpy_index"], None),
                                     (np.ascontiguousarray(p[finite]), None, None)):
        from oracle import dedup as odedup
        want_np = odedup.numpy_index(pts) if want_np is None else want_np
        want_torch = odedup.torch_compat_index(pts) if np.isfinite(pts).all() else odedup.torch_compat_index_numpy(pts)
        pcd = o3d.PointCloud()
        pcd.point["positions"] = o3d.Tensor(torch.from_numpy(pts).cuda())
        pcd.point["ring"] = o3d.Tensor(torch.arange(pts.shape[0], dtype=torch.int16).reshape(-1, 1).cuda())
        out, msg = utils.remove_duplicates(pcd, backend="numpy")
        assert msg == ""
        got = out.point.positions.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(pts[want_np]))
        assert np.array_equal(got.view(np.uint32)[~np.isnan(got)], pts[want_np].view(np.uint32)[~np.isnan(pts[want_np])])
        assert np.array_equal(out.point.ring.cpu().numpy().reshape(-1), want_np.astype(np.int16))
        out, msg = utils.remove_duplicates(pcd, backend="torch")
        assert len(out.point.positions) == pts.shape[0]                  # the reference's N-row result
        assert np.array_equal(out.point.ring.cpu().numpy().reshape(-1), want_torch.astype(np.int16))


@pytest.mark.parametrize("mode", ["numpy", "torch"])
@pytest.mark.parametrize("layout", ["xyzi16", "xyzirt22"])
def test_frontend_sorted_modes(env, mode, layout):
    from oracle import dedup as odedup
    from oracle import filters, pc2
    ctx, engine, synth, capi = env["ctx"], env["engine"], env["synth"], env["capi"]
    scan = synth.lidar_scan(seed=5, n_beams=32, n_az=512)
    scan["positions"][5] = [np.inf, 1.0, 2.0]
    scan["positions"][77] = [-0.0, 3.0, 2.0]
    scan["positions"][78] = [0.0, 3.0, 2.0]
    msg = synth.pack_cloud(scan, layout, is_dense=False)
    data = torch.frombuffer(bytearray(msg.data), dtype=torch.uint8).cuda()
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, data)
    crop = dict(min=[-20.1, -33.3, -1.7], max=[27.3, 19.9, 2.9], invert=False, mode=capi.CROP_TORCH)
    dm = capi.DEDUP_NUMPY if mode == "numpy" else capi.DEDUP_TORCH_COMPAT
    cfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=dm, remove_nan=True, remove_inf=True,
                                 transforms=[T_A], crop=crop)
    for _ in range(2):
        xyzi, src, _, cnt = ctx.frontend([desc], cfg, want_src=True, want_stage=False)
        ctx.check()
        arr = pc2.read_points(msg, skip_nans=False)
        pos = np.vstack((arr["x"], arr["y"], arr["z"])).T.astype(np.float32)
        nanskip = pc2.read_points_mask(msg, skip_nans=True)
        fn = odedup.numpy_index if mode == "numpy" else odedup.torch_compat_index_numpy
        p_ref, src_ref = filters.frontend_sorted(pos, fn, nanskip_mask=nanskip, remove_nan=True, remove_infinite=True,
                                                 transforms=[T_A], crop=crop)
        m = int(cnt.item())
        assert m == len(src_ref)
        assert np.array_equal(src[:m].cpu().numpy().astype(np.uint32), src_ref)
        assert np.array_equal(xyzi[:m, :3].cpu().numpy().view(np.uint32), p_ref.view(np.uint32))
        assert np.array_equal(xyzi[:m, 3].cpu().numpy(), arr["intensity"].astype(np.float32)[src_ref])


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("mode", ["numpy", "torch"])
def test_pipeline_sorted_modes(env, mode, graph):
    """Whole pipeline with the reference's default CPU back ends for duplicate removal, eager and as a
    replayed graph, against the oracle pipeline (bit-exact)."""
    from oracle import dedup as odedup
    from oracle import pipeline as opipe
    ctx, engine, synth, capi = env["ctx"], env["engine"], env["synth"], env["capi"]
    msg = synth.pack_cloud(synth.lidar_scan(seed=33, n_beams=32, n_az=1024), "xyzirt22")
    data = torch.frombuffer(bytearray(msg.data), dtype=torch.uint8).cuda()
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, data)
    crop = dict(min=[-60.0, -60.0, -20.0], max=[60.0, 60.0, 20.0], invert=False, mode=capi.CROP_NUMPY)
    dm = capi.DEDUP_NUMPY if mode == "numpy" else capi.DEDUP_TORCH_COMPAT
    fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=dm, remove_nan=True, remove_inf=True,
                                  transforms=[T_A], crop=crop)
    stages = dict(voxel_size=0.1, radius=dict(nb_points=4, radius=0.4),
                  ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=60, probability=0.99, seed=3))
    pcfg = engine.make_pipeline_cfg(fcfg, **stages)
    out = torch.zeros((msg.width, 4), device="cuda")
    counts = torch.zeros(8, dtype=torch.int32, device="cuda")
    plane = torch.zeros(8, dtype=torch.float64, device="cuda")
    if graph:
        g = ctx.capture_pipeline([desc], pcfg, out, counts, plane)
        out.zero_(); counts.zero_()
        for _ in range(3):
            ctx.launch_graph(g)
    else:
        ctx.pipeline_run([desc], pcfg, out, counts, plane)
    ctx.check()
    cfg = opipe.default_config()
    cfg.update(dedup_mode=odedup.DEDUP_NUMPY if mode == "numpy" else odedup.DEDUP_TORCH_COMPAT,
               transforms=[T_A], crop=crop, **stages)
    ref = opipe.preprocess(msg, cfg)
    c = counts.cpu().numpy()
    assert c[capi.CNT_STATUS] == 0 and c[capi.CNT_FILTERED] == ref["n_filtered"]
    n_out = int(c[capi.CNT_OUTPUT])
    assert n_out == ref["positions"].shape[0]
    got = out[:n_out].cpu().numpy()
    assert np.array_equal(got[:, :3].view(np.uint32), ref["positions"].view(np.uint32))
    assert np.array_equal(got[:, 3].view(np.uint32), ref["intensity"].view(np.uint32))

"""GPU parity tests: every stage of the CUDA path, called through the C-ABI, against the CPU
oracle on the same seeded inputs.  Integer / mask / index results are compared bit-exact;
floating-point tolerances are written next to each assertion."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def env():
    from autodriver_pointcloud_preprocessor_b200 import _capi, engine, synth
    ctx = engine.Context(max_points=300_000)
    yield dict(ctx=ctx, engine=engine, capi=_capi, synth=synth)
    ctx.close()


def same_f32(a, b):
    """Bit-exact float32 equality; NaNs compare equal to NaNs (CPU and GPU propagate different
    NaN payloads through arithmetic, which carries no information)."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def dev_bytes(msg):
    return torch.frombuffer(bytearray(msg.data), dtype=torch.uint8).cuda()


def small_scan(synth, seed=0, n_beams=32, n_az=512, **kw):
    return synth.lidar_scan(seed=seed, n_beams=n_beams, n_az=n_az, **kw)


def oracle_frontend(msg, cfg_kw, dedup=False):
    from oracle import dedup as odedup
    from oracle import filters, pc2
    arr = pc2.read_points(msg, skip_nans=False)
    pos = np.vstack((arr["x"], arr["y"], arr["z"])).T.astype(np.float32)
    meta = pc2.get_pointcloud_metadata(arr.dtype.names)
    inten = arr[meta["intensity_field_name"]].astype(np.float32) if meta["has_intensity"] else np.zeros(len(arr), np.float32)
    nanskip = pc2.read_points_mask(msg, skip_nans=cfg_kw.get("skip_nans", False))
    p, src, stage = filters.frontend(pos, nanskip_mask=nanskip, dedup_mask_fn=odedup.open3d_mask if dedup else None,
                                     remove_nan=cfg_kw.get("remove_nan", False),
                                     remove_infinite=cfg_kw.get("remove_inf", False),
                                     transforms=cfg_kw.get("transforms", ()), crop=cfg_kw.get("crop"))
    return p, inten[src], src, stage


T_A = np.array([[0.9986295, -0.0523360, 0.0, 1.5], [0.0523360, 0.9986295, 0.0, -0.25], [0.0, 0.0, 1.0, 1.8],
                [0.0, 0.0, 0.0, 1.0]])
T_B = np.array([[1.0, 0.0, 0.0, 0.1], [0.0, 0.9998477, -0.0174524, 0.0], [0.0, 0.0174524, 0.9998477, 0.05],
                [0.0, 0.0, 0.0, 1.0]])


@pytest.mark.parametrize("layout", ["xyzi16", "xyzirt22", "ouster48"])
@pytest.mark.parametrize("crop_mode,invert", [(0, False), (1, False), (2, False), (0, True), (2, True)])
def test_frontend_parity(env, layout, crop_mode, invert):
    ctx, engine, synth = env["ctx"], env["engine"], env["synth"]
    scan = small_scan(synth, seed=3)
    scan["positions"][5] = [np.inf, 1.0, 2.0]
    scan["positions"][77] = [1.0, -np.inf, 2.0]
    scan["intensity"][9] = np.nan                                  # read_points drops rows with a NaN in ANY field
    msg = synth.pack_cloud(scan, layout, is_dense=False)
    crop = dict(min=[-20.1, -33.3, -1.7], max=[27.3, 19.9, 2.9], invert=invert, mode=crop_mode)
    kw = dict(skip_nans=True, remove_nan=True, remove_inf=True, transforms=[T_A, T_B], crop=crop)
    data = dev_bytes(msg)
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width * msg.height, data)
    cfg = engine.make_filter_cfg(dedup_mode=env["capi"].DEDUP_OPEN3D, **kw)
    xyzi, src, stage, cnt = ctx.frontend([desc], cfg, want_src=True, want_stage=True)
    ctx.check()
    m = int(cnt.item())
    p_ref, i_ref, src_ref, stage_ref = oracle_frontend(msg, kw, dedup=True)
    assert np.array_equal(stage.cpu().numpy(), stage_ref)          # every per-stage mask, bit-exact
    assert m == len(src_ref)
    assert np.array_equal(src[:m].cpu().numpy().astype(np.uint32), src_ref)   # surviving indices, in order
    got = xyzi[:m].cpu().numpy()
    assert np.array_equal(got[:, :3].view(np.uint32), p_ref.view(np.uint32))  # float32 transform bit-exact
    assert np.array_equal(got[:, 3].view(np.uint32), i_ref.view(np.uint32))


def test_frontend_dense_flag_and_no_filters(env):
    """is_dense clouds keep their NaNs through read_points (utils.py:209 skip_nans && !is_dense)."""
    ctx, engine, synth = env["ctx"], env["engine"], env["synth"]
    msg = synth.pack_cloud(small_scan(synth, seed=4), "xyzi16", is_dense=True)
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, dev_bytes(msg))
    kw = dict(skip_nans=True and not msg.is_dense)
    xyzi, src, stage, cnt = ctx.frontend([desc], engine.make_filter_cfg(**kw), want_stage=True)
    ctx.check()
    assert int(cnt.item()) == msg.width
    p_ref, _, _, stage_ref = oracle_frontend(msg, kw)
    assert np.array_equal(stage.cpu().numpy(), stage_ref)
    assert np.array_equal(xyzi[:, :3].cpu().numpy().view(np.uint32), p_ref.view(np.uint32))


@pytest.mark.parametrize("n", [0, 1, 31, 1024, 1025, 4097])
def test_frontend_ragged_sizes(env, n):
    ctx, engine, synth = env["ctx"], env["engine"], env["synth"]
    scan = synth.lidar_scan(seed=11, n_beams=16, n_az=512, n_points=max(n, 1))
    if n == 0:
        scan = {k: v[:0] for k, v in scan.items()}
    msg = synth.pack_cloud(scan, "xyzirt22")
    data = dev_bytes(msg) if n else torch.zeros(16, dtype=torch.uint8, device="cuda")
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, n, data)
    kw = dict(skip_nans=True, remove_nan=True, remove_inf=True,
              crop=dict(min=[-30, -30, -3], max=[30, 30, 10], invert=False, mode=2))
    xyzi, src, stage, cnt = ctx.frontend([desc], engine.make_filter_cfg(**kw), want_stage=True)
    ctx.check()
    p_ref, _, src_ref, _ = oracle_frontend(msg, kw)
    m = int(cnt.item())
    assert m == len(src_ref)
    assert np.array_equal(src[:m].cpu().numpy().astype(np.uint32), src_ref)


def test_generic_layout_datatypes(env, golden_dir):
    """x/y/z as float64, integer intensity, unaligned records: the casts of utils.py:102-121."""
    import json
    from oracle import pc2
    from autodriver_pointcloud_preprocessor_b200.msgs import PointCloud2, PointField
    ctx, engine = env["ctx"], env["engine"]
    g = np.load(os.path.join(golden_dir, "convert.npz"))
    meta = json.load(open(os.path.join(golden_dir, "metadata.json")))
    np2pf = {"<f4": 7, "<f8": 8, "|u1": 2, "<u2": 4, "<u4": 6, "|i1": 1, "<i2": 3, "<i4": 5}
    for name in ("velodyne", "autoware", "livox", "f64xyz"):
        spec = meta[name]
        fields = [PointField(name=f[0], offset=f[2], datatype=np2pf[f[1]], count=1) for f in spec["fields"]]
        raw = g[f"{name}__bytes"]
        n = raw.size // spec["itemsize"]
        msg = PointCloud2(width=n, fields=fields, point_step=spec["itemsize"], data=raw.tobytes(), is_dense=True)
        desc = engine.make_cloud_desc(fields, spec["itemsize"], n, torch.from_numpy(raw.copy()).cuda())
        xyzi = ctx.unpack(desc).cpu().numpy()
        assert np.array_equal(xyzi[:, :3].view(np.uint32), g[f"{name}__positions"].view(np.uint32)), name
        assert np.array_equal(xyzi[:, 3], g[f"{name}__intensity"]), name     # reference's own conversion


def test_concat_multi_sensor(env):
    """Config C3 in small: 4 sensors, per-sensor extrinsics, one launch, then a common crop."""
    from oracle import pipeline as opipe
    ctx, engine, synth = env["ctx"], env["engine"], env["synth"]
    T = synth.sensor_extrinsics(4)
    scans = [small_scan(synth, seed=20 + s, n_beams=16, n_az=300 + 37 * s) for s in range(4)]
    layouts = ["xyzi16", "xyzirt22", "ouster48", "xyzi16"]
    msgs = [synth.pack_cloud(sc, lay) for sc, lay in zip(scans, layouts)]
    datas = [dev_bytes(m) for m in msgs]
    descs = [engine.make_cloud_desc(m.fields, m.point_step, m.width, d, transform=T[s])
             for s, (m, d) in enumerate(zip(msgs, datas))]
    xyzi, src, _, cnt = ctx.frontend(descs, engine.make_filter_cfg(), want_src=True)
    ctx.check()
    ref = opipe.concat(scans, list(T))
    n = int(cnt.item())
    assert n == ref["positions"].shape[0]
    got = xyzi[:n].cpu().numpy()
    assert np.isnan(ref["positions"]).any()                       # NaN returns travel through the transform
    assert same_f32(got[:, :3], ref["positions"])
    assert same_f32(got[:, 3], ref["intensity"])
    assert np.array_equal(src[:n].cpu().numpy(), np.arange(n))


def filtered_cloud(env, seed=5, n_beams=32, n_az=1024):
    """A finite, cropped SoA cloud on the device + its numpy copy."""
    ctx, engine, synth = env["ctx"], env["engine"], env["synth"]
    msg = synth.pack_cloud(small_scan(synth, seed=seed, n_beams=n_beams, n_az=n_az), "xyzi16")
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, dev_bytes(msg))
    cfg = engine.make_filter_cfg(skip_nans=True, remove_nan=True, remove_inf=True,
                                 crop=dict(min=[-60, -60, -20], max=[60, 60, 20], invert=False, mode=2))
    xyzi, _, _, cnt = ctx.frontend([desc], cfg, want_src=False)
    n = int(cnt.item())
    xyzi = xyzi[:n].contiguous()
    return xyzi, xyzi.cpu().numpy()


@pytest.mark.parametrize("voxel_size", [0.1, 0.5, 0.01])
def test_voxel_parity(env, voxel_size):
    from oracle import voxel as ovox
    ctx = env["ctx"]
    xyzi, host = filtered_cloud(env)
    for rep in range(2):                                            # twice: the table must self-clean
        out, p2v, vc, cnt = ctx.voxel_downsample(xyzi, voxel_size, want_p2v=True, want_counts=True)
        ctx.check()
        v = int(cnt.item())
        ref = ovox.voxel_down_sample(host[:, :3], voxel_size, host[:, 3], fixed=True)
        assert v == ref["positions"].shape[0]
        assert np.array_equal(p2v.cpu().numpy(), ref["p2v"])        # voxel membership, bit-exact
        assert np.array_equal(vc[:v].cpu().numpy().astype(np.uint32), ref["counts"])
        got = out[:v].cpu().numpy()
        assert np.array_equal(got[:, :3].view(np.uint32), ref["positions"].view(np.uint32))   # fixed-point mean
        assert np.array_equal(got[:, 3].view(np.uint32), ref["intensity"].view(np.uint32))
        # and within 1e-5 (relative to the coordinate scale) of Open3D's float32 serial sums
        o3d = ovox.centroids_o3d(host[:, :3], ref["p2v"], v)
        assert np.max(np.abs(got[:, :3] - o3d) / np.maximum(1.0, np.abs(o3d))) < 1e-5


def test_voxel_key_range_error(env):
    ctx = env["ctx"]
    xyzi = torch.tensor([[0.0, 0.0, 0.0, 1.0], [70000.0, 0.0, 0.0, 1.0]], device="cuda")
    ctx.voxel_downsample(xyzi, 0.1)
    with pytest.raises(env["capi"].ApcError) as e:
        ctx.check()
    assert e.value.code == env["capi"].APC_ERR_KEY_RANGE
    # the context recovers: tables were wiped
    out, _, _, cnt = ctx.voxel_downsample(xyzi[:1].contiguous(), 0.1)
    ctx.check()
    assert int(cnt.item()) == 1


def test_masks_select_gather(env, golden_dir):
    from oracle import dedup as odedup
    from oracle import filters
    ctx = env["ctx"]
    g = np.load(os.path.join(golden_dir, "crop.npz"))
    p = g["points"]
    xyzi = torch.from_numpy(np.concatenate([p, np.zeros((p.shape[0], 1), np.float32)], 1)).cuda()
    for case in ("roi", "frac"):
        lo, hi = g[f"{case}_min"].tolist(), g[f"{case}_max"].tolist()
        for backend, mode in (("numpy", 0), ("torch", 1)):
            for invert in (False, True):
                m = ctx.crop_mask(xyzi, lo, hi, mode=mode, invert=invert).cpu().numpy().astype(bool)
                assert np.array_equal(m, g[f"{case}_{backend}_{'inv' if invert else 'fwd'}"])   # reference golden
        for invert in (False, True):
            m = ctx.crop_mask(xyzi, lo, hi, mode=2, invert=invert).cpu().numpy().astype(bool)
            assert np.array_equal(m, filters.crop_mask(p, lo, hi, invert, filters.CROP_OPEN3D))
    for rn, ri in ((True, True), (True, False), (False, True)):
        m = ctx.non_finite_mask(xyzi, rn, ri).cpu().numpy().astype(bool)
        assert np.array_equal(m, filters.non_finite_mask(p, rn, ri))
    d = np.load(os.path.join(golden_dir, "dedup.npz"))["points"]
    dx = torch.from_numpy(np.concatenate([d, np.zeros((d.shape[0], 1), np.float32)], 1)).cuda()
    for rep in range(2):
        m = ctx.duplicate_mask(dx).cpu().numpy().astype(bool)
        assert np.array_equal(m, odedup.open3d_mask(d))
    mask = torch.from_numpy(odedup.open3d_mask(d).astype(np.uint8)).cuda()
    out, idx, cnt = ctx.select_by_mask(dx, mask)
    k = int(cnt.item())
    want = np.flatnonzero(odedup.open3d_mask(d))
    assert np.array_equal(idx[:k].cpu().numpy(), want)
    assert np.array_equal(out[:k].cpu().numpy()[:, :3].view(np.uint32), d[want].view(np.uint32))
    out2, idx2, cnt2 = ctx.select_by_mask(dx, mask, invert=True)
    assert int(cnt2.item()) == d.shape[0] - k
    ring = torch.arange(d.shape[0], dtype=torch.int16, device="cuda")
    assert np.array_equal(ctx.gather(ring, idx[:k].contiguous()).cpu().numpy(), want.astype(np.int16))
    t64 = torch.arange(d.shape[0], dtype=torch.float64, device="cuda") * 0.5
    assert np.array_equal(ctx.gather(t64, idx[:k].contiguous()).cpu().numpy(), want * 0.5)
    T = T_A
    tx = ctx.transform(dx, T).cpu().numpy()
    ref = filters.transform(d, T)
    assert same_f32(tx[:, :3], ref)


def voxelised(env, voxel_size=0.1, **kw):
    ctx = env["ctx"]
    xyzi, _ = filtered_cloud(env, **kw)
    out, _, _, cnt = ctx.voxel_downsample(xyzi, voxel_size)
    ctx.check()
    v = out[:int(cnt.item())].contiguous()
    return v, v.cpu().numpy()


def test_radius_outliers_parity(env):
    from oracle import outliers
    ctx = env["ctx"]
    v, host = voxelised(env, 0.1)
    for nb, r in ((5, 0.5), (3, 0.25), (12, 1.0)):
        for rep in range(2):
            mask, counts = ctx.radius_outliers(v, nb, r, want_counts=True)
            ctx.check()
        ref = outliers.radius_mask(host[:, :3], nb, r)
        assert np.array_equal(mask.cpu().numpy().astype(bool), ref)          # outlier decisions, bit-exact
    # neighbour counts against the exhaustive float32 oracle on a subset small enough for N^2
    sub = v[:3000].contiguous()
    mask, counts = ctx.radius_outliers(sub, 5, 0.5, want_counts=True)
    ctx.check()
    assert np.array_equal(counts.cpu().numpy().astype(np.uint32), outliers.radius_counts_brute(host[:3000, :3], 0.5))
    mask2, _ = ctx.radius_outliers(sub, 5, 0.5, want_counts=False)             # early-exit path
    assert np.array_equal(mask.cpu().numpy(), mask2.cpu().numpy())


@pytest.mark.parametrize("centre", [(0.0, 0.0, 0.0), (30000.0, -20000.0, 5.0), (-65000.0, 65000.0, -100.0)])
def test_radius_query_box_pruning_is_exact(env, centre):
    """Cells of 2r with box-pruned neighbour cells (k_radius_query_oct): exact counts and keep / drop
    decisions equal the exhaustive float32 count for a dense random blob - every face / edge / corner case
    of the pruning - also tens of kilometres from the origin, where a float32 ulp of the coordinates
    (4 mm at 65 km) approaches the pruning slack and the query falls back to the 27-cell walk."""
    from oracle import outliers
    ctx = env["ctx"]
    rng = np.random.default_rng(17)
    pts = np.zeros((4000, 4), dtype=np.float32)
    pts[:, :3] = (np.asarray(centre) + rng.uniform(-2.0, 2.0, size=(4000, 3))).astype(np.float32)
    dev = torch.from_numpy(pts).cuda()
    for nb, r in ((5, 0.5), (9, 0.35), (40, 1.0)):
        ref = outliers.radius_counts_brute(pts[:, :3], r)
        mask, counts = ctx.radius_outliers(dev, nb, r, want_counts=True)
        ctx.check()
        assert np.array_equal(counts.cpu().numpy().astype(np.uint32), ref)
        mask2, _ = ctx.radius_outliers(dev, nb, r, want_counts=False)         # early exit, scan outwards from the query
        ctx.check()
        assert np.array_equal(mask2.cpu().numpy().astype(bool), ref >= nb)
        assert np.array_equal(mask.cpu().numpy(), mask2.cpu().numpy())


def test_statistical_outliers_parity(env):
    from oracle import outliers
    ctx = env["ctx"]
    v, host = voxelised(env, 0.1)
    # a few isolated far points exercise the coarse levels and the brute-force straggler path
    far = torch.tensor([[300.0, 300.0, 50.0, 0.0], [-250.0, 40.0, 90.0, 0.0], [1000.0, -900.0, 5.0, 0.0]], device="cuda")
    v = torch.cat([v, far]).contiguous()
    host = v.cpu().numpy()
    for k, ratio in ((20, 2.0), (8, 1.0), (32, 2.0), (33, 1.5), (64, 2.0)):     # k <= 32: register lists; above: shared buffer
        for rep in range(2):
            mask, avg, stats = ctx.statistical_outliers(v, k, ratio)
            ctx.check()
        ref_mask, ref_avg = outliers.statistical_mask(host[:, :3], k, ratio)
        assert np.array_equal(avg.cpu().numpy().view(np.uint32), ref_avg.view(np.uint32))   # float32 KNN means
        mu, sigma, thr = outliers.statistical_threshold(ref_avg, ratio)
        st = stats.cpu().numpy()
        assert st[0] == mu and st[1] == sigma and st[2] == thr                # float64 tree reduction
        assert np.array_equal(mask.cpu().numpy().astype(bool), ref_mask)
    # tiny clouds: fewer points than k, P == 1
    for n in (1, 2, 7):
        tiny = v[:n].contiguous()
        mask, avg, _ = ctx.statistical_outliers(tiny, 20, 2.0)
        ctx.check()
        ref_mask, ref_avg = outliers.statistical_mask(host[:n, :3], 20, 2.0, brute=True)
        assert np.array_equal(avg.cpu().numpy().view(np.uint32), ref_avg.view(np.uint32))
        assert np.array_equal(mask.cpu().numpy().astype(bool), ref_mask)


@pytest.mark.parametrize("ransac_n,iters", [(3, 64), (5, 100), (8, 250)])
def test_segment_plane_parity(env, ransac_n, iters):
    from oracle import ransac
    ctx = env["ctx"]
    v, host = voxelised(env, 0.1)
    plane8, mask, info = ctx.segment_plane(v, 0.2, ransac_n, iters, 0.99, seed=1234)
    ctx.check()
    ref_plane, ref_inl, ri = ransac.segment_plane(host[:, :3], 0.2, ransac_n, iters, 0.99, seed=1234)
    info = info.cpu().numpy()
    # every hypothesis' tallies: float64 inlier count and the integer error term, bit-exact
    got_scores = ctx.segment_plane_scores(iters).cpu().numpy()
    valid = np.any(ri["planes"] != 0.0, axis=1)
    ref_scores = np.array([s if ok else (0, 0) for s, ok in zip(ri["scores"], valid)], dtype=np.int64)
    assert np.array_equal(got_scores[valid], ref_scores[valid])
    assert int(info[0]) == ri["best_it"]                                    # same winning hypothesis
    p8 = plane8.cpu().numpy()
    assert np.array_equal(p8[4:], ri["best_plane"])                         # float64 hypothesis fit, bit-exact
    got_inl = np.flatnonzero(mask.cpu().numpy())
    assert np.array_equal(got_inl, ref_inl)                                 # same inlier set
    assert int(info[1]) == ref_inl.size
    assert np.allclose(p8[:4], ref_plane, rtol=0, atol=1e-5)                # refit plane within 1e-5
    # explicit sample table gives the same result as the generator that produced it
    plane8b, maskb, infob = ctx.segment_plane(v, 0.2, ransac_n, iters, 0.99, sample_table=ri["samples"])
    assert np.array_equal(plane8b.cpu().numpy(), p8) and np.array_equal(maskb.cpu().numpy(), mask.cpu().numpy())


def test_repack_roundtrip(env):
    """prepare_pointcloud: same field names/datatypes re-packed without padding; unknown fields zero."""
    from oracle import pc2
    ctx, engine, synth = env["ctx"], env["engine"], env["synth"]
    scan = small_scan(synth, seed=8, nan_frac=0.0)
    msg = synth.pack_cloud(scan, "ouster48")
    n = msg.width
    pos = torch.from_numpy(scan["positions"]).cuda()
    xyzi = torch.cat([pos, torch.from_numpy(scan["intensity"]).cuda()[:, None]], 1).contiguous()
    meta = pc2.get_pointcloud_metadata([f.name for f in msg.fields])
    names = [f.name for f in msg.fields]
    dts = [f.datatype for f in msg.fields]
    packed, step = pc2.packed_fields(names, dts)
    ring = torch.from_numpy(scan["ring"].astype(np.int16)).cuda()
    tfield = torch.from_numpy(scan["time"].astype(np.float64)).cuda()
    src_of = {"x": (1, None), "y": (2, None), "z": (3, None), meta["intensity_field_name"]: (4, None),
              meta["ring_field_name"]: (5, ring), meta["time_field_name"]: (5, tfield)}
    out_fields = [(off, dt, *src_of.get(name, (0, None))) for (name, off, dt) in packed]
    raw = ctx.repack(xyzi, out_fields, step).cpu().numpy()
    ctx.check()
    ref = pc2.repack(msg.fields, scan["positions"], {"intensity": scan["intensity"], "ring": scan["ring"],
                                                      "time": scan["time"].astype(np.float64)}, meta)
    assert ref.dtype.itemsize == step
    assert np.array_equal(raw[:n * step], np.frombuffer(ref.tobytes(), dtype=np.uint8))


PIPE_CASES = {
    "C1": dict(voxel_size=0.1, ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=7)),
    "C2": dict(voxel_size=0.1, radius=dict(nb_points=5, radius=0.5),
               ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=7)),
    "C4": dict(voxel_size=0.05, statistical=dict(nb_neighbors=20, std_ratio=2.0)),
    "crop_only": dict(voxel_size=0.0),
    # both neighbour grids (KNN levels + radius cells) built inside one run
    "all": dict(voxel_size=0.1, statistical=dict(nb_neighbors=8, std_ratio=1.5), radius=dict(nb_points=4, radius=0.4),
                ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=50, probability=0.99, seed=5)),
}


@pytest.mark.parametrize("case", list(PIPE_CASES))
@pytest.mark.parametrize("graph", [False, True])
def test_pipeline_parity(env, case, graph):
    """preprocess() end to end, eager and as a replayed CUDA graph, against the oracle pipeline."""
    from oracle import pipeline as opipe
    ctx, engine, synth, capi = env["ctx"], env["engine"], env["synth"], env["capi"]
    stages = PIPE_CASES[case]
    msg = synth.pack_cloud(small_scan(synth, seed=31, n_beams=32, n_az=1024), "xyzirt22")
    data = dev_bytes(msg)
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, data)
    crop = dict(min=[-60.0, -60.0, -20.0], max=[60.0, 60.0, 20.0], invert=False, mode=capi.CROP_OPEN3D)
    fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                                  transforms=[T_A], crop=crop)
    pcfg = engine.make_pipeline_cfg(fcfg, **stages)
    out = torch.zeros((msg.width, 4), device="cuda")
    counts = torch.zeros(8, dtype=torch.int32, device="cuda")
    plane = torch.zeros(8, dtype=torch.float64, device="cuda")
    if graph:
        g = ctx.capture_pipeline([desc], pcfg, out, counts, plane)
        out.zero_(); counts.zero_()
        for _ in range(3):                                        # replays must be idempotent
            ctx.launch_graph(g)
    else:
        ctx.pipeline_run([desc], pcfg, out, counts, plane)
    ctx.check()
    cfg = opipe.default_config()
    cfg.update(transforms=[T_A], crop=crop, voxel_size=stages.get("voxel_size", 0.0),
               statistical=stages.get("statistical"), radius=stages.get("radius"), ground=stages.get("ground"))
    ref = opipe.preprocess(msg, cfg)
    c = counts.cpu().numpy()
    assert c[capi.CNT_STATUS] == 0
    assert c[capi.CNT_INPUT] == msg.width and c[capi.CNT_FILTERED] == ref["n_filtered"]
    n_out = int(c[capi.CNT_OUTPUT])
    assert n_out == ref["positions"].shape[0]
    got = out[:n_out].cpu().numpy()
    assert np.array_equal(got[:, :3].view(np.uint32), ref["positions"].view(np.uint32))     # whole pipeline bit-exact
    assert np.array_equal(got[:, 3].view(np.uint32), ref["intensity"].view(np.uint32))
    if stages.get("ground"):
        assert int(c[capi.CNT_GROUND_INLIERS]) == ref["ground_inliers"].size
        assert np.allclose(plane.cpu().numpy()[:4], ref["plane"], rtol=0, atol=1e-5)


def test_pipeline_index_maps(env):
    """apc_pipeline_run_maps: src_idx / p2v / voxel_counts / out_row against the oracle pipeline's
    intermediates (index work: bit-exact), with and without selection stages after the voxel stage."""
    from oracle import pipeline as opipe
    ctx, engine, synth, capi = env["ctx"], env["engine"], env["synth"], env["capi"]
    msg = synth.pack_cloud(small_scan(synth, seed=37, n_beams=32, n_az=1024), "xyzirt22")
    data = dev_bytes(msg)                                   # keep the device buffer alive: desc holds a raw pointer
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, data)
    crop = dict(min=[-60.0, -60.0, -20.0], max=[60.0, 60.0, 20.0], invert=False, mode=capi.CROP_OPEN3D)
    fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                                  transforms=[T_A], crop=crop)
    cases = [dict(voxel_size=0.1),
             dict(voxel_size=0.1, radius=dict(nb_points=4, radius=0.4)),
             dict(voxel_size=0.1, statistical=dict(nb_neighbors=8, std_ratio=1.5), radius=dict(nb_points=4, radius=0.4),
                  ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=50, probability=0.99, seed=5)),
             dict(ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=50, probability=0.99, seed=5))]
    for stages in cases:
        out, counts, plane, maps = ctx.pipeline_run_maps([desc], engine.make_pipeline_cfg(fcfg, **stages))
        ctx.check()
        cfg = opipe.default_config()
        cfg.update(transforms=[T_A], crop=crop, voxel_size=stages.get("voxel_size", 0.0), statistical=stages.get("statistical"),
                   radius=stages.get("radius"), ground=stages.get("ground"))
        ref = opipe.preprocess(msg, cfg)
        c = counts.cpu().numpy()
        m, n_out = int(c[capi.CNT_FILTERED]), int(c[capi.CNT_OUTPUT])
        assert np.array_equal(maps["src_idx"][:m].cpu().numpy().astype(np.uint32), ref["src_idx"])
        rows = np.arange(m)
        if stages.get("voxel_size"):
            v = int(c[capi.CNT_VOXELS])
            assert np.array_equal(maps["p2v"][:m].cpu().numpy(), ref["p2v"])
            assert np.array_equal(maps["voxel_counts"][:v].cpu().numpy(), ref["voxel_counts"])
            rows = np.arange(v)
        if stages.get("statistical"):
            rows = rows[ref["statistical_mask"]]
        if stages.get("radius"):
            rows = rows[ref["radius_mask"]]
        if stages.get("ground"):
            keep = np.ones(rows.shape[0], dtype=bool)
            keep[ref["ground_inliers"]] = False
            rows = rows[keep]
        assert n_out == rows.shape[0] == ref["positions"].shape[0]
        assert np.array_equal(maps["out_row"][:n_out].cpu().numpy(), rows)
        assert np.array_equal(out[:n_out, :3].cpu().numpy().view(np.uint32), ref["positions"].view(np.uint32))


@pytest.mark.parametrize("seed", list(range(12)))
def test_pipeline_random_stage_combinations(env, seed):
    """Seeded random combinations of layouts, duplicate-removal back ends, crop modes, transforms and
    stages (the stages share scratch: tables, scan states, counters, cursors) against the oracle
    pipeline, eager and replayed; everything bit-exact except the refit plane (1e-5)."""
    from oracle import dedup as odedup
    from oracle import pipeline as opipe
    ctx, engine, synth, capi = env["ctx"], env["engine"], env["synth"], env["capi"]
    rng = np.random.default_rng(1000 + seed)
    layout = ["xyzi16", "xyzirt22", "ouster48"][rng.integers(3)]
    scan = small_scan(synth, seed=200 + seed, n_beams=int(rng.choice([16, 32])), n_az=int(rng.choice([256, 512, 1024])))
    msg = synth.pack_cloud(scan, layout, is_dense=bool(rng.integers(2)) and False)
    data = dev_bytes(msg)
    desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, data)
    dedup = int(rng.choice([capi.DEDUP_OFF, capi.DEDUP_OPEN3D, capi.DEDUP_NUMPY, capi.DEDUP_TORCH_COMPAT]))
    crop = None
    if rng.integers(4):
        crop = dict(min=[-40.0, -35.5, -2.5], max=[38.0, 44.0, 6.0], invert=bool(rng.integers(4) == 0), mode=int(rng.integers(3)))
    transforms = [T_A, T_B][:int(rng.integers(3))]
    stages = {}
    if rng.integers(4):
        stages["voxel_size"] = float(rng.choice([0.1, 0.25, 0.5]))
    if rng.integers(2):
        stages["statistical"] = dict(nb_neighbors=int(rng.choice([4, 10, 20])), std_ratio=float(rng.choice([1.0, 2.0])))
    if rng.integers(2):
        stages["radius"] = dict(nb_points=int(rng.choice([2, 5])), radius=float(rng.choice([0.3, 0.6, 1.2])))
    if rng.integers(2):
        stages["ground"] = dict(distance_threshold=0.2, ransac_n=int(rng.choice([3, 5])), num_iterations=int(rng.choice([30, 100])),
                                probability=0.99, seed=int(rng.integers(100)))
    fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=dedup, remove_nan=True, remove_inf=True,
                                  transforms=transforms, crop=crop)
    pcfg = engine.make_pipeline_cfg(fcfg, **stages)
    cfg = opipe.default_config()
    cfg.update(dedup_mode={capi.DEDUP_OFF: odedup.DEDUP_OFF, capi.DEDUP_OPEN3D: odedup.DEDUP_OPEN3D,
                           capi.DEDUP_NUMPY: odedup.DEDUP_NUMPY, capi.DEDUP_TORCH_COMPAT: odedup.DEDUP_TORCH_COMPAT}[dedup],
               transforms=transforms, crop=crop, voxel_size=stages.get("voxel_size", 0.0),
               statistical=stages.get("statistical"), radius=stages.get("radius"), ground=stages.get("ground"))
    ref = opipe.preprocess(msg, cfg)
    out = torch.zeros((msg.width, 4), device="cuda")
    counts = torch.zeros(8, dtype=torch.int32, device="cuda")
    plane = torch.zeros(8, dtype=torch.float64, device="cuda")
    what = (layout, dedup, crop, len(transforms), stages)
    for mode in ("eager", "graph"):
        out.zero_(); counts.zero_()
        if mode == "graph":
            g = ctx.capture_pipeline([desc], pcfg, out, counts, plane)
            out.zero_(); counts.zero_()
            ctx.launch_graph(g)
            ctx.launch_graph(g)
        else:
            ctx.pipeline_run([desc], pcfg, out, counts, plane)
        ctx.check()
        c = counts.cpu().numpy()
        assert c[capi.CNT_STATUS] == 0 and c[capi.CNT_FILTERED] == ref["n_filtered"], what
        n_out = int(c[capi.CNT_OUTPUT])
        assert n_out == ref["positions"].shape[0], (mode, what)
        got = out[:n_out].cpu().numpy()
        assert same_f32(got[:, :3], ref["positions"]), (mode, what)
        assert same_f32(got[:, 3], ref["intensity"]), (mode, what)
        if stages.get("ground") and ref["ground_inliers"].size >= 3:
            assert np.allclose(plane.cpu().numpy()[:4], ref["plane"], rtol=0, atol=1e-5), (mode, what)


@pytest.mark.parametrize("voxel_size", [0.1, 0.37])
def test_voxel_sorted_equals_hash(env, voxel_size):
    """The sort-based voxel grid (onesweep radix sort + segmented reduce, the alternative north_star (3)
    names) produces the same voxels as the hash: identical counts and bit-identical centroids, in
    ascending (ix, iy, iz) order instead of first-occurrence order."""
    ctx = env["ctx"]
    xyzi, pts = filtered_cloud(env)
    out_h, p2v, vc_h, cnt_h = ctx.voxel_downsample(xyzi, voxel_size, want_p2v=True, want_counts=True)
    out_s, vc_s, cnt_s = ctx.voxel_downsample_sorted(xyzi, voxel_size, want_counts=True)
    ctx.check()
    v = int(cnt_h.item())
    assert v == int(cnt_s.item()) and v > 100
    h = np.concatenate([out_h[:v].cpu().numpy().view(np.uint32), vc_h[:v].cpu().numpy().view(np.uint32)[:, None]], 1)
    s = np.concatenate([out_s[:v].cpu().numpy().view(np.uint32), vc_s[:v].cpu().numpy().view(np.uint32)[:, None]], 1)
    # same multiset of (centroid bits, count) rows
    assert np.array_equal(h[np.lexsort(h.T[::-1])], s[np.lexsort(s.T[::-1])])
    # the sorted path's order: strictly ascending voxel index, x-major.  The index of a voxel is that of
    # its member points (floor(x / voxel_size) in float32, the kernels' arithmetic), looked up through
    # the hash path's point -> voxel map; rows are matched by their centroid bits.
    q = np.floor(pts[:, :3] / np.float32(voxel_size)).astype(np.int64)
    key_of_point = ((q[:, 0] + (1 << 20)) << 42) | ((q[:, 1] + (1 << 20)) << 21) | (q[:, 2] + (1 << 20))
    key_of_row = np.zeros(v, dtype=np.int64)
    key_of_row[p2v.cpu().numpy()[:pts.shape[0]]] = key_of_point
    by_bits = {tuple(r): k for r, k in zip(map(tuple, h), key_of_row)}
    assert len(by_bits) == v                                     # centroid + count identify a voxel here
    keys_sorted = np.array([by_bits[tuple(r)] for r in s])
    assert np.all(np.diff(keys_sorted) > 0)

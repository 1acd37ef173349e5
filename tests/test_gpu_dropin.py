"""GPU tests of the drop-in layer: the Open3D-shaped carrier, the utils functions and the node
callback (fused and staged paths) against the oracle and the reference-minted goldens."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

T_A = np.array([[0.9986295, -0.0523360, 0.0, 1.5], [0.0523360, 0.9986295, 0.0, -0.25], [0.0, 0.0, 1.0, 1.8],
                [0.0, 0.0, 0.0, 1.0]])


def scan_msg(layout="xyzirt22", seed=41, n_beams=32, n_az=512, **kw):
    from autodriver_pointcloud_preprocessor_b200 import synth
    scan = synth.lidar_scan(seed=seed, n_beams=n_beams, n_az=n_az, **kw)
    return scan, synth.pack_cloud(scan, layout, frame_id="lidar")


def test_pointcloud_to_dict_matches_reference_conversion():
    """utils.pointcloud_to_dict on the GPU == read_points + convert_pointcloud_to_numpy."""
    from autodriver_pointcloud_preprocessor_b200 import utils
    from oracle import pc2
    for layout in ("xyzi16", "xyzirt22", "ouster48"):
        scan, msg = scan_msg(layout)
        d, meta = utils.pointcloud_to_dict(msg, None, True, False, None)
        ref, ref_meta = pc2.pointcloud_to_dict(msg, None, True, False, None)
        assert {k: v for k, v in meta.items() if k != "header"} == {k: v for k, v in ref_meta.items() if k != "header"}
        assert set(d) == set(ref)
        for k in ref:
            got = d[k].cpu().numpy()
            assert got.dtype == ref[k].dtype, (layout, k)
            assert np.array_equal(got, ref[k], equal_nan=True), (layout, k)


def test_big_endian_message_is_byte_swapped_like_read_points():
    """read_points byte-swaps a message whose endianness differs from the host's (utils.py:206-211): the same
    records stored big-endian give the same carrier tensors, and the node publishes the same cloud."""
    from autodriver_pointcloud_preprocessor_b200 import utils
    from oracle import pc2
    for layout in ("xyzirt22", "ouster48"):
        scan, msg = scan_msg(layout, seed=57, n_beams=16, n_az=256)
        le = np.frombuffer(msg.data, dtype=pc2.dtype_from_fields(msg.fields, msg.point_step))
        be_dtype = np.dtype({"names": list(le.dtype.names), "formats": [le.dtype[n].newbyteorder(">") for n in le.dtype.names],
                             "offsets": [le.dtype.fields[n][1] for n in le.dtype.names], "itemsize": le.dtype.itemsize})
        be = np.zeros(le.shape, dtype=be_dtype)
        for name in le.dtype.names:
            be[name] = le[name]
        import copy
        msg_be = copy.copy(msg)
        msg_be.data, msg_be.is_bigendian = be.tobytes(), True
        assert msg_be.data != msg.data
        a, _ = utils.pointcloud_to_dict(msg, None, True, False, None)
        b, _ = utils.pointcloud_to_dict(msg_be, None, True, False, None)
        assert a.keys() == b.keys()
        for k in a:
            if k != "header":
                x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
                assert x.dtype == y.dtype and x.tobytes() == y.tobytes(), (layout, k)
        ref, _ = pc2.pointcloud_to_dict(msg_be, None, True, False, None)          # the oracle runs read_points' byteswap
        assert np.array_equal(b["positions"].cpu().numpy().view(np.uint32), ref["positions"].view(np.uint32))
        outs = []
        for m in (msg, msg_be):
            node = make_node({"use_gpu": True, "voxel_size": 0.2, "estimate_normals": False, "remove_ground": True,
                              "remove_ground.seed": 3})
            node.callback(m)
            assert len(node.pointcloud_pub.messages) == 1, "callback dropped the frame"
            outs.append(bytes(node.pointcloud_pub.messages[0].data))
        assert outs[0] == outs[1]


def test_carrier_ops_against_oracle(golden_dir):
    from autodriver_pointcloud_preprocessor_b200 import geometry as o3d
    from autodriver_pointcloud_preprocessor_b200 import utils
    from oracle import dedup as odedup
    from oracle import filters, voxel
    scan, msg = scan_msg("xyzirt22", nan_frac=0.01)
    d, meta = utils.pointcloud_to_dict(msg, None, False, False, None)       # keep NaNs: skip_nans False
    for key in ("intensity", "ring", "time"):
        d = utils.get_fields_from_dicts(key, d, meta)
    pcd = utils.dict_to_open3d_tensor_pointcloud(d, device="CUDA:0")
    pos = scan["positions"]
    # duplicates -> non finite -> transform -> crop, like pp.py:450-506
    p1, msg1 = utils.remove_duplicates(pcd, backend="open3d")
    m_dup = odedup.open3d_mask(pos)
    assert "remove_duplicated_points" in msg1 and len(p1.point.positions) == m_dup.sum()
    p2, mask2 = p1.remove_non_finite_points(remove_nan=True, remove_infinite=True)
    m_fin = filters.non_finite_mask(pos[m_dup])
    assert np.array_equal(mask2.cpu().numpy(), m_fin)
    keep = np.flatnonzero(m_dup)[m_fin]
    p2.transform(o3d.Tensor(T_A, dtype=o3d.float32))
    ref_pos = filters.transform(pos[keep], T_A)
    assert np.array_equal(p2.point.positions.cpu().numpy().view(np.uint32), ref_pos.view(np.uint32))
    for backend, mode in (("numpy", 0), ("torch", 1), ("open3d", 2)):
        for invert in (False, True):
            p3, cmsg = utils.crop_pointcloud(p2, backend=backend, min_bound=[-20.1, -33.3, -1.7],
                                             max_bound=[27.3, 19.9, 2.9], invert=invert)
            m = filters.crop_mask(ref_pos, [-20.1, -33.3, -1.7], [27.3, 19.9, 2.9], invert, mode)
            assert len(p3.point.positions) == m.sum(), (backend, invert)
            # every attribute travels with its point (select_by_mask gathers all of them)
            assert np.array_equal(p3.point.ring.cpu().numpy().reshape(-1), scan["ring"][keep][m])
            assert np.array_equal(p3.point.time.cpu().numpy().reshape(-1), scan["time"][keep][m].astype(np.float64))
            if backend == "open3d":
                assert "Using Open3D pointcloud.crop()" in cmsg
    p3, _ = utils.crop_pointcloud(p2, backend="open3d", min_bound=[-60, -60, -20], max_bound=[60, 60, 20])
    m = filters.crop_mask(ref_pos, [-60, -60, -20], [60, 60, 20])
    # voxel grid: positions/intensity fixed-point mean, other attributes float32 mean then cast back
    v = p3.voxel_down_sample(0.25)
    ref = voxel.voxel_down_sample(ref_pos[m], 0.25, scan["intensity"][keep][m], fixed=True)
    assert np.array_equal(v.point.positions.cpu().numpy().view(np.uint32), ref["positions"].view(np.uint32))
    assert np.array_equal(v.point.intensity.cpu().numpy().reshape(-1).view(np.uint32), ref["intensity"].view(np.uint32))
    # ring (uint16): exact integer sums, one divide, cast back - bit-equal to the oracle's fixed-point mean
    # AND to Open3D's float32 serial index_add (the sums stay far below 2^24)
    nv = ref["counts"].size
    ring_in = scan["ring"][keep][m].astype(np.float32)
    got_ring = v.point.ring.cpu().numpy().reshape(-1)
    assert got_ring.dtype == np.uint16
    assert np.array_equal(got_ring, voxel.centroids_fixed(ring_in, ref["p2v"], nv, scale=1.0).astype(np.uint16))
    assert np.array_equal(got_ring, voxel.centroids_o3d(ring_in, ref["p2v"], nv).astype(np.uint16))
    # time (float64 in the carrier, averaged in float32): order-independent 2^-20 fixed point, bit-equal to
    # the oracle; within 1e-6 of Open3D's float32 serial sum
    time_in = scan["time"][keep][m].astype(np.float32)
    got_time = v.point.time.cpu().numpy().reshape(-1)
    assert got_time.dtype == np.float64
    assert np.array_equal(got_time, voxel.centroids_fixed(time_in, ref["p2v"], nv, scale=float(1 << 20)).astype(np.float64))
    assert np.allclose(got_time, voxel.centroids_o3d(time_in, ref["p2v"], nv), rtol=0, atol=1e-6)
    # select_by_index(invert=True) == complement in order (pp.py:542)
    idx = o3d.Tensor(torch.tensor([0, 5, 7], dtype=torch.int64))
    rest = v.select_by_index(idx, invert=True)
    assert len(rest.point.positions) == len(v.point.positions) - 3
    assert np.array_equal(rest.point.positions.cpu().numpy()[:4], np.delete(ref["positions"], [0, 5, 7], axis=0)[:4])
    # CPU-resident carrier: same results, tensors stay on the host
    pc_cpu = p3.cpu()
    v_cpu = pc_cpu.voxel_down_sample(0.25)
    assert v_cpu.point.positions.is_cpu
    assert np.array_equal(v_cpu.point.positions.numpy(), v.point.positions.cpu().numpy())


def make_node(overrides, **kw):
    from autodriver_pointcloud_preprocessor_b200.pointcloud_preprocessor import PointcloudPreprocessorNode
    return PointcloudPreprocessorNode(parameter_overrides=overrides, **kw)


def oracle_published(msg, cfg_update, T=None):
    """What the reference would publish: oracle pipeline + prepare_pointcloud re-pack."""
    from oracle import pc2
    from oracle import pipeline as opipe
    cfg = opipe.default_config()
    cfg.update(cfg_update)
    if T is not None:
        cfg["transforms"] = [T]
    return opipe.preprocess(msg, cfg), cfg


def test_node_declares_reference_parameters(golden_dir):
    import json
    node = make_node({})
    ref = json.load(open(os.path.join(golden_dir, "node_contract.json")))["parameters"]
    for p in ref:
        assert node.has_parameter(p["name"]), p["name"]
        if p["name"] != "offset_pointcloud_matrix":
            assert node.get_parameter(p["name"]).value == p["default"]
    ns = make_node({}, parameter_namespace="front")
    assert ns.has_parameter("front.voxel_size") and ns.parameter_namespace == "front."


@pytest.mark.parametrize("layout,fused", [("xyzi16", "auto"), ("xyzirt22", "auto"), ("xyzirt22", "true"),
                                          ("ouster48", "false")])
def test_node_callback_end_to_end(layout, fused):
    """callback(): bytes in -> published PointCloud2 bytes out, fused and staged paths."""
    from autodriver_pointcloud_preprocessor_b200 import synth
    from oracle import pc2
    scan, msg = scan_msg(layout, seed=43, n_beams=32, n_az=1024)
    ground = dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=3)
    node = make_node({"use_gpu": True, "voxel_size": 0.1, "remove_ground": True, "remove_ground.seed": 3,
                      "remove_radius_outliers": True, "estimate_normals": False, "robot_frame": "base_link",
                      "fused_pipeline": fused})
    q = [0.0, 0.0, np.sin(0.02), np.cos(0.02)]
    node.tf_buffer.set_transform("base_link", "lidar", (1.5, -0.25, 1.8), q)
    node.callback(msg)
    assert len(node.pointcloud_pub.messages) == 1, "callback dropped the frame"
    out = node.pointcloud_pub.messages[0]
    T = node.camera_to_robot_tf.cpu().numpy()
    ref, cfg = oracle_published(msg, dict(voxel_size=0.1, ground=ground, radius=dict(nb_points=5, radius=0.5)), T)
    assert out.width == ref["positions"].shape[0] and out.height == 1
    assert out.header.frame_id == "base_link"                                   # pp.py:633-634
    assert [f.name for f in out.fields] == [f.name for f in msg.fields]         # same names / types, re-packed
    packed, step = pc2.packed_fields([f.name for f in msg.fields], [f.datatype for f in msg.fields])
    assert out.point_step == step and [(f.name, f.offset, f.datatype) for f in out.fields] == packed
    arr = np.frombuffer(out.data, dtype=pc2.dtype_from_fields(out.fields, out.point_step))
    got = np.stack([arr["x"], arr["y"], arr["z"]], 1)
    assert np.array_equal(got.view(np.uint32), ref["positions"].view(np.uint32))
    assert np.array_equal(arr["intensity"].view(np.uint32), ref["intensity"].view(np.uint32))
    if "ring" in arr.dtype.names:
        assert arr["ring"].any()                    # both paths carry / voxel-average every known attribute
    assert set(node.processing_times) >= {"ros_to_numpy", "preprocessing_time", "pointcloud_msg_parsing",
                                          "pointcloud_pub", "total_callback_time", "tf_lookup"}
    assert out.is_dense == (msg.is_dense and True)


def test_node_dynamic_parameters():
    from autodriver_pointcloud_preprocessor_b200._ros_compat import Parameter
    node = make_node({"use_gpu": True, "estimate_normals": False})
    ok = node.parameter_change_callback([Parameter("voxel_size", value=0.2)])
    assert ok.successful and node.voxel_size == 0.2
    assert not node.parameter_change_callback([Parameter("voxel_size", value="big")]).successful      # type guard
    assert not node.parameter_change_callback([Parameter("no_such_parameter", value=1)]).successful   # pp.py:1001
    assert not node.parameter_change_callback([Parameter("roi_min", value=[0.0, 1.0])]).successful    # length 3
    assert node.parameter_change_callback([Parameter("roi_min", value=[-5.0, -5.0, -1.0])]).successful
    assert node.roi_min == [-5.0, -5.0, -1.0]
    # quirks kept from the reference: neither of these can be set through the callback
    assert not node.parameter_change_callback([Parameter("offset_pointcloud_matrix", value=[1.0] * 16)]).successful
    assert not node.parameter_change_callback([Parameter("pointcloud_save_ascii", value=True)]).successful
    _, msg = scan_msg("xyzi16", seed=44)
    node.callback(msg)
    arr = node.pointcloud_pub.messages[-1]
    assert arr.width > 0


def test_concatenate_then_voxel():
    """Config C3 in small through the host API: 4 sensors merged in one launch, then voxel."""
    from autodriver_pointcloud_preprocessor_b200 import synth
    from autodriver_pointcloud_preprocessor_b200.pointcloud_concatenator import concatenate
    from oracle import pipeline as opipe
    from oracle import voxel
    T = synth.sensor_extrinsics(4)
    scans = [synth.lidar_scan(seed=60 + s, n_beams=16, n_az=512, nan_frac=0.0) for s in range(4)]
    msgs = [synth.pack_cloud(sc, "xyzi16") for sc in scans]
    xyzi, n, counts = concatenate(msgs, list(T), stages=dict(voxel_size=0.1))
    merged = opipe.concat(scans, list(T))
    ref = voxel.voxel_down_sample(merged["positions"], 0.1, merged["intensity"], fixed=True)
    assert n == ref["positions"].shape[0]
    assert np.array_equal(xyzi.cpu().numpy()[:, :3].view(np.uint32), ref["positions"].view(np.uint32))


def test_concatenator_node_publishes_the_merged_cloud():
    """Three sensors with different layouts and extrinsics through PointcloudConcatenatorNode: the
    published PointCloud2 (x, y, z, intensity, frame = target_frame, stamp = newest of the set) holds
    the oracle's concatenation bit for bit; a sensor without a transform is left out, not misplaced."""
    from autodriver_pointcloud_preprocessor_b200 import pointcloud_concatenator as pcn
    from autodriver_pointcloud_preprocessor_b200 import synth
    from autodriver_pointcloud_preprocessor_b200.msgs import Time
    from oracle import pc2
    from oracle import pipeline as opipe
    layouts = ["xyzi16", "xyzirt22", "ouster48"]
    scans = [synth.lidar_scan(seed=70 + s, n_beams=16, n_az=256, nan_frac=0.0) for s in range(3)]
    msgs = [synth.pack_cloud(sc, lay, frame_id=f"lidar{s}") for s, (sc, lay) in enumerate(zip(scans, layouts))]
    for s, m in enumerate(msgs):
        m.header.stamp = Time(100, 10_000_000 * s)
    node = pcn.PointcloudConcatenatorNode(parameter_overrides={
        "input_topics": ["/l0/points", "/l1/points", "/l2/points"], "target_frame": "base_link", "sync_mode": "sync"})
    quats = [(0.0, 0.0, np.sin(a / 2), np.cos(a / 2)) for a in (0.3, -1.2, 2.0)]
    trans = [(1.5, 0.0, 1.8), (-1.0, 0.5, 1.7), (0.0, -0.8, 2.0)]
    for s in range(3):
        node.tf_buffer.set_transform("base_link", f"lidar{s}", trans[s], quats[s])
    for s in range(3):
        node.subs[s][1](msgs[s])
    assert len(node.pointcloud_pub.messages) == 1
    out = node.pointcloud_pub.messages[0]
    assert out.header.frame_id == "base_link" and (out.header.stamp.sec, out.header.stamp.nanosec) == (100, 20_000_000)
    assert [(f.name, f.offset, f.datatype) for f in out.fields] == [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 12, 7)]
    Ts = [pcn._quat_to_matrix(trans[s], quats[s]).astype(np.float32) for s in range(3)]
    ref = opipe.concat(scans, Ts)
    arr = np.frombuffer(out.data, dtype=pc2.dtype_from_fields(out.fields, out.point_step))
    assert out.width == ref["positions"].shape[0]
    got = np.stack([arr["x"], arr["y"], arr["z"]], 1)
    assert np.array_equal(got.view(np.uint32), ref["positions"].view(np.uint32))
    assert np.array_equal(arr["intensity"].view(np.uint32), ref["intensity"].view(np.uint32))
    # a sensor whose frame is unknown to TF is skipped (robust), the others still go out
    msgs[1].header.frame_id = "unknown"
    node.sensor_tf[1] = None
    for s in range(3):
        node.subs[s][1](msgs[s])
    out2 = node.pointcloud_pub.messages[-1]
    assert out2.width == scans[0]["positions"].shape[0] + scans[2]["positions"].shape[0]


@pytest.mark.parametrize("layout,fused,backend", [("xyzi16", "auto", "open3d"), ("xyzirt22", "auto", "numpy"),
                                                  ("xyzirt22", "true", "torch")])
def test_node_callback_with_normals_and_reference_backends(layout, fused, backend):
    """The reference's defaults that the other node tests switch off: estimate_normals=True (adds
    normal_x/y/z to the published cloud, pp.py:560-567, 620-624) and the numpy / torch duplicate
    removal back ends (sorted unique rows / points[inverse], utils.py:520-542)."""
    from oracle import dedup as odedup
    from oracle import pc2
    scan, msg = scan_msg(layout, seed=47, n_beams=32, n_az=1024)
    ground = dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=3)
    node = make_node({"use_gpu": backend == "open3d", "cpu_backend": backend, "voxel_size": 0.1, "remove_ground": True,
                      "remove_ground.seed": 3, "estimate_normals.search_radius": 0.5, "robot_frame": "base_link",
                      "fused_pipeline": fused})
    assert node.estimate_normals is True                                        # reference default, pp.py:176
    q = [0.0, 0.0, np.sin(0.02), np.cos(0.02)]
    node.tf_buffer.set_transform("base_link", "lidar", (1.5, -0.25, 1.8), q)
    node.callback(msg)
    assert len(node.pointcloud_pub.messages) == 1, "callback dropped the frame"
    out = node.pointcloud_pub.messages[0]
    T = node.camera_to_robot_tf.cpu().numpy()
    mode = {"open3d": odedup.DEDUP_OPEN3D, "numpy": odedup.DEDUP_NUMPY, "torch": odedup.DEDUP_TORCH_COMPAT}[backend]
    crop_mode = {"open3d": 2, "numpy": 0, "torch": 1}[backend]
    ref, cfg = oracle_published(msg, dict(voxel_size=0.1, ground=ground, normals=dict(radius=0.5, max_nn=30),
                                          dedup_mode=mode,
                                          crop=dict(min=[-60.0, -60.0, -20.0], max=[60.0, 60.0, 20.0], invert=False,
                                                    mode=crop_mode)), T)
    names = [f.name for f in msg.fields] + ["normal_x", "normal_y", "normal_z"]
    assert [f.name for f in out.fields] == names
    packed, step = pc2.packed_fields(names, [f.datatype for f in msg.fields] + [7, 7, 7])
    assert out.point_step == step and [(f.name, f.offset, f.datatype) for f in out.fields] == packed
    assert out.width == ref["positions"].shape[0]
    arr = np.frombuffer(out.data, dtype=pc2.dtype_from_fields(out.fields, out.point_step))
    got = np.stack([arr["x"], arr["y"], arr["z"]], 1)
    assert np.array_equal(got.view(np.uint32), ref["positions"].view(np.uint32))
    nrm = np.stack([arr["normal_x"], arr["normal_y"], arr["normal_z"]], 1)
    assert np.allclose(np.linalg.norm(nrm.astype(np.float64), axis=1), 1.0, atol=1e-5)
    # float tolerance: 1e-5 per component wherever the eigenvector is well conditioned (two smallest
    # eigenvalues separated); EVERY normal is a unit eigenvector of the smallest eigenvalue of the oracle's
    # covariance (Rayleigh quotient within 1e-4 of the largest eigenvalue): the bound on the worst deviation
    from test_gpu_normals import check_normals
    check_normals(nrm, ref["normal_cov"], ref["normals"])
    assert "normal_estimation" in node.processing_times


def test_fused_and_staged_paths_publish_the_same_cloud():
    """Velodyne-style layout (x, y, z, intensity, ring u16, time f32): the fused pipeline (attributes
    carried through its index maps) and the staged carrier path publish the same message, byte for
    byte - the attribute voxel means are order-independent fixed-point sums in both paths - and the
    attributes equal the oracle's (``oracle.voxel.centroids_fixed`` over the oracle's own voxel map)."""
    from oracle import pc2
    scan, msg = scan_msg("xyzirt22", seed=49, n_beams=32, n_az=1024)
    outs = {}
    for fused in ("true", "false"):
        node = make_node({"use_gpu": True, "voxel_size": 0.1, "remove_ground": True, "remove_ground.seed": 3,
                          "remove_radius_outliers": True, "estimate_normals": False, "fused_pipeline": fused})
        node.callback(msg)
        assert len(node.pointcloud_pub.messages) == 1, "callback dropped the frame"
        out = node.pointcloud_pub.messages[0]
        outs[fused] = np.frombuffer(out.data, dtype=pc2.dtype_from_fields(out.fields, out.point_step))
    a, b = outs["true"], outs["false"]
    assert a.shape == b.shape and a.dtype == b.dtype
    assert a["ring"].any()
    assert a.tobytes() == b.tobytes()
    # against the oracle: pipeline with the same stages, attributes averaged over ITS point -> voxel map
    from oracle import pipeline as opipe
    from oracle import voxel
    cfg = opipe.default_config()
    cfg.update(voxel_size=0.1, radius=dict(nb_points=node.remove_radius_outliers_nb_points,
                                           radius=node.remove_radius_outliers_search_radius),
               ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=3))
    ref = opipe.preprocess(msg, cfg)
    assert a.shape[0] == ref["positions"].shape[0]
    assert np.array_equal(np.stack([a["x"], a["y"], a["z"]], 1).view(np.uint32), ref["positions"].view(np.uint32))
    nv = ref["voxel_counts"].size
    keep = ref["radius_mask"].copy()
    rows = np.flatnonzero(keep)
    ground = np.ones(rows.size, dtype=bool)
    ground[ref["ground_inliers"]] = False
    rows = rows[ground]                                            # voxel rows that reach the output, in order
    ring_in = scan["ring"][ref["src_idx"]].astype(np.float32)
    time_in = scan["time"][ref["src_idx"]].astype(np.float32)
    assert np.array_equal(a["ring"], voxel.centroids_fixed(ring_in, ref["p2v"], nv, scale=1.0).astype(np.uint16)[rows])
    assert np.array_equal(a["time"].view(np.uint32),
                          voxel.centroids_fixed(time_in, ref["p2v"], nv, scale=float(1 << 20))[rows].view(np.uint32))


def test_node_saves_published_cloud_as_pcd(tmp_path):
    """save_pointcloud (pp.py:1010-1018): the published records land in <dir>/<prefix><frame>.pcd."""
    from autodriver_pointcloud_preprocessor_b200 import pointcloud_loader as pl
    _, msg = scan_msg("xyzirt22", seed=51, n_beams=16, n_az=512)
    node = make_node({"use_gpu": True, "voxel_size": 0.2, "estimate_normals": False, "save_pointcloud": True,
                      "pointcloud_save_directory": str(tmp_path), "pointcloud_save_prepend_str": "scan_"})
    node.callback(msg)
    node.callback(msg)
    out = node.pointcloud_pub.messages[-1]
    files = sorted(os.listdir(tmp_path))
    assert files == ["scan_00000000.pcd", "scan_00000001.pcd"]
    back = pl.read_pcd(str(tmp_path / files[-1]))
    assert back.width == out.width and back.point_step == out.point_step
    assert [(f.name, f.offset, f.datatype) for f in back.fields] == [(f.name, f.offset, f.datatype) for f in out.fields]
    assert bytes(back.data) == bytes(out.data)


def test_rgb_cloud_through_the_fused_pipeline():
    """Colour clouds (x, y, z, packed float32 rgb; utils.py:110-119, pp.py:429-431, 598-603) take the fused
    pipeline too: the three channels are cut from the message bytes, scaled to float [0, 1], carried
    through the pipeline's index maps and re-packed.  Without voxels the fused and the staged carrier
    path publish the same bytes; with voxels the channel means agree to 1/255."""
    from autodriver_pointcloud_preprocessor_b200 import synth
    from autodriver_pointcloud_preprocessor_b200.msgs import Header, PointCloud2
    from oracle import pc2
    scan = synth.lidar_scan(seed=53, n_beams=16, n_az=512)
    n = scan["positions"].shape[0]
    rng = np.random.default_rng(5)
    rgb8 = rng.integers(0, 256, size=(n, 3), dtype=np.uint32)
    dt = np.dtype({"names": ["x", "y", "z", "rgb"], "formats": ["<f4"] * 4, "offsets": [0, 4, 8, 12], "itemsize": 16})
    arr = np.zeros(n, dtype=dt)
    arr["x"], arr["y"], arr["z"] = scan["positions"].T
    arr["rgb"] = ((rgb8[:, 0] << 16) | (rgb8[:, 1] << 8) | rgb8[:, 2]).astype(np.uint32).view(np.float32)
    msg = PointCloud2(header=Header(frame_id="lidar"), height=1, width=n, fields=synth.fields_from_dtype(dt),
                      is_bigendian=False, point_step=16, row_step=16 * n, data=arr.tobytes(), is_dense=False)
    for voxel_size in (0.0, 0.2):
        outs = {}
        for fused in ("true", "false"):
            node = make_node({"use_gpu": True, "voxel_size": voxel_size, "estimate_normals": False,
                              "remove_radius_outliers": True, "fused_pipeline": fused})
            node.callback(msg)
            assert len(node.pointcloud_pub.messages) == 1, "callback dropped the frame"
            assert node._use_fused() == (fused == "true")
            out = node.pointcloud_pub.messages[0]
            assert [f.name for f in out.fields] == ["x", "y", "z", "rgb"]
            outs[fused] = np.frombuffer(out.data, dtype=pc2.dtype_from_fields(out.fields, out.point_step))
        a, b = outs["true"], outs["false"]
        assert a.shape == b.shape and a.shape[0] > 1000
        for k in ("x", "y", "z"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
        ca, cb = a["rgb"].view(np.uint32), b["rgb"].view(np.uint32)
        assert ca.any()
        if voxel_size == 0.0:
            assert np.array_equal(ca, cb)
            # and they are the input's colours of the surviving points (order preserved)
            src = {}
            for p, c in zip(map(tuple, scan["positions"].view(np.uint32)), arr["rgb"].view(np.uint32)):
                src.setdefault(p, c)                         # duplicates: the lowest index survives
            got = np.stack([a["x"], a["y"], a["z"]], 1).view(np.uint32)
            assert all(src[tuple(p)] == c for p, c in zip(map(tuple, got[:500]), ca[:500]))
        else:
            for sh in (16, 8, 0):
                d = ((ca >> sh) & 0xFF).astype(np.int64) - ((cb >> sh) & 0xFF).astype(np.int64)
                assert np.abs(d).max() <= 1

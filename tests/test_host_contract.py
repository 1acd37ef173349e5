"""Drop-in contract of the host-side mirror (CPU only): parameter table, method names, utils
signatures and the pure-host helpers, all against fixtures minted from the reference
(tests/golden/make_golden.py), plus the frame sharding / gather logic on gloo."""
import inspect
import json
import os
import socket

import numpy as np
import pytest


@pytest.fixture(scope="module")
def contract(golden_dir):
    return json.load(open(os.path.join(golden_dir, "node_contract.json")))


def test_parameter_table_matches_reference(contract):
    from autodriver_pointcloud_preprocessor_b200 import pointcloud_preprocessor as pp
    ref = contract["parameters"]
    assert [p["name"] for p in ref] == [p[0] for p in pp.PARAMETERS]          # same names, same order
    for r, (name, default, ptype) in zip(ref, pp.PARAMETERS):
        if name == "offset_pointcloud_matrix":                                  # default is an expression: eye(4) flat
            assert default == np.eye(4).flatten().tolist()
        else:
            assert default == r["default"], name
            assert type(default) is type(r["default"]), name
        want = None if r["type"] is None else getattr(pp.ParameterType, r["type"])
        assert ptype == want, name
    # additive parameters never shadow a reference name
    assert not {p[0] for p in pp.EXTRA_PARAMETERS} & {p["name"] for p in ref}
    assert sorted(pp.PROCESSING_TIME_KEYS) == contract["processing_times_keys"]


def test_node_methods_match_reference(contract):
    from autodriver_pointcloud_preprocessor_b200 import pointcloud_preprocessor as pp
    cls = pp.PointcloudPreprocessorNode
    for m in contract["methods"]:
        assert hasattr(cls, m["name"]), m["name"]
        if m["name"] == "__init__":
            continue
        got = list(inspect.signature(getattr(cls, m["name"])).parameters)
        assert got[:len(m["args"])] == m["args"], m["name"]
    init = inspect.signature(cls.__init__).parameters
    assert list(init)[:4] == ["self", "node_name", "enabled", "parameter_namespace"]
    assert init["node_name"].default == "pointcloud_preprocessor" and init["enabled"].default is True
    assert callable(pp.main)


def test_utils_signatures_match_reference(golden_dir):
    from autodriver_pointcloud_preprocessor_b200 import utils
    sigs = json.load(open(os.path.join(golden_dir, "utils_signatures.json")))
    for name, args in sigs.items():
        fn = getattr(utils, name)
        params = inspect.signature(fn).parameters
        got = [p for p in params if not p.startswith("_")]
        assert got == [a[0] for a in args], name
        for (arg, default) in args:
            if default is not None:
                assert repr(params[arg].default) == default or str(params[arg].default) == default.strip("'\""), (name, arg)
    for const in ("FIELD_DTYPE_MAP", "FIELD_DTYPE_MAP_INV", "VENDOR_MAPPINGS"):
        assert hasattr(utils, const)


def test_host_helpers_match_reference(golden_dir):
    from autodriver_pointcloud_preprocessor_b200 import utils
    meta = json.load(open(os.path.join(golden_dir, "metadata.json")))
    for entry in meta["mappings"]:
        assert utils.get_pointcloud_metadata(entry["names"]) == entry["metadata"]
    for entry in meta["packed"]:
        fields, step = utils.numpy_struct_to_pointcloud2(entry["names"], entry["datatypes"])
        assert step == entry["point_step"]
        assert [[f.name, f.offset, f.datatype, f.count] for f in fields] == entry["fields"]
    g = np.load(os.path.join(golden_dir, "rgb.npz"))
    assert np.array_equal(utils.merge_rgb_fields(g["r"], g["g"], g["b"]).view(np.uint32), g["merged_float"].view(np.uint32))
    assert np.array_equal(utils.merge_rgb_fields(g["r"], g["g"], g["b"], return_int=True), g["merged_int"])
    assert np.array_equal(utils.extract_rgb_from_pointcloud(g["merged_float"]), g["extracted"])
    assert np.array_equal(utils.rgb_int_to_float(g["colors"]).view(np.uint32), g["packed"].view(np.uint32))
    assert np.array_equal(utils.rgb_to_intensity(g["colors"]), g["luminance"])
    c = np.load(os.path.join(golden_dir, "convert.npz"))
    for name in ("velodyne", "autoware", "livox", "f64xyz", "rgb"):
        spec = meta[name]
        dt = np.dtype({"names": [f[0] for f in spec["fields"]], "formats": [f[1] for f in spec["fields"]],
                       "offsets": [f[2] for f in spec["fields"]], "itemsize": spec["itemsize"]})
        arr = np.frombuffer(c[f"{name}__bytes"].tobytes(), dtype=dt)
        d = utils.convert_pointcloud_to_numpy(arr, dict(spec["metadata"], field_names=dt.names))
        for k, v in d.items():
            assert np.array_equal(v, c[f"{name}__{k}"], equal_nan=True) and v.dtype == c[f"{name}__{k}"].dtype


def test_dedup_backends_have_no_cpu_path():
    """numpy / torch duplicate removal run on the device (sorted unique rows); without a CUDA device
    they fail loudly instead of falling back to np.unique / torch.unique."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the gpu tests")
    from autodriver_pointcloud_preprocessor_b200 import geometry, utils
    pcd = geometry.PointCloud()
    pcd.point["positions"] = geometry.Tensor(torch.zeros((4, 3), dtype=torch.float32))
    for backend in ("numpy", "torch", "open3d"):
        with pytest.raises((RuntimeError, AssertionError)):
            utils.remove_duplicates(pcd, backend=backend)


def test_sensor_synchronizer():
    from autodriver_pointcloud_preprocessor_b200.msgs import Header, PointCloud2, Time
    from autodriver_pointcloud_preprocessor_b200.pointcloud_concatenator import SensorSynchronizer

    def msg(t):
        return PointCloud2(header=Header(stamp=Time(int(t), int((t % 1) * 1e9))))

    s = SensorSynchronizer(3, mode="sync", slop=0.05)
    assert s.add(0, msg(10.00)) is None and s.add(1, msg(10.01)) is None
    ids, msgs = s.add(2, msg(10.02))
    assert ids == [0, 1, 2] and len(msgs) == 3
    assert s.add(0, msg(10.10)) is None and s.add(1, msg(10.11)) is None
    assert s.add(2, msg(10.30)) is None                        # outside the slop: no set
    r = SensorSynchronizer(3, mode="robust", slop=0.05, timeout=0.2)
    assert r.add(0, msg(20.00)) is not None                    # only live sensor so far
    assert r.add(0, msg(20.10)) is not None
    r.add(1, msg(20.19))
    out = r.add(0, msg(20.20))                                 # sensor 2 never reported: not waited for
    assert out is not None and out[0] == [0, 1]


def test_concatenator_node_policies():
    """PointcloudConcatenatorNode (pointcloud_concatenator.py:1-5): parameters, one subscription per
    sensor, sets handed to publish_set by the two policies, per-sensor TF looked up with Duration / Time
    and cached.  No GPU: the merge itself is stubbed out here (tests/test_gpu_dropin.py runs it)."""
    from autodriver_pointcloud_preprocessor_b200 import pointcloud_concatenator as pcn
    from autodriver_pointcloud_preprocessor_b200.msgs import Header, PointCloud2, Time

    def msg(t, frame):
        return PointCloud2(header=Header(stamp=Time(int(t), int((t % 1) * 1e9)), frame_id=frame), width=1, height=1)

    sets = []
    Node = pcn.PointcloudConcatenatorNode
    node = Node(parameter_overrides={"input_topics": ["/a/points", "/b/points", "/c/points"], "target_frame": "base_link",
                                     "sync_mode": "sync", "slop": 0.05})
    node.publish_set = lambda ids, msgs: sets.append((ids, [m.header.frame_id for m in msgs]))
    assert [s[0] for s in node.subs] == ["/a/points", "/b/points", "/c/points"]
    assert node.pointcloud_pub.topic == "/lidar/points_concatenated"
    for i, fr in enumerate("abc"):
        node.subs[i][1](msg(5.0 + 0.01 * i, fr))               # the subscription callbacks
    assert sets == [([0, 1, 2], ["a", "b", "c"])]
    node.subs[0][1](msg(6.0, "a"))
    node.subs[1][1](msg(6.2, "b"))
    node.subs[2][1](msg(6.21, "c"))                             # a's message is outside the slop: no set
    assert len(sets) == 1
    robust = Node(parameter_overrides={"input_topics": ["/a/points", "/b/points"], "sync_mode": "robust", "timeout": 0.2})
    robust.publish_set = lambda ids, msgs: sets.append((ids, None))
    robust.subs[0][1](msg(1.0, "a"))
    robust.subs[0][1](msg(1.1, "a"))                            # sensor b is dead: a alone is published
    assert [s[0] for s in sets[1:]] == [[0], [0]]
    # TF: target == sensor frame needs none; otherwise one lookup, cached
    assert node.lookup_sensor_tf(0, "base_link") is None
    node.tf_buffer.set_transform("base_link", "a", (1.0, 2.0, 3.0), (0.0, 0.0, 0.0, 1.0))
    T = node.lookup_sensor_tf(0, "a", Time(1, 0))
    assert T.shape == (4, 4) and T[:3, 3].tolist() == [1.0, 2.0, 3.0] and node.lookup_sensor_tf(0, "a") is T
    with pytest.raises(Exception):
        node.lookup_sensor_tf(1, "b")                           # unknown frame: tf2 LookupException
    with pytest.raises(ValueError):
        Node(parameter_overrides={"input_topics": [f"/s{i}" for i in range(9)]})
    # quaternion -> matrix: 90 degrees about z
    R = pcn._quat_to_matrix((0, 0, 0), (0.0, 0.0, np.sin(np.pi / 4), np.cos(np.pi / 4)))
    assert np.allclose(R[:3, :3] @ [1.0, 0.0, 0.0], [0.0, 1.0, 0.0])


def test_shard_frames_covers_everything():
    from autodriver_pointcloud_preprocessor_b200.replay import shard_frames
    for n, w in ((1024, 8), (1024, 3), (5, 8), (0, 2)):
        got = [i for r in range(w) for i in shard_frames(n, w, r)]
        assert got == list(range(n))
        sizes = [len(shard_frames(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from autodriver_pointcloud_preprocessor_b200.replay import gather_outputs, shard_frames
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    frames = shard_frames(6, world, rank)
    send = torch.zeros((len(frames), 4, 4))
    counts = torch.zeros(len(frames), dtype=torch.int32)
    for j, f in enumerate(frames):
        counts[j] = f % 4 + 1
        send[j, :counts[j]] = float(f)
    recv, all_counts = gather_outputs(send, counts)
    q.put((rank, recv.numpy(), all_counts.numpy()))
    dist.destroy_process_group()


def test_gather_outputs_gloo_world2():
    """The multi-GPU result exchange on CPU: 2 ranks, gloo, frame-parallel shards."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, recv, counts in results:
        assert recv.shape == (2, 3, 4, 4) and counts.shape == (2, 3)
        for g in range(2):
            for j, f in enumerate(range(g * 3, g * 3 + 3)):
                assert counts[g, j] == f % 4 + 1
                assert (recv[g, j, :counts[g, j]] == f).all() and (recv[g, j, counts[g, j]:] == 0).all()


def test_pcd_loader_round_trip(tmp_path):
    """File replay shell (pointcloud_loader.py:1-5 states the intent only): binary and ascii PCD files
    become the PointCloud2 byte buffers a driver would publish, in name order, optionally looping."""
    from autodriver_pointcloud_preprocessor_b200 import pointcloud_loader as pl
    from autodriver_pointcloud_preprocessor_b200 import synth
    from oracle import pc2
    scan = synth.lidar_scan(seed=3, n_beams=8, n_az=64)
    msg = synth.pack_cloud(scan, "xyzirt22")
    for k, binary in enumerate((True, False)):
        pl.write_pcd(str(tmp_path / f"scan_{k}.pcd"), msg, binary=binary)
    replay = pl.DirectoryReplay(str(tmp_path))
    assert len(replay) == 2
    want = np.frombuffer(bytes(msg.data), dtype=pc2.dtype_from_fields(msg.fields, msg.point_step))
    for got_msg in replay:
        assert [(f.name, f.offset, f.datatype) for f in got_msg.fields] == [(f.name, f.offset, f.datatype) for f in msg.fields]
        assert got_msg.width == msg.width and got_msg.point_step == msg.point_step and not got_msg.is_dense
        got = np.frombuffer(bytes(got_msg.data), dtype=pc2.dtype_from_fields(got_msg.fields, got_msg.point_step))
        for name in want.dtype.names:
            assert np.array_equal(got[name], want[name], equal_nan=True), name      # repr() round-trips floats exactly
    looping = iter(pl.DirectoryReplay(str(tmp_path), loop=True))
    assert [next(looping).width for _ in range(5)] == [msg.width] * 5


def test_tf_lookup_uses_duration_and_time_like_the_reference():
    """pp.py:476-477,714-719: the stamp goes through rclpy.time.Time.from_msg and the timeout through
    rclpy.duration.Duration; tf2_ros.Buffer adds them, so a bare float raises TypeError there.  The
    stand-in buffer enforces the same signature, and the node must call it accordingly."""
    from autodriver_pointcloud_preprocessor_b200 import _ros_compat as rc
    from autodriver_pointcloud_preprocessor_b200 import msgs
    from autodriver_pointcloud_preprocessor_b200 import pointcloud_preprocessor as pp
    if rc.HAVE_ROS:
        pytest.skip("real tf2_ros present")
    buf = rc.Buffer()
    buf.set_transform("base_link", "lidar", (1.0, 2.0, 3.0), (0.0, 0.0, 0.0, 1.0))
    with pytest.raises(TypeError):
        buf.lookup_transform("base_link", "lidar", rc.Time(), 0.1)
    with pytest.raises(TypeError):
        buf.lookup_transform("base_link", "lidar", msgs.Header().stamp, rc.Duration(seconds=0.1))
    assert buf.lookup_transform("base_link", "lidar", rc.Time.from_msg(msgs.Header().stamp),
                                rc.Duration(seconds=0.1)).transform.translation.x == 1.0

    class Fake:                                       # just enough of the node for get_camera_to_robot_tf
        camera_to_robot_tf, static_camera_to_robot_tf = None, True
        robot_frame, transform_timeout, tf_buffer = "base_link", 0.1, buf

        def transform_to_matrix(self, t):
            return ("matrix", t.transform.translation.y)

    node = Fake()
    pp.PointcloudPreprocessorNode.get_camera_to_robot_tf(node, "lidar", rc.Time.from_msg(msgs.Header().stamp))
    assert node.camera_to_robot_tf == ("matrix", 2.0)
    node.camera_to_robot_tf = None
    pp.PointcloudPreprocessorNode.get_camera_to_robot_tf(node, "lidar")      # timestamp=None -> Time()
    assert node.camera_to_robot_tf == ("matrix", 2.0)

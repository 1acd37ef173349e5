"""Multi-LiDAR concatenation.

The reference's ``pointcloud_concatenator.py`` is a five-line statement of intent
(``pointcloud_concatenator.py:1-5``): concatenate and synchronise n point clouds into one,
optionally transformed to a target frame, with a message_filters-style synchronised mode and
a "robust" mode that publishes even if some sensors fail.  This module defines that behaviour:

* the merge itself is ONE launch of the fused front end over up to 8 PointCloud2 byte
  buffers of possibly different layouts - per-sensor float32 4x4 extrinsic, then the common
  filters / transforms / crop - output in sensor order with each sensor's point order kept
  (``oracle/pipeline.py:concat`` is the CPU statement of the same rule);
* ``SensorSynchronizer`` implements the two host-side policies on message stamps.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi, engine, geometry


def concatenate(clouds, transforms=None, filter_kw=None, stages=None, want_src=False):
    """Merge PointCloud2 messages (``clouds``) into one device cloud.

    ``transforms``: per-sensor 4x4 (or None) mapping each sensor frame to the target frame.
    ``filter_kw``: optional :func:`engine.make_filter_cfg` arguments applied to the merged
    cloud in the same launch; ``stages``: optional :func:`engine.make_pipeline_cfg` stages
    (voxel / outliers / ground) to run after the merge.
    Returns ``(xyzi[n,4] device tensor, n, src_idx | counts)``.
    """
    if not 1 <= len(clouds) <= _capi.APC_MAX_CLOUDS:
        raise ValueError(f"1..{_capi.APC_MAX_CLOUDS} sensors per launch")
    transforms = transforms or [None] * len(clouds)
    total = sum(c.width * c.height for c in clouds)
    ctx = geometry.get_context(total)
    bufs, descs = [], []
    for c, T in zip(clouds, transforms):
        n = c.width * c.height
        raw = (torch.frombuffer(bytearray(c.data), dtype=torch.uint8) if n else torch.zeros(16, dtype=torch.uint8)).cuda()
        bufs.append(raw)
        descs.append(engine.make_cloud_desc(c.fields, c.point_step, n, raw, transform=T))
    fcfg = engine.make_filter_cfg(**(filter_kw or {}))
    if stages:
        out, counts, _ = ctx.pipeline_run(descs, engine.make_pipeline_cfg(fcfg, **stages))
        ctx.check()
        c = counts.cpu().numpy()
        n_out = int(c[_capi.CNT_OUTPUT])
        return out[:n_out], n_out, c
    xyzi, src, _, cnt = ctx.frontend(descs, fcfg, want_src=want_src)
    ctx.check()
    n_out = int(cnt.item())
    return xyzi[:n_out], n_out, (src[:n_out] if src is not None else None)


class SensorSynchronizer:
    """Host-side pairing of the latest message per sensor.

    ``mode='sync'``   - emit only when every sensor has a message and all stamps lie within
                        ``slop`` seconds (ApproximateTime-like);
    ``mode='robust'`` - emit as soon as every *live* sensor has reported, dropping sensors
                        whose latest message is older than ``timeout`` seconds, so one dead
                        LiDAR does not stall the output.
    """

    def __init__(self, n_sensors: int, mode: str = "sync", slop: float = 0.05, timeout: float = 0.2):
        if mode not in ("sync", "robust"):
            raise ValueError("mode must be 'sync' or 'robust'")
        self.n, self.mode, self.slop, self.timeout = n_sensors, mode, slop, timeout
        self.latest = [None] * n_sensors
        self.last_seen = [None] * n_sensors

    @staticmethod
    def _stamp(msg) -> float:
        st = msg.header.stamp
        return float(getattr(st, "sec", 0)) + 1e-9 * float(getattr(st, "nanosec", 0))

    def add(self, sensor: int, msg):
        """Store ``msg``; returns ``(sensor_ids, msgs)`` when a set is ready, else ``None``."""
        self.latest[sensor] = msg
        self.last_seen[sensor] = self._stamp(msg)
        have = [(i, m) for i, m in enumerate(self.latest) if m is not None]
        now = max(self._stamp(m) for _, m in have)
        if self.mode == "sync":
            if len(have) < self.n:
                return None
            stamps = np.array([self._stamp(m) for _, m in have])
            if stamps.max() - stamps.min() > self.slop:
                return None
        else:
            # sensors heard from within `timeout` are live; wait until each of them has a
            # buffered message of the current sweep (within `slop` of the newest stamp)
            live = [i for i in range(self.n) if self.last_seen[i] is not None and now - self.last_seen[i] <= self.timeout]
            have = [(i, m) for i, m in have if i in live and now - self._stamp(m) <= self.slop]
            if len(have) < len(live):
                return None
        ids, msgs = [i for i, _ in have], [m for _, m in have]
        for i in ids:
            self.latest[i] = None
        return ids, msgs


# ---- the node ------------------------------------------------------------------------------------
def _quat_to_matrix(t, q) -> np.ndarray:
    """4x4 float64 from a translation (x, y, z) and a unit quaternion (x, y, z, w)."""
    x, y, z, w = (float(v) for v in q)
    T = np.eye(4)
    T[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    T[:3, 3] = [float(v) for v in t]
    return T


def _make_node_class():
    from ._ros_compat import (HAVE_ROS, Buffer, ConnectivityException, Duration, ExtrapolationException, Header,
                              LookupException, Node, PointCloud2, PointField, QoSHistoryPolicy, QoSProfile,
                              QoSReliabilityPolicy, Time, TransformListener, point_cloud2)
    from .utils import numpy_struct_to_pointcloud2

    class PointcloudConcatenatorNode(Node):
        """The node ``pointcloud_concatenator.py:1-5`` describes: n ``PointCloud2`` topics in, one merged
        cloud out, in ``target_frame`` when one is given (per-sensor TF looked up once and cached, as the
        preprocessor does for its static transform, ``pp.py:704-732``).

        ``sync_mode='sync'``: with ROS 2 present the sets come from
        ``message_filters.ApproximateTimeSynchronizer`` (the import the reference already carries,
        ``pp.py:102``); without it from :class:`SensorSynchronizer` with the same ``slop``.
        ``sync_mode='robust'``: :class:`SensorSynchronizer` in robust mode - a sensor that has been silent
        for ``timeout`` seconds no longer holds the output back.

        The merge + per-sensor transform (+ optional voxel grid) is ONE launch chain on the GPU
        (:func:`concatenate`); the merged records (x, y, z, intensity as float32) are packed on the device
        and copied to the host once.
        """

        def __init__(self, node_name='pointcloud_concatenator', **kw):
            super().__init__(node_name, **kw)
            self.declare_parameter('input_topics', ['/lidar_front/points', '/lidar_rear/points'])
            self.declare_parameter('output_topic', '/lidar/points_concatenated')
            self.declare_parameter('target_frame', '')
            self.declare_parameter('sync_mode', 'sync')
            self.declare_parameter('slop', 0.05)
            self.declare_parameter('timeout', 0.2)
            self.declare_parameter('queue_size', 10)
            self.declare_parameter('transform_timeout', 0.1)
            self.declare_parameter('voxel_size', 0.0)
            self.declare_parameter('remove_nans', True)
            self.declare_parameter('qos', 'SENSOR_DATA')
            gp = lambda n: self.get_parameter(n).value  # noqa: E731
            self.input_topics = list(gp('input_topics'))
            if not 1 <= len(self.input_topics) <= _capi.APC_MAX_CLOUDS:
                raise ValueError(f"input_topics: 1..{_capi.APC_MAX_CLOUDS} sensors")
            self.target_frame = str(gp('target_frame'))
            self.sync_mode = str(gp('sync_mode')).lower()
            self.slop, self.timeout = float(gp('slop')), float(gp('timeout'))
            self.transform_timeout = float(gp('transform_timeout'))
            self.voxel_size = float(gp('voxel_size'))
            self.remove_nans = bool(gp('remove_nans'))
            qos_name = str(gp('qos')).lower()
            qos = QoSProfile(reliability=QoSReliabilityPolicy.BEST_EFFORT if qos_name == 'sensor_data'
                             else QoSReliabilityPolicy.RELIABLE, history=QoSHistoryPolicy.KEEP_LAST, depth=1)
            self.tf_buffer = Buffer()
            self.tf_listener = TransformListener(self.tf_buffer, self)
            self.sensor_tf = [None] * len(self.input_topics)       # cached 4x4 per sensor (static extrinsics)
            self.frame_count = 0
            self.pointfields, self.point_step = numpy_struct_to_pointcloud2(
                ['x', 'y', 'z', 'intensity'], [PointField.FLOAT32] * 4)
            self._pinned = [None, None]
            self.pointcloud_pub = self.create_publisher(PointCloud2, str(gp('output_topic')), qos)
            self.synchronizer = None
            self.filter_sync = None
            if self.sync_mode == 'sync' and HAVE_ROS:           # the message_filters seam
                from message_filters import ApproximateTimeSynchronizer, Subscriber
                subs = [Subscriber(self, PointCloud2, t, qos_profile=qos) for t in self.input_topics]
                self.filter_sync = ApproximateTimeSynchronizer(subs, int(gp('queue_size')), self.slop)
                self.filter_sync.registerCallback(lambda *msgs: self.publish_set(list(range(len(msgs))), list(msgs)))
                self.subs = subs
            else:
                if self.sync_mode not in ('sync', 'robust'):
                    raise ValueError("sync_mode must be 'sync' or 'robust'")
                self.synchronizer = SensorSynchronizer(len(self.input_topics), mode=self.sync_mode, slop=self.slop,
                                                       timeout=self.timeout)
                self.subs = [self.create_subscription(PointCloud2, t, (lambda m, i=i: self.sensor_callback(i, m)), qos)
                             for i, t in enumerate(self.input_topics)]
            self._tf_errors = (LookupException, ConnectivityException, ExtrapolationException)
            self._Time, self._Duration, self._Header, self._pc2 = Time, Duration, Header, point_cloud2

        # one sensor's message arrived (SensorSynchronizer modes)
        def sensor_callback(self, sensor: int, msg):
            ready = self.synchronizer.add(sensor, msg)
            if ready is not None:
                self.publish_set(*ready)

        def lookup_sensor_tf(self, sensor: int, frame_id: str, stamp=None):
            """4x4 sensor frame -> target frame, cached after the first successful lookup."""
            if not self.target_frame or frame_id == self.target_frame:
                return None
            if self.sensor_tf[sensor] is not None:
                return self.sensor_tf[sensor]
            tf = self.tf_buffer.lookup_transform(self.target_frame, frame_id, self._Time.from_msg(stamp),
                                                 self._Duration(seconds=self.transform_timeout))
            tr, ro = tf.transform.translation, tf.transform.rotation
            self.sensor_tf[sensor] = _quat_to_matrix((tr.x, tr.y, tr.z), (ro.x, ro.y, ro.z, ro.w))
            return self.sensor_tf[sensor]

        def publish_set(self, sensor_ids, msgs):
            """Merge one synchronised set and publish it.  A sensor whose transform cannot be resolved
            is left out of this set (robust behaviour) rather than published in the wrong frame."""
            if self.pointcloud_pub.get_subscription_count() == 0:
                return None
            clouds, transforms = [], []
            for i, m in zip(sensor_ids, msgs):
                if m.width * m.height == 0:
                    continue
                try:
                    T = self.lookup_sensor_tf(i, m.header.frame_id, m.header.stamp)
                except self._tf_errors as e:
                    self.get_logger().warn(f"no transform {m.header.frame_id} -> {self.target_frame}: {e}; sensor {i} skipped",
                                           throttle_duration_sec=2.0)
                    continue
                clouds.append(m)
                transforms.append(T)
            if not clouds:
                return None
            stages = dict(voxel_size=self.voxel_size) if self.voxel_size > 0.0 else None
            filter_kw = dict(skip_nans=self.remove_nans, remove_nan=self.remove_nans, remove_inf=self.remove_nans)
            xyzi, n, _ = concatenate(clouds, transforms, filter_kw=filter_kw, stages=stages)
            ctx = geometry.get_context(max(n, 1))
            fields = [(f.offset, f.datatype, src, None) for f, src in zip(self.pointfields, (1, 2, 3, 4))]
            raw = ctx.repack(xyzi.contiguous(), fields, self.point_step) if n else torch.zeros(0, dtype=torch.uint8)
            nbytes = n * self.point_step
            slot = self.frame_count & 1
            if self._pinned[slot] is None or self._pinned[slot].numel() < nbytes:
                self._pinned[slot] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8).pin_memory()
            if n:
                self._pinned[slot][:nbytes].copy_(raw[:nbytes], non_blocking=True)
                torch.cuda.current_stream().synchronize()
            dt = np.dtype({'names': ['x', 'y', 'z', 'intensity'], 'formats': ['<f4'] * 4, 'offsets': [0, 4, 8, 12],
                           'itemsize': self.point_step})
            records = np.frombuffer(self._pinned[slot].numpy(), dtype=dt, count=n)
            newest = max(msgs, key=SensorSynchronizer._stamp)
            header = self._Header()
            header.stamp = newest.header.stamp
            header.frame_id = self.target_frame or clouds[0].header.frame_id
            out = self._pc2.create_cloud(header, self.pointfields, records)
            out.is_dense = bool(self.remove_nans)
            self.pointcloud_pub.publish(out)
            self.frame_count += 1
            return out

    return PointcloudConcatenatorNode


def __getattr__(name):                      # the node class needs the ROS seam: built on first use
    if name == "PointcloudConcatenatorNode":
        cls = _make_node_class()
        globals()[name] = cls
        return cls
    raise AttributeError(name)


def main(args=None):                        # pragma: no cover - needs a ROS 2 installation
    from ._ros_compat import rclpy
    rclpy.init(args=args)
    node = __getattr__("PointcloudConcatenatorNode")()
    try:
        rclpy.spin(node)
    finally:
        node.destroy_node()
        rclpy.shutdown()

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

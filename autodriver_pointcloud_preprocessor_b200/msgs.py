"""Minimal stand-ins for the ROS 2 message types the hot path touches.

When ROS 2 is installed the real ``sensor_msgs.msg.PointCloud2`` / ``PointField`` /
``std_msgs.msg.Header`` are used unchanged (they are duck-compatible); when it is absent
(this image) these plain classes carry the same attributes so the unpack / repack path
and its tests run without ROS.

Reference: the attributes read by ``utils.py:202-223`` (``pointcloud_to_dict``),
``pointcloud_preprocessor.py:546-574`` (``set_fields``) and ``:762-769``
(``tensor_to_ros_cloud`` -> ``create_cloud``).
"""
from __future__ import annotations

import sys

try:  # pragma: no cover - ROS is not present in the build image
    from sensor_msgs.msg import PointCloud2, PointField  # type: ignore
    from std_msgs.msg import Header  # type: ignore
    ROS_MESSAGES = True
except ImportError:
    ROS_MESSAGES = False

    class Time:
        __slots__ = ("sec", "nanosec")

        def __init__(self, sec: int = 0, nanosec: int = 0):
            self.sec = int(sec)
            self.nanosec = int(nanosec)

        def __repr__(self):
            return f"Time(sec={self.sec}, nanosec={self.nanosec})"

    class Header:
        __slots__ = ("stamp", "frame_id")

        def __init__(self, stamp=None, frame_id: str = ""):
            self.stamp = stamp if stamp is not None else Time()
            self.frame_id = frame_id

        def __repr__(self):
            return f"Header(stamp={self.stamp}, frame_id={self.frame_id!r})"

    class PointField:
        """sensor_msgs/PointField datatype constants and slots."""
        INT8 = 1
        UINT8 = 2
        INT16 = 3
        UINT16 = 4
        INT32 = 5
        UINT32 = 6
        FLOAT32 = 7
        FLOAT64 = 8
        __slots__ = ("name", "offset", "datatype", "count")

        def __init__(self, name: str = "", offset: int = 0, datatype: int = 0, count: int = 1):
            self.name = name
            self.offset = int(offset)
            self.datatype = int(datatype)
            self.count = int(count)

        def __repr__(self):
            return (f"PointField(name={self.name!r}, offset={self.offset}, "
                    f"datatype={self.datatype}, count={self.count})")

    class PointCloud2:
        __slots__ = ("header", "height", "width", "fields", "is_bigendian",
                     "point_step", "row_step", "data", "is_dense")

        def __init__(self, header=None, height: int = 1, width: int = 0, fields=None,
                     is_bigendian: bool = (sys.byteorder != "little"), point_step: int = 0,
                     row_step: int = 0, data=b"", is_dense: bool = False):
            self.header = header if header is not None else Header()
            self.height = int(height)
            self.width = int(width)
            self.fields = list(fields) if fields is not None else []
            self.is_bigendian = bool(is_bigendian)
            self.point_step = int(point_step)
            self.row_step = int(row_step)
            self.data = data
            self.is_dense = bool(is_dense)

        def __repr__(self):
            return (f"PointCloud2(height={self.height}, width={self.width}, "
                    f"point_step={self.point_step}, is_dense={self.is_dense}, "
                    f"fields={[f.name for f in self.fields]})")


#: byte size of each PointField datatype (index = datatype constant)
DATATYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 2, 5: 4, 6: 4, 7: 4, 8: 8}

"""ROS 2 import seam.

With ROS 2 installed the node derives from the real ``rclpy.node.Node`` and uses the real
QoS / parameter / TF classes.  Without it (this image has no rclpy, sensor_msgs, tf2_ros) the
small stand-ins below provide the same call surface the node touches, so the node class, its
parameter table and its callback can be instantiated and driven by tests and by the replay
tools with ``msgs.PointCloud2`` objects.
"""
from __future__ import annotations

import logging
import sys
import time

from .msgs import Header, PointCloud2, PointField, ROS_MESSAGES  # noqa: F401

try:  # pragma: no cover - not available in the build image
    import rclpy
    from rclpy.node import Node
    from rclpy.parameter import Parameter
    from rclpy.qos import QoSProfile, QoSReliabilityPolicy, QoSHistoryPolicy
    from rcl_interfaces.msg import ParameterDescriptor, ParameterType, SetParametersResult
    import tf2_ros
    from tf2_ros import Buffer, TransformListener, LookupException, ConnectivityException, ExtrapolationException
    from sensor_msgs_py import point_cloud2
    from rclpy.duration import Duration
    from rclpy.time import Time
    HAVE_ROS = True
except ImportError:
    HAVE_ROS = False
    rclpy = None

    class _Type:
        NOT_SET, BOOL, INTEGER, DOUBLE, STRING, BYTE_ARRAY, BOOL_ARRAY, INTEGER_ARRAY, DOUBLE_ARRAY, STRING_ARRAY = range(10)
        INT = INTEGER

        @staticmethod
        def from_value(v):
            if isinstance(v, bool):
                return _Type.BOOL
            if isinstance(v, int):
                return _Type.INTEGER
            if isinstance(v, float):
                return _Type.DOUBLE
            if isinstance(v, str):
                return _Type.STRING
            if isinstance(v, (list, tuple)):
                if len(v) and all(isinstance(x, bool) for x in v):
                    return _Type.BOOL_ARRAY
                if len(v) and all(isinstance(x, int) for x in v):
                    return _Type.INTEGER_ARRAY
                if len(v) and all(isinstance(x, str) for x in v):
                    return _Type.STRING_ARRAY
                return _Type.DOUBLE_ARRAY
            return _Type.NOT_SET

    class ParameterType:
        PARAMETER_NOT_SET, PARAMETER_BOOL, PARAMETER_INTEGER, PARAMETER_DOUBLE, PARAMETER_STRING = range(5)
        PARAMETER_BYTE_ARRAY, PARAMETER_BOOL_ARRAY, PARAMETER_INTEGER_ARRAY, PARAMETER_DOUBLE_ARRAY, \
            PARAMETER_STRING_ARRAY = range(5, 10)

    class ParameterDescriptor:
        def __init__(self, description="", type=0, **_kw):
            self.description, self.type = description, type

    class _ParameterValue:
        def __init__(self, v):
            self._v = v

        bool_value = property(lambda s: bool(s._v))
        integer_value = property(lambda s: int(s._v))
        double_value = property(lambda s: float(s._v))
        string_value = property(lambda s: str(s._v))
        double_array_value = property(lambda s: [float(x) for x in s._v])

    class Parameter:
        Type = _Type

        def __init__(self, name, type_=None, value=None):
            self.name, self.value = name, value
            self.type_ = _Type.from_value(value) if type_ is None else type_

        def get_parameter_value(self):
            return _ParameterValue(self.value)

    class SetParametersResult:
        def __init__(self, successful=True, reason=""):
            self.successful, self.reason = successful, reason

    class QoSReliabilityPolicy:
        RELIABLE, BEST_EFFORT = 1, 2

    class QoSHistoryPolicy:
        KEEP_LAST, KEEP_ALL = 1, 2

    class QoSProfile:
        def __init__(self, reliability=QoSReliabilityPolicy.RELIABLE, history=QoSHistoryPolicy.KEEP_LAST, depth=1):
            self.reliability, self.history, self.depth = reliability, history, depth

    class Duration:
        """rclpy.duration.Duration stand-in (pp.py:718 wraps transform_timeout in one)."""

        def __init__(self, *, seconds=0.0, nanoseconds=0):
            self.nanoseconds = int(round(float(seconds) * 1e9)) + int(nanoseconds)

    class Time:
        """rclpy.time.Time stand-in (pp.py:477 converts the header stamp with Time.from_msg)."""

        def __init__(self, *, seconds=0, nanoseconds=0):
            self.nanoseconds = int(seconds) * 1_000_000_000 + int(nanoseconds)

        @classmethod
        def from_msg(cls, msg):
            if msg is None:
                return cls()
            return cls(seconds=getattr(msg, "sec", 0), nanoseconds=getattr(msg, "nanosec", 0))

    class LookupException(Exception):
        pass

    class ConnectivityException(Exception):
        pass

    class ExtrapolationException(Exception):
        pass

    class _Vec:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class TransformStamped:
        def __init__(self, translation=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0, 1.0)):
            self.transform = _Vec(translation=_Vec(x=translation[0], y=translation[1], z=translation[2]),
                                  rotation=_Vec(x=rotation[0], y=rotation[1], z=rotation[2], w=rotation[3]))

    class Buffer:
        """tf2 buffer stand-in: transforms are registered with :meth:`set_transform`."""

        def __init__(self):
            self._tf = {}

        def set_transform(self, target_frame, source_frame, translation, rotation_xyzw):
            self._tf[(target_frame, source_frame)] = TransformStamped(translation, rotation_xyzw)

        def lookup_transform(self, target_frame, source_frame, time=None, timeout=None):
            # tf2_ros.Buffer computes `start_time + timeout`: anything but a Duration / Time pair raises
            if timeout is not None and not isinstance(timeout, Duration):
                raise TypeError("lookup_transform: timeout must be a Duration (tf2_ros.Buffer signature)")
            if time is not None and not isinstance(time, Time):
                raise TypeError("lookup_transform: time must be a Time (tf2_ros.Buffer signature)")
            try:
                return self._tf[(target_frame, source_frame)]
            except KeyError:
                raise LookupException(f'"{target_frame}" passed to lookupTransform argument target_frame does not exist.')

    class TransformListener:
        def __init__(self, buffer, node):
            self.buffer = buffer

    class _tf2_ros:
        LookupException, ConnectivityException, ExtrapolationException = (LookupException, ConnectivityException,
                                                                          ExtrapolationException)
        Buffer, TransformListener = Buffer, TransformListener

        class TransformBroadcaster:
            def __init__(self, node):
                pass

    tf2_ros = _tf2_ros

    class _Publisher:
        def __init__(self, topic, depth):
            self.topic, self.depth, self.messages, self.subscribers = topic, depth, [], 1

        def get_subscription_count(self):
            return self.subscribers

        def publish(self, msg):
            self.messages.append(msg)
            del self.messages[:-max(1, self.depth)]

    class _Clock:
        class _Now:
            def to_msg(self):
                from .msgs import Time
                t = time.time()
                return Time(int(t), int((t % 1) * 1e9))

        def now(self):
            return self._Now()

    class _Logger:
        def __init__(self, name):
            self._log = logging.getLogger(name)
            self._last = {}

        def _emit(self, level, msg, throttle_duration_sec=None, **_kw):
            if throttle_duration_sec is not None:
                now = time.monotonic()
                if now - self._last.get(msg, -1e9) < throttle_duration_sec:
                    return
                self._last[msg] = now
            self._log.log(level, msg)

        def info(self, msg, **kw):
            self._emit(logging.INFO, msg, **kw)

        def warn(self, msg, **kw):
            self._emit(logging.WARNING, msg, **kw)

        warning = warn

        def error(self, msg, **kw):
            self._emit(logging.ERROR, msg, **kw)

        def debug(self, msg, **kw):
            self._emit(logging.DEBUG, msg, **kw)

    class Node:
        """The slice of ``rclpy.node.Node`` the preprocessor uses."""

        def __init__(self, node_name, parameter_overrides=None, **_kw):
            self._node_name = node_name
            self._params = {}
            self._overrides = dict(parameter_overrides or {})
            self._param_callbacks = []
            self._logger = _Logger(node_name)
            self._clock = _Clock()
            self.subscriptions_, self.publishers_ = [], []
            self.declare_parameter("use_sim_time", False)

        def declare_parameter(self, name, value=None, descriptor=None):
            v = self._overrides.get(name, value)
            self._params[name] = Parameter(name, value=v)
            return self._params[name]

        def has_parameter(self, name):
            return name in self._params

        def get_parameter(self, name):
            return self._params[name]

        def set_parameters(self, params):
            results = []
            for p in params:
                res = SetParametersResult(True)
                for cb in self._param_callbacks:
                    res = cb([p])
                    if not res.successful:
                        break
                if res.successful:
                    self._params[p.name] = p
                results.append(res)
            return results

        def add_on_set_parameters_callback(self, cb):
            self._param_callbacks.append(cb)

        def create_subscription(self, msg_type, topic, callback, qos_profile=None):
            sub = (topic, callback, qos_profile)
            self.subscriptions_.append(sub)
            return sub

        def create_publisher(self, msg_type, topic, qos):
            pub = _Publisher(topic, qos if isinstance(qos, int) else getattr(qos, "depth", 1))
            self.publishers_.append(pub)
            return pub

        def get_logger(self):
            return self._logger

        def get_clock(self):
            return self._clock

        def get_fully_qualified_name(self):
            return f"/{self._node_name}"

        def destroy_node(self):
            pass

    class _PointCloud2Module:
        """``sensor_msgs_py.point_cloud2`` stand-in: only ``create_cloud`` (pp.py:769)."""

        @staticmethod
        def create_cloud(header, fields, points):
            import numpy as np
            pts = np.ascontiguousarray(points)
            return PointCloud2(header=header, height=1, width=int(pts.shape[0]), fields=list(fields),
                               is_bigendian=sys.byteorder != "little", point_step=int(pts.dtype.itemsize),
                               row_step=int(pts.dtype.itemsize * pts.shape[0]), data=pts.tobytes(), is_dense=False)

    point_cloud2 = _PointCloud2Module

"""File replay for the preprocessing path: PCD files -> ``PointCloud2`` messages.

The reference only states the intent (``pointcloud_loader.py:1-5``: "load pointclouds from a
directory of .pcds ... add support for looping"; ``pcap_player.py`` is empty).  This module is the
host-side shell of the batched replay configuration: it turns files into the same
``sensor_msgs/PointCloud2`` byte buffers a driver would publish, so that they can be fed to
``PointcloudPreprocessorNode.callback`` or to ``replay.ScanPipeline.process_host`` unchanged.
Pure host code (file parsing); nothing here touches the GPU.

Supported: PCD v0.7 ``DATA ascii`` and ``DATA binary`` with scalar fields (COUNT 1) of the
PointField types (I/U/F, sizes 1/2/4/8).  ``binary_compressed`` is not supported.
"""
from __future__ import annotations

import os

import numpy as np

from .msgs import Header, PointCloud2, PointField

_PCD_TO_DATATYPE = {("I", 1): PointField.INT8, ("U", 1): PointField.UINT8, ("I", 2): PointField.INT16,
                    ("U", 2): PointField.UINT16, ("I", 4): PointField.INT32, ("U", 4): PointField.UINT32,
                    ("F", 4): PointField.FLOAT32, ("F", 8): PointField.FLOAT64}
_NP_OF = {PointField.INT8: "i1", PointField.UINT8: "u1", PointField.INT16: "<i2", PointField.UINT16: "<u2",
          PointField.INT32: "<i4", PointField.UINT32: "<u4", PointField.FLOAT32: "<f4", PointField.FLOAT64: "<f8"}


def read_pcd(path: str, frame_id: str = "lidar") -> PointCloud2:
    """One PCD file as an unorganised (height 1), little-endian ``PointCloud2`` whose records are the
    file's fields packed in file order."""
    with open(path, "rb") as f:
        raw = f.read()
    header, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, rest = line.partition(" ")
        header[key.upper()] = rest.split()
        if key.upper() == "DATA":
            break
    names = header["FIELDS"]
    sizes = [int(v) for v in header["SIZE"]]
    types = header["TYPE"]
    counts = [int(v) for v in header.get("COUNT", ["1"] * len(names))]
    if any(c != 1 for c in counts):
        raise NotImplementedError("PCD fields with COUNT != 1 are not supported")
    n = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header["HEIGHT"][0])
    try:
        datatypes = [_PCD_TO_DATATYPE[(t.upper(), s)] for t, s in zip(types, sizes)]
    except KeyError as e:
        raise NotImplementedError(f"unsupported PCD field type/size {e.args[0]}") from None
    dtype = np.dtype([(nm, _NP_OF[dt]) for nm, dt in zip(names, datatypes)])
    kind = header["DATA"][0].lower()
    if kind == "binary":
        arr = np.frombuffer(raw, dtype=dtype, count=n, offset=pos)
    elif kind == "ascii":
        rows = np.loadtxt(raw[pos:].decode("ascii").splitlines(), dtype=np.float64, ndmin=2)
        if rows.shape[0] != n:
            raise ValueError(f"{path}: {rows.shape[0]} rows, header says {n}")
        arr = np.zeros(n, dtype=dtype)
        for c, nm in enumerate(names):
            arr[nm] = rows[:, c].astype(dtype[nm])
    else:
        raise NotImplementedError(f"PCD DATA {kind} is not supported")
    fields, offset = [], 0
    for nm, dt in zip(names, datatypes):
        fields.append(PointField(name=nm, offset=offset, datatype=dt, count=1))
        offset += np.dtype(_NP_OF[dt]).itemsize
    msg = PointCloud2()
    msg.header = Header(frame_id=frame_id)
    msg.height, msg.width = 1, n
    msg.fields = fields
    msg.is_bigendian = False
    msg.point_step = offset
    msg.row_step = offset * n
    msg.data = arr.tobytes()
    finite = [nm for nm, dt in zip(names, datatypes) if dt in (PointField.FLOAT32, PointField.FLOAT64)]
    msg.is_dense = bool(all(np.isfinite(arr[nm]).all() for nm in finite))
    return msg


def write_pcd(path: str, msg: PointCloud2, binary: bool = True) -> None:
    """Inverse of :func:`read_pcd` for packed little-endian clouds (used by the tests and to dump the
    node's output without Open3D, cf. ``pointcloud_saver`` pp.py:1006-1030)."""
    inv = {v: k for k, v in _PCD_TO_DATATYPE.items()}
    dtype = np.dtype({"names": [f.name for f in msg.fields], "formats": [_NP_OF[f.datatype] for f in msg.fields],
                      "offsets": [f.offset for f in msg.fields], "itemsize": msg.point_step})
    arr = np.frombuffer(bytes(msg.data), dtype=dtype, count=msg.width * msg.height)
    packed = np.zeros(arr.shape[0], dtype=np.dtype([(f.name, _NP_OF[f.datatype]) for f in msg.fields]))
    for f in msg.fields:
        packed[f.name] = arr[f.name]
    n = arr.shape[0]
    head = ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7",
            "FIELDS " + " ".join(f.name for f in msg.fields),
            "SIZE " + " ".join(str(inv[f.datatype][1]) for f in msg.fields),
            "TYPE " + " ".join(inv[f.datatype][0] for f in msg.fields),
            "COUNT " + " ".join("1" for _ in msg.fields),
            f"WIDTH {n}", "HEIGHT 1", "VIEWPOINT 0 0 0 1 0 0 0", f"POINTS {n}",
            "DATA " + ("binary" if binary else "ascii")]
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode("ascii"))
        if binary:
            f.write(packed.tobytes())
        else:
            for row in packed:
                f.write((" ".join(repr(v.item()) for v in row) + "\n").encode("ascii"))


class DirectoryReplay:
    """Iterates over the ``.pcd`` files of a directory in name order, optionally looping forever
    ("add support for looping", ``pointcloud_loader.py:4``)."""

    def __init__(self, directory: str, loop: bool = False, frame_id: str = "lidar"):
        self.paths = sorted(os.path.join(directory, f) for f in os.listdir(directory) if f.lower().endswith(".pcd"))
        if not self.paths:
            raise FileNotFoundError(f"no .pcd files in {directory}")
        self.loop, self.frame_id = loop, frame_id

    def __len__(self):
        return len(self.paths)

    def __iter__(self):
        while True:
            for p in self.paths:
                yield read_pcd(p, self.frame_id)
            if not self.loop:
                return

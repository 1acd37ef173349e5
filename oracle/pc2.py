"""PointCloud2 <-> numpy: CPU oracle (test infrastructure, see oracle/__init__.py).

Restates ``sensor_msgs_py.point_cloud2.read_points`` / ``create_cloud`` (dependency
``ros2/common_interfaces`` - not vendored in the reference, distro unpinned; the Humble+
numpy-based API is what ``utils.py:206-211`` and ``pp.py:769`` call) and the reference's
own conversion helpers.  PARITY UNPINNED for the sensor_msgs_py part; the helpers that
live in the reference (``convert_pointcloud_to_numpy`` etc.) are pinned by
``tests/golden/``.
"""
from __future__ import annotations

import sys

import numpy as np

# PointField datatype -> numpy dtype (utils.py:28-37)
FIELD_DTYPE_MAP = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32,
                   6: np.uint32, 7: np.float32, 8: np.float64}

# utils.py:41-48
VENDOR_MAPPINGS = {
    "intensity": ["I", "intensity"],
    "ring": ["C", "ring", "line"],
    "time": ["t", "time", "timestamp"],
    "return_type": ["return_type", "tag", "R"],
    "azimuth": ["azimuth"],
    "distance": ["distance", "depth", "d"],
}


def dtype_from_fields(fields, point_step=None) -> np.dtype:
    """sensor_msgs_py ``dtype_from_fields``: structured dtype with explicit offsets."""
    names, formats, offsets = [], [], []
    for f in fields:
        base = np.dtype(FIELD_DTYPE_MAP[f.datatype])
        if f.count == 1:
            names.append(f.name)
            formats.append(base)
            offsets.append(f.offset)
        else:
            for a in range(f.count):
                names.append(f"{f.name}_{a}")
                formats.append(base)
                offsets.append(f.offset + a * base.itemsize)
    spec = {"names": names, "formats": formats, "offsets": offsets}
    if point_step is not None:
        spec["itemsize"] = point_step
    return np.dtype(spec)


def read_points(cloud, field_names=None, skip_nans=False, reshape_organized_cloud=False):
    """``read_points`` as called at utils.py:206-211.

    NaN rows are dropped only when ``skip_nans and not cloud.is_dense``; the test looks at
    *every selected field* (NaN only - inf is kept).
    """
    points = np.ndarray(shape=(cloud.width * cloud.height,),
                        dtype=dtype_from_fields(cloud.fields, point_step=cloud.point_step),
                        buffer=cloud.data)
    if field_names is not None:
        assert all(n in points.dtype.names for n in field_names)
        points = points[list(field_names)]
    if bool(sys.byteorder != "little") != bool(cloud.is_bigendian):
        points = points.byteswap()
    if skip_nans and not cloud.is_dense:
        keep = np.ones(len(points), dtype=bool)
        for name in points.dtype.names:
            keep &= ~np.isnan(points[name])
        points = points[keep]
    if reshape_organized_cloud and cloud.height > 1:
        points = points.reshape(cloud.width, cloud.height)
    return points


def read_points_mask(cloud, field_names=None, skip_nans=False):
    """The keep-mask ``read_points`` applies (all True when nothing is skipped)."""
    points = np.ndarray(shape=(cloud.width * cloud.height,),
                        dtype=dtype_from_fields(cloud.fields, point_step=cloud.point_step),
                        buffer=cloud.data)
    if field_names is not None:
        points = points[list(field_names)]
    keep = np.ones(len(points), dtype=bool)
    if skip_nans and not cloud.is_dense:
        for name in points.dtype.names:
            keep &= ~np.isnan(points[name])
    return keep


def parse_differing_fields(options, field_names):
    """utils.py:423-438 - last matching alias wins; the alias spelling is returned."""
    if isinstance(options, str):
        options = [options]
    found, name = [], None
    for option in options:
        if option.lower() in field_names:
            found.append(option)
            name = option
    return any(found), name


def get_pointcloud_metadata(field_names, vendor_mappings=None):
    """utils.py:441-472."""
    if vendor_mappings is None:
        vendor_mappings = VENDOR_MAPPINGS
    field_names = [f.lower() for f in field_names]
    if {"r", "g", "b"}.issubset(field_names):
        has_rgb = True
    else:
        has_rgb, _ = parse_differing_fields("rgb", field_names)
    has_i, n_i = parse_differing_fields(vendor_mappings["intensity"], field_names)
    has_r, n_r = parse_differing_fields(vendor_mappings["ring"], field_names)
    has_t, n_t = parse_differing_fields(vendor_mappings["time"], field_names)
    has_rt, n_rt = parse_differing_fields(vendor_mappings["return_type"], field_names)
    return {"has_rgb": has_rgb, "has_intensity": has_i, "intensity_field_name": n_i,
            "has_ring": has_r, "ring_field_name": n_r, "has_time": has_t, "time_field_name": n_t,
            "has_return_type": has_rt, "return_type_field_name": n_rt}


def extract_rgb_from_pointcloud(rgb):
    """utils.py:324-345: packed float32 rgb -> (N,3) uint8."""
    b = rgb.view(np.uint32)
    return np.vstack((((b >> 16) & 0xFF).astype(np.uint8), ((b >> 8) & 0xFF).astype(np.uint8),
                      (b & 0xFF).astype(np.uint8))).T.astype(np.uint8)


def convert_pointcloud_to_numpy(arr, meta):
    """utils.py:51-133: structured array -> SoA dict with the reference's casts."""
    out = {"positions": np.vstack((arr["x"], arr["y"], arr["z"])).T.astype(np.float32)}
    field_names = meta.get("field_names", ["x", "y", "z"])
    if meta.get("has_rgb", False):
        if {"r", "g", "b"}.issubset(field_names):
            out["rgb"] = np.vstack((arr["r"].astype(np.uint8), arr["g"].astype(np.uint8),
                                    arr["b"].astype(np.uint8))).T
        else:
            out["rgb"] = extract_rgb_from_pointcloud(arr["rgb"].astype(np.float32))
    if meta.get("has_intensity", False):
        out["intensity"] = arr[meta["intensity_field_name"]].astype(np.float32)
    if meta.get("has_ring", False):
        out["ring"] = arr[meta["ring_field_name"]].astype(np.uint16)
    if meta.get("has_time", False):
        out["time"] = arr[meta["time_field_name"]].astype(np.float64)
    if meta.get("has_return_type", False):
        out["return_type"] = arr[meta["return_type_field_name"]].astype(np.uint8)
    return out


def pointcloud_to_dict(cloud, field_names=None, skip_nans=True, organize_cloud=False, metadata_dict=None):
    """utils.py:202-223."""
    if not metadata_dict:
        metadata_dict = {}
    metadata_dict.update({"header": cloud.header, "field_names": None})
    arr = read_points(cloud, field_names=field_names, skip_nans=skip_nans,
                      reshape_organized_cloud=organize_cloud)
    metadata_dict["field_names"] = arr.dtype.names
    metadata_dict["num_fields"] = len(arr.dtype.names)
    if not metadata_dict.get("has_intensity", False):
        metadata_dict.update(get_pointcloud_metadata(metadata_dict["field_names"]))
    return convert_pointcloud_to_numpy(arr, metadata_dict), metadata_dict


def packed_fields(field_names, field_datatypes):
    """utils.py:140-199 ``numpy_struct_to_pointcloud2``: cumulative offsets, no padding.

    Returns ``([(name, offset, datatype)], point_step)``.
    """
    out, offset = [], 0
    for name, dt in zip(field_names, field_datatypes):
        out.append((name, offset, dt))
        offset += np.dtype(FIELD_DTYPE_MAP[dt]).itemsize
    return out, offset


def repack(cloud_fields, positions, attrs: dict, meta: dict, normals=None, estimate_normals=False):
    """pp.py:546-625 ``set_fields`` + ``prepare_pointcloud``: output bytes of the published cloud.

    Same field names/datatypes as the input, re-packed with no padding; x,y,z filled from
    ``positions``; intensity / ring / time / return_type cast back to their input dtype;
    every other input field is emitted as zeros.  ``attrs`` maps the canonical attribute
    names ('intensity', 'ring', 'time', 'return_type') to arrays of length P.
    """
    names = [f.name for f in cloud_fields]
    dts = [f.datatype for f in cloud_fields]
    if estimate_normals:
        names += ["normal_x", "normal_y", "normal_z"]
        dts += [7, 7, 7]
    new_dtype = np.dtype([(n, FIELD_DTYPE_MAP[d]) for n, d in zip(names, dts)])
    out = np.zeros(positions.shape[0], dtype=new_dtype)
    out["x"], out["y"], out["z"] = positions[:, 0], positions[:, 1], positions[:, 2]
    for key in ("intensity", "ring", "time", "return_type"):
        fname = meta.get(f"{key}_field_name")
        if meta.get(f"has_{key}") and key in attrs and fname is not None:
            out[fname] = np.asarray(attrs[key]).reshape(-1).astype(out[fname].dtype)
    if estimate_normals and normals is not None:
        out["normal_x"], out["normal_y"], out["normal_z"] = normals[:, 0], normals[:, 1], normals[:, 2]
    return out

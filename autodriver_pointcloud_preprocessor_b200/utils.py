"""Drop-in counterpart of the reference's ``autodriver_pointcloud_preprocessor/utils.py``.

Same module-level names, argument lists, return shapes and message strings (checked against
``tests/golden/utils_signatures.json``); the heavy lifting goes through the CUDA library
instead of numpy / torch / Open3D.  ``o3c`` / ``t`` / ``o3d`` are the carrier module of this
package (``geometry``) so call sites such as ``o3c.Tensor.from_numpy`` keep working.

Host-side helpers that are pure metadata or tiny numpy glue in the reference (field-name
mapping, packed field tables, rgb bit packing, timers) stay host-side here as well.
"""
from __future__ import annotations

import sys
import time
from typing import Any

import numpy as np

from . import _capi, geometry
from .msgs import PointCloud2, PointField  # noqa: F401  (same names the reference imports, utils.py:6)

try:
    import torch
    from torch.utils.dlpack import from_dlpack as torch_from_dlpack
    from torch.utils.dlpack import to_dlpack as torch_to_dlpack
except ImportError:  # pragma: no cover
    torch = None
    torch_from_dlpack = None
    torch_to_dlpack = None

# the reference's ``import open3d as o3d / open3d.core as o3c / open3d.t.geometry as t`` (utils.py:19-22)
o3d = geometry
o3c = geometry
t = geometry

FIELD_DTYPE_MAP = {                                   # utils.py:28-37
    PointField.INT8: np.int8,
    PointField.UINT8: np.uint8,
    PointField.INT16: np.int16,
    PointField.UINT16: np.uint16,
    PointField.INT32: np.int32,
    PointField.UINT32: np.uint32,
    PointField.FLOAT32: np.float32,
    PointField.FLOAT64: np.float64,
}

FIELD_DTYPE_MAP_INV = {v: k for k, v in FIELD_DTYPE_MAP.items()}

VENDOR_MAPPINGS = {                                   # utils.py:41-48
    "intensity": ["I", "intensity"],
    "ring": ["C", "ring", "line"],
    "time": ["t", "time", "timestamp"],
    "return_type": ["return_type", "tag", "R"],
    "azimuth": ["azimuth"],
    "distance": ["distance", "depth", "d"],
}

_TORCH_OF = {np.dtype(np.int8): "int8", np.dtype(np.uint8): "uint8", np.dtype(np.int16): "int16",
             np.dtype(np.uint16): "uint16", np.dtype(np.int32): "int32", np.dtype(np.uint32): "uint32",
             np.dtype(np.float32): "float32", np.dtype(np.float64): "float64"}


def _structured_dtype(fields, point_step):
    names, formats, offsets = [], [], []
    for f in fields:
        base = np.dtype(FIELD_DTYPE_MAP[f.datatype])
        if f.count == 1:
            names.append(f.name), formats.append(base), offsets.append(f.offset)
        else:
            for a in range(f.count):
                names.append(f"{f.name}_{a}"), formats.append(base), offsets.append(f.offset + a * base.itemsize)
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": point_step})


def convert_pointcloud_to_numpy(structured_cloud_array, metadata_dict):
    """utils.py:51-133: structured array -> SoA dict (host arrays) with the reference's casts.

    Kept for callers that already hold a structured host array; the node itself goes through
    :func:`pointcloud_to_dict`, which unpacks on the GPU.
    """
    has_rgb = metadata_dict.get('has_rgb', False)
    field_names = metadata_dict.get('field_names', ['x', 'y', 'z'])
    positions_arr = np.vstack(
        (structured_cloud_array["x"], structured_cloud_array["y"], structured_cloud_array["z"])
    ).T.astype(np.float32)
    pointcloud_dictionary = {'positions': positions_arr}
    if has_rgb:
        if {"r", "g", "b"}.issubset(field_names):
            rgb_arr = merge_rgb_fields(structured_cloud_array["r"], structured_cloud_array["g"],
                                       structured_cloud_array["b"], return_int=True)
        else:
            rgb_arr = extract_rgb_from_pointcloud(structured_cloud_array["rgb"].astype(np.float32))
        pointcloud_dictionary['rgb'] = rgb_arr
    for key, dtype in (('intensity', np.float32), ('ring', np.uint16), ('time', np.float64),
                       ('return_type', np.uint8)):
        if metadata_dict.get(f'has_{key}', False):
            pointcloud_dictionary[key] = structured_cloud_array[metadata_dict[f'{key}_field_name']].astype(dtype)
    return pointcloud_dictionary


def dict_to_open3d_tensor_pointcloud(pointcloud_dict, device="CPU:0"):
    """utils.py:135-137."""
    pointcloud = t.PointCloud(pointcloud_dict).to(device)
    return pointcloud


def numpy_struct_to_pointcloud2(field_names: list,
                                field_datatypes: list, is_dense: bool = True) -> tuple[list[Any], int | Any]:
    """utils.py:140-199: PointField list with cumulative offsets (no padding) and the point step."""
    fields = []
    offset = 0
    for name, datatype in zip(field_names, field_datatypes):
        np_dt = FIELD_DTYPE_MAP[datatype]
        byte_size = np.dtype(np_dt).itemsize
        pf = PointField()
        pf.name = name
        pf.offset = offset
        pf.datatype = datatype
        pf.count = 1
        fields.append(pf)
        offset += byte_size
    return fields, offset


def packed_message_bytes(ros_cloud):
    """The message's point records as one tightly packed host uint8 tensor ``[n * point_step]``.
    ``read_points`` honours ``row_step`` (organised clouds may pad their rows); the kernels expect
    ``width * height`` consecutive records, so padded rows are compacted here, on the way to the GPU."""
    n = ros_cloud.width * ros_cloud.height
    if n == 0:
        return torch.zeros(16, dtype=torch.uint8)
    raw = torch.frombuffer(bytearray(ros_cloud.data), dtype=torch.uint8)
    tight = ros_cloud.width * ros_cloud.point_step
    row_step = int(getattr(ros_cloud, "row_step", 0) or tight)
    if row_step != tight:
        if row_step < tight or raw.numel() < row_step * ros_cloud.height:
            raise ValueError(f"PointCloud2 row_step {row_step} is inconsistent with width*point_step {tight}")
        raw = raw[:row_step * ros_cloud.height].view(ros_cloud.height, row_step)[:, :tight].contiguous().view(-1)
    return raw


def is_foreign_endian(ros_cloud) -> bool:
    """read_points byte-swaps the records when the message's endianness differs from the host's
    (sensor_msgs_py.point_cloud2.read_points)."""
    return bool(sys.byteorder != 'little') != bool(ros_cloud.is_bigendian)


def device_native_endian(data_dev, ros_cloud):
    """Byte-swap every multi-byte field of the uploaded records IN PLACE on the device so that the kernels
    (which read native little-endian fields) see what ``read_points`` would return.  Off the hot path:
    sensors publish little-endian; a handful of strided device copies per message."""
    n = ros_cloud.width * ros_cloud.height
    if n == 0 or not is_foreign_endian(ros_cloud):
        return data_dev
    rows = data_dev[:n * ros_cloud.point_step].view(n, ros_cloud.point_step)
    for f in ros_cloud.fields:
        size = np.dtype(FIELD_DTYPE_MAP[f.datatype]).itemsize
        for c in range(max(1, int(getattr(f, 'count', 1) or 1))):
            a = f.offset + c * size
            if size > 1:
                rows[:, a:a + size] = rows[:, a:a + size].flip(1)
    return data_dev


def raw_column(rows, field):
    """One field of every record of a device byte buffer viewed as ``[n, point_step]``: a contiguous
    device tensor of the field's own dtype (the per-field slice ``read_points`` returns)."""
    np_dt = np.dtype(FIELD_DTYPE_MAP[field.datatype])
    col = rows[:, field.offset:field.offset + np_dt.itemsize].contiguous().view(getattr(torch, _TORCH_OF[np_dt]))
    return col.reshape(-1)


def pointcloud_to_dict(ros_cloud, field_names=None, skip_nans=True, organize_cloud=False, metadata_dict=None,
                       _data_dev=None):
    """utils.py:202-223, with ``read_points`` + ``convert_pointcloud_to_numpy`` executed on the GPU.

    The message bytes are uploaded once; positions / intensity come from the fused unpack
    kernel (NaN skip on every selected field iff ``skip_nans and not is_dense``), ring / time /
    return_type / rgb are cut out of the same device buffer and gathered with the surviving
    indices.  The returned dict holds device-resident carrier tensors (``.cpu().numpy()``
    gives the arrays the reference would have produced).
    """
    from . import engine
    if not metadata_dict:
        metadata_dict = {}
    metadata_dict.update({'header': ros_cloud.header, 'field_names': None})
    all_names = []
    for f in ros_cloud.fields:
        all_names += [f.name] if f.count == 1 else [f"{f.name}_{a}" for a in range(f.count)]
    if field_names is not None:
        assert all(name in all_names for name in field_names)
        names = tuple(field_names)
    else:
        names = tuple(all_names)
    metadata_dict['field_names'] = names
    metadata_dict['num_fields'] = len(names)
    if not metadata_dict.get('has_intensity', False):
        metadata_dict.update(get_pointcloud_metadata(metadata_dict['field_names']))
    n = ros_cloud.width * ros_cloud.height
    if _data_dev is not None:              # the node uploads the message once (already in native byte order) and shares the buffer
        data = _data_dev
    else:
        data = device_native_endian(packed_message_bytes(ros_cloud).cuda(), ros_cloud)
    ctx = geometry.get_context(n)
    desc = engine.make_cloud_desc(ros_cloud.fields, ros_cloud.point_step, n, data, field_names=field_names)
    cfg = engine.make_filter_cfg(skip_nans=bool(skip_nans and not ros_cloud.is_dense))
    xyzi, src, _, cnt = ctx.frontend([desc], cfg, want_src=True)
    m = int(cnt.item())
    pos, inten = ctx.split_xyzi(xyzi, m, want_intensity=bool(metadata_dict.get('has_intensity')))
    cloud_dict = {'positions': geometry.Tensor(pos)}
    if metadata_dict.get('has_intensity'):
        cloud_dict['intensity'] = geometry.Tensor(inten)
    src = src[:m].contiguous()
    by_name = {f.name: f for f in ros_cloud.fields}
    rows = data[:n * ros_cloud.point_step].view(n, ros_cloud.point_step) if n else None

    def column(name, np_dtype):
        return ctx.gather(raw_column(rows, by_name[name]), src, m).to(getattr(torch, _TORCH_OF[np.dtype(np_dtype)]))

    for key, np_dtype in (('ring', np.uint16), ('time', np.float64), ('return_type', np.uint8)):
        if metadata_dict.get(f'has_{key}') and n:
            cloud_dict[key] = geometry.Tensor(column(metadata_dict[f'{key}_field_name'], np_dtype))
    if metadata_dict.get('has_rgb') and n:
        if {"r", "g", "b"}.issubset(names):
            cols = [column(c, np.uint8) for c in ("r", "g", "b")]
            cloud_dict['rgb'] = geometry.Tensor(torch.stack(cols, 1))
        else:
            packed = column("rgb", np.float32).view(torch.int32)
            cloud_dict['rgb'] = geometry.Tensor(torch.stack([(packed >> 16) & 0xFF, (packed >> 8) & 0xFF, packed & 0xFF],
                                                            1).to(torch.uint8))
    if organize_cloud and ros_cloud.height > 1:
        metadata_dict['organized_shape'] = (ros_cloud.width, ros_cloud.height)
    return cloud_dict, metadata_dict


def check_field(field, pointcloud_dict, metadata_dict):
    """utils.py:226-229."""
    if pointcloud_dict.get(field, None) is not None or metadata_dict.get(f'has_{field}', None):
        return True
    return False


def get_fields_from_dicts(key, pointcloud_dict, metadata_dict):
    """utils.py:231-237: attribute -> (N, 1) carrier tensor."""
    key_tensor = None
    if check_field(key, pointcloud_dict, metadata_dict):
        value = pointcloud_dict[key]
        if isinstance(value, geometry.Tensor):
            key_tensor = value.reshape(-1, 1)
        else:
            key_tensor = o3c.Tensor.from_numpy(value.reshape(-1, 1))
    if key_tensor is not None:
        pointcloud_dict[key] = key_tensor
    return pointcloud_dict


def crop_pointcloud(pointcloud, backend='open3d', min_bound=None, max_bound=None, invert=False, aabb=None):
    """utils.py:240-301.  ``backend`` selects the *comparison semantics* of the reference's three
    branches (numpy: float64 compare; torch: float32 compare; anything else: Open3D AABB with
    ``invert`` = logical NOT); all three run in the same CUDA kernel."""
    msg = ''
    if backend.lower() in ['np', 'numpy']:
        if not pointcloud.point.positions.is_cpu:
            msg = f'{msg} Transferring points to cpu...'
        msg = f'{msg} Converting points to numpy...'
        crop_mask = pointcloud.crop_mask(min_bound, max_bound, _capi.CROP_NUMPY, invert)
        pointcloud = pointcloud.select_by_mask(crop_mask)
    elif backend.lower() in ['torch', 'pytorch']:
        crop_mask = pointcloud.crop_mask(min_bound, max_bound, _capi.CROP_TORCH, invert)
        pointcloud = pointcloud.select_by_mask(crop_mask)
    else:
        if aabb is None:
            aabb = t.AxisAlignedBoundingBox(min_bound, max_bound)
        pointcloud = pointcloud.crop(aabb, invert=invert)
        msg = f'{msg} Using Open3D pointcloud.crop()'
    return pointcloud, msg


def merge_rgb_fields(r, g, b, return_int=False):
    """utils.py:304-322."""
    if return_int:
        rgb_arr = np.vstack((r.astype(np.uint8), g.astype(np.uint8), b.astype(np.uint8))).T
    else:
        r, g, b = r.astype(np.uint32), g.astype(np.uint32), b.astype(np.uint32)
        rgb_arr = np.array((r << 16) | (g << 8) | (b << 0)).view(np.float32)
    return rgb_arr


def extract_rgb_from_pointcloud(rgb):
    """utils.py:324-345: packed float32 rgb -> (N, 3) uint8."""
    rgb_bytes = rgb.view(np.uint32)
    r = ((rgb_bytes >> 16) & 0xFF).astype(np.uint8)
    g = ((rgb_bytes >> 8) & 0xFF).astype(np.uint8)
    b = (rgb_bytes & 0xFF).astype(np.uint8)
    return np.vstack((r, g, b)).T.astype(np.uint8)


def rgb_int_to_float(rgb_np):
    """utils.py:347-356."""
    colors_u8 = (rgb_np * 255).clip(0, 255).astype(np.uint8)
    r_u, g_u, b_u = (colors_u8[:, c].astype(np.uint32) for c in range(3))
    return ((r_u << 16) | (g_u << 8) | b_u).view(np.float32)


def rgb_to_intensity(color):
    """utils.py:358-367."""
    rgb = np.asarray(color)
    return (0.2126 * rgb[:, 0] + 0.7152 * rgb[:, 1] + 0.0722 * rgb[:, 2]).astype(np.float32)


def intensity_to_rgb(intensity):
    """utils.py:370-421: min-max normalised grey colours as a carrier tensor."""
    intensity = intensity.astype(np.float32)
    i_min, i_max = intensity.min(), intensity.max()
    i_norm = (intensity - i_min) / max(i_max - i_min, 1e-6)
    rgb = np.stack([i_norm, i_norm, i_norm], axis=1).astype(np.float32)
    return o3d.Tensor(rgb)


def parse_differing_fields(options, field_names):
    """utils.py:423-438: last matching alias wins and its spelling is returned."""
    if isinstance(options, str):
        options = [options]
    option_in_field_names = []
    corresponding_field_name = None
    for option in options:
        if option.lower() in field_names:
            option_in_field_names.append(option)
            corresponding_field_name = option
    return any(option_in_field_names), corresponding_field_name


def get_pointcloud_metadata(field_names, vendor_mappings: dict = None):
    """utils.py:441-472."""
    if vendor_mappings is None:
        vendor_mappings = VENDOR_MAPPINGS
    field_names = [field_name.lower() for field_name in field_names]
    if {"r", "g", "b"}.issubset(field_names):
        has_rgb = True
    else:
        has_rgb, _ = parse_differing_fields("rgb", field_names)
    has_intensity, intensity_field_name = parse_differing_fields(vendor_mappings["intensity"], field_names)
    has_ring, ring_field_name = parse_differing_fields(vendor_mappings["ring"], field_names)
    has_time, time_field_name = parse_differing_fields(vendor_mappings["time"], field_names)
    has_return_type, return_type_field_name = parse_differing_fields(vendor_mappings["return_type"], field_names)
    return {
        'has_rgb': has_rgb,
        'has_intensity': has_intensity,
        'intensity_field_name': intensity_field_name,
        'has_ring': has_ring,
        'ring_field_name': ring_field_name,
        'has_time': has_time,
        'time_field_name': time_field_name,
        'has_return_type': has_return_type,
        'return_type_field_name': return_type_field_name,
    }


def get_current_time(monotonic=True):
    """utils.py:474-483."""
    if not monotonic:
        return time.time()
    return time.perf_counter()


def get_time_difference(start_time, end_time, return_absolute_difference=False):
    """utils.py:486-500."""
    time_difference = end_time - start_time
    if return_absolute_difference:
        return abs(end_time - start_time)
    return time_difference


def structured_numpy_array_to_open3d_tensor_pointcloud(structured_numpy_array):
    """utils.py:503-506 (the reference assigns ``.points``, which Open3D's tensor cloud does not
    have; here the positions land where every other function expects them)."""
    pointcloud = o3d.PointCloud()
    pointcloud.point["positions"] = o3c.Tensor.from_numpy(structured_numpy_array["positions"])
    return pointcloud


def remove_duplicates(pointcloud, backend='torch'):
    """utils.py:509-546, every back end on the device.

    * ``'np'`` / ``'numpy'`` -> ``np.unique(points, axis=0, return_index=True)`` then
      ``select_by_index(first_index)`` (utils.py:532-534): the *lexicographically sorted* unique
      rows, each represented by its lowest input index; -0.0 == +0.0 merge, NaN rows never do.
      Here a stable 96-bit-key radix sort + head flags (``apc_unique_rows``).
    * ``'torch'`` / ``'pytorch'`` -> the reference passes ``torch.unique``'s *inverse* map to
      ``select_by_index`` (utils.py:538-542), so the result has N rows ``points[inverse]``.
      Reproduced as written (same sort, inverse map); identical to torch on NaN-free input, which
      is what reaches this stage once read_points has skipped NaN rows.
    * anything else -> ``remove_duplicated_points()``: bit-pattern keys, lowest index kept, order
      preserved (GPU hash).
    """
    msg = ''
    if backend.lower() in ['np', 'numpy']:
        first_index = pointcloud.unique_rows_index()
        pointcloud = pointcloud.select_by_index(first_index)
    elif backend.lower() in ['torch', 'pytorch']:
        inverse = pointcloud.unique_rows_inverse()
        pointcloud = pointcloud.select_by_index(inverse)
    else:
        pointcloud, duplicates_mask = pointcloud.remove_duplicated_points()
        msg = f'{msg} Using Open3D pointcloud.remove_duplicated_points()'
    return pointcloud, msg

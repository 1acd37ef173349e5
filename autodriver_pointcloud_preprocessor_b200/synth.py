"""Seeded synthetic spinning-LiDAR scans and PointCloud2 byte-buffer builders.

The reference ships no data (``setup.py:55`` globs a ``data/`` directory that does not
exist), so the workloads named in BASELINE.json are generated here: a ray-cast scene
(ground plane, an enclosing box, a few random boxes and cylinders) sampled on an
``n_beams x n_az`` spinning pattern in azimuth-major order, the order LiDAR drivers emit.
Everything is numpy on the host; this module is test/bench input generation, not part of
the device hot path.

Layouts:
  * ``xyzi16``  - x,y,z,intensity float32, point_step 16 (128-bit aligned primary layout)
  * ``xyzirt22`` - Velodyne: x,y,z,intensity f32; ring u16 @16; time f32 @18; point_step 22
  * ``ouster48`` - Ouster-like padded: x,y,z f32 @0/4/8; intensity f32 @16; t u32 @20;
                   reflectivity u16 @24; ring u16 @26; ambient u16 @28; range u32 @32;
                   point_step 48
"""
from __future__ import annotations

import numpy as np

from .msgs import Header, PointCloud2, PointField

LAYOUTS = ("xyzi16", "xyzirt22", "ouster48")


def _layout_dtype(layout: str) -> np.dtype:
    if layout == "xyzi16":
        return np.dtype({"names": ["x", "y", "z", "intensity"],
                         "formats": ["<f4", "<f4", "<f4", "<f4"],
                         "offsets": [0, 4, 8, 12], "itemsize": 16})
    if layout == "xyzirt22":
        return np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                         "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                         "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22})
    if layout == "ouster48":
        return np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring",
                                   "ambient", "range"],
                         "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "<u2", "<u2", "<u4"],
                         "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48})
    raise ValueError(f"unknown layout {layout!r}")


_NP_TO_PF = {np.dtype("i1"): PointField.INT8, np.dtype("u1"): PointField.UINT8,
             np.dtype("<i2"): PointField.INT16, np.dtype("<u2"): PointField.UINT16,
             np.dtype("<i4"): PointField.INT32, np.dtype("<u4"): PointField.UINT32,
             np.dtype("<f4"): PointField.FLOAT32, np.dtype("<f8"): PointField.FLOAT64}


def fields_from_dtype(dtype: np.dtype):
    """PointField list describing a structured dtype (offsets preserved)."""
    out = []
    for name in dtype.names:
        sub, off = dtype.fields[name][:2]
        out.append(PointField(name=name, offset=int(off), datatype=_NP_TO_PF[np.dtype(sub)], count=1))
    return out


def make_scene(rng: np.random.Generator, n_boxes: int = 6, n_cyl: int = 6):
    """Random obstacles inside the 40 m enclosing box."""
    boxes = []
    for _ in range(n_boxes):
        c = np.array([rng.uniform(-30, 30), rng.uniform(-30, 30)])
        if np.linalg.norm(c) < 4.0:
            c += 6.0
        half = rng.uniform(0.5, 3.0, size=2)
        h = rng.uniform(0.5, 4.0)
        boxes.append((c[0] - half[0], c[0] + half[0], c[1] - half[1], c[1] + half[1], -1.8, -1.8 + h))
    cyls = []
    for _ in range(n_cyl):
        c = np.array([rng.uniform(-25, 25), rng.uniform(-25, 25)])
        if np.linalg.norm(c) < 3.0:
            c += 5.0
        cyls.append((c[0], c[1], rng.uniform(0.15, 0.8), -1.8 + rng.uniform(1.0, 6.0)))
    return boxes, cyls


def _ray_cast(d: np.ndarray, boxes, cyls) -> np.ndarray:
    """Nearest hit range along unit directions ``d`` (R,3) from the origin (float64)."""
    big = 1e9
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        # ground plane z = -1.8
        t = np.where(dz < -1e-9, -1.8 / dz, big)
        # enclosing box walls |x|,|y| = 40, ceiling z = 15
        tx = np.where(np.abs(dx) > 1e-12, 40.0 / np.abs(dx), big)
        ty = np.where(np.abs(dy) > 1e-12, 40.0 / np.abs(dy), big)
        tz = np.where(dz > 1e-9, 15.0 / dz, big)
        t = np.minimum(np.minimum(t, tx), np.minimum(ty, tz))
        inv = 1.0 / d
        for (x0, x1, y0, y1, z0, z1) in boxes:
            lo = np.stack([x0 * inv[:, 0], y0 * inv[:, 1], z0 * inv[:, 2]], axis=1)
            hi = np.stack([x1 * inv[:, 0], y1 * inv[:, 1], z1 * inv[:, 2]], axis=1)
            tmin = np.nanmax(np.minimum(lo, hi), axis=1)
            tmax = np.nanmin(np.maximum(lo, hi), axis=1)
            hit = (tmax >= tmin) & (tmin > 0)
            t = np.where(hit & (tmin < t), tmin, t)
        for (cx, cy, r, ztop) in cyls:
            a = dx * dx + dy * dy
            b = -2.0 * (dx * cx + dy * cy)
            c = cx * cx + cy * cy - r * r
            disc = b * b - 4 * a * c
            tc = (-b - np.sqrt(np.maximum(disc, 0.0))) / (2 * a)
            zc = tc * dz
            hit = (disc > 0) & (tc > 0) & (zc <= ztop) & (zc >= -1.8)
            t = np.where(hit & (tc < t), tc, t)
    return t


def lidar_scan(seed: int = 0, n_beams: int = 128, n_az: int = 2048, nan_frac: float = 0.005,
               dup_frac: float = 0.001, noise_sigma: float = 0.02, n_points: int | None = None):
    """One synthetic sweep.  Returns a dict of SoA numpy arrays in azimuth-major order.

    ``positions`` float32 (N,3); ``intensity`` float32; ``ring`` uint16; ``time`` float32.
    ``nan_frac`` of the returns are NaN (no echo) and ``dup_frac`` are exact copies of an
    earlier point (drivers that re-emit a return), as SURVEY.md section 8d prescribes.
    """
    rng = np.random.default_rng(seed)
    boxes, cyls = make_scene(rng)
    elev = np.deg2rad(np.linspace(-22.5, 22.5, n_beams))
    az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False)
    azg, elg = np.meshgrid(az, elev, indexing="ij")          # (n_az, n_beams): azimuth-major
    ce = np.cos(elg)
    d = np.stack([ce * np.cos(azg), ce * np.sin(azg), np.sin(elg)], axis=-1).reshape(-1, 3)
    rngs = _ray_cast(d, boxes, cyls)
    rngs = rngs + rng.normal(0.0, noise_sigma, size=rngs.shape)
    pos = (d * rngs[:, None]).astype(np.float32)
    n = pos.shape[0]
    ring = np.tile(np.arange(n_beams, dtype=np.uint16), n_az)
    tcol = np.repeat((np.arange(n_az, dtype=np.float64) / n_az * 0.1), n_beams).astype(np.float32)
    inten = rng.uniform(0.0, 255.0, size=n).astype(np.float32)
    if n_points is not None:                                   # truncate / tile to an exact size
        reps = -(-n_points // n)
        if reps > 1:
            jitter = rng.normal(0.0, 0.01, size=(reps * n, 3)).astype(np.float32)
            pos = np.tile(pos, (reps, 1)) + jitter
            ring, tcol, inten = np.tile(ring, reps), np.tile(tcol, reps), np.tile(inten, reps)
        pos, ring, tcol, inten = pos[:n_points], ring[:n_points], tcol[:n_points], inten[:n_points]
        n = n_points
    pos = np.ascontiguousarray(pos)
    if dup_frac > 0:
        k = int(n * dup_frac)
        dst = rng.choice(np.arange(1, n), size=k, replace=False)
        src = (dst * rng.uniform(0.0, 1.0, size=k)).astype(np.int64)   # an earlier point
        pos[dst] = pos[src]
    if nan_frac > 0:
        k = int(n * nan_frac)
        idx = rng.choice(n, size=k, replace=False)
        pos[idx] = np.nan
    return {"positions": pos, "intensity": inten, "ring": ring, "time": tcol}


def pack_cloud(scan: dict, layout: str = "xyzi16", frame_id: str = "lidar", is_dense: bool = False,
               stamp=None) -> PointCloud2:
    """Pack a SoA scan into a PointCloud2 message with the given byte layout."""
    dt = _layout_dtype(layout)
    pos = scan["positions"]
    n = pos.shape[0]
    arr = np.zeros(n, dtype=dt)
    arr["x"], arr["y"], arr["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
    arr["intensity"] = scan["intensity"]
    if "ring" in dt.names:
        arr["ring"] = scan["ring"]
    if "time" in dt.names:
        arr["time"] = scan["time"]
    if "t" in dt.names:
        arr["t"] = (scan["time"].astype(np.float64) * 1e9).astype(np.uint32)
    if "reflectivity" in dt.names:
        arr["reflectivity"] = scan["intensity"].astype(np.uint16)
    if "range" in dt.names:
        with np.errstate(invalid="ignore"):
            r = np.nan_to_num(np.linalg.norm(pos.astype(np.float64), axis=1) * 1000.0, nan=0.0)
        arr["range"] = r.astype(np.uint32)
    msg = PointCloud2(header=Header(stamp=stamp, frame_id=frame_id), height=1, width=n,
                      fields=fields_from_dtype(dt), is_bigendian=False, point_step=dt.itemsize,
                      row_step=dt.itemsize * n, data=arr.tobytes(), is_dense=is_dense)
    return msg


def sensor_extrinsics(n_sensors: int = 4) -> np.ndarray:
    """Distinct rigid float32 4x4 transforms: yaw 0/90/180/270 deg, +-1 m offsets (config C3)."""
    out = np.zeros((n_sensors, 4, 4), dtype=np.float64)
    for s in range(n_sensors):
        yaw = np.deg2rad(90.0 * s + 3.0 * s)
        c, si = np.cos(yaw), np.sin(yaw)
        out[s] = np.array([[c, -si, 0.0, 1.0 if s % 2 == 0 else -1.0],
                           [si, c, 0.0, 1.0 if s < 2 else -1.0],
                           [0.0, 0.0, 1.0, 0.1 * s],
                           [0.0, 0.0, 0.0, 1.0]])
    return out.astype(np.float32)


#: the five BASELINE.json configurations (sizes and stage parameters)
CONFIGS = {
    "C1": dict(n_beams=64, n_az=2048, voxel_size=0.1, remove_ground=True),
    "C2": dict(n_beams=128, n_az=2048, voxel_size=0.1, remove_ground=True, radius_outliers=True),
    "C3": dict(n_beams=128, n_az=2048, n_sensors=4, voxel_size=0.1),
    "C4": dict(n_points=1_500_000, voxel_size=0.05, statistical_outliers=True),
    "C5": dict(n_beams=128, n_az=2048, n_frames=1024),
}

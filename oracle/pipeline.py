"""Multi-sensor concatenation and the full per-scan pipeline: CPU oracle (test infrastructure).

``concat`` defines the semantics ``pointcloud_concatenator.py:1-5`` only describes in prose
(no code exists in the reference): for sensors i = 1..S with float32 4x4 ``T_i`` the output
is the concatenation, in sensor order with each sensor's point order preserved, of
``transform(T_i, P_i)`` (same float32 arithmetic as ``filters.transform``), attributes
carried along.

``preprocess`` follows the stage order of ``pp.py:447-544``:
dedup -> non-finite -> transform(s) -> crop -> voxel -> statistical outliers ->
[radius outliers - absent from the reference, placed after the statistical stage] ->
normals -> RANSAC ground removal.
"""
from __future__ import annotations

import numpy as np

from . import dedup as odedup
from . import filters, normals, outliers, pc2, ransac, voxel


def concat(clouds, transforms):
    """``clouds``: list of dicts with 'positions' (+ optional 'intensity'); returns one dict
    plus ``sensor_id`` (uint8) and ``src_idx`` (index within the sensor)."""
    pos, inten, sid, src = [], [], [], []
    for s, (c, T) in enumerate(zip(clouds, transforms)):
        p = c["positions"].astype(np.float32)
        pos.append(filters.transform(p, T) if T is not None else p)
        inten.append(c.get("intensity", np.zeros(p.shape[0], np.float32)).astype(np.float32))
        sid.append(np.full(p.shape[0], s, dtype=np.uint8))
        src.append(np.arange(p.shape[0], dtype=np.uint32))
    return {"positions": np.concatenate(pos), "intensity": np.concatenate(inten),
            "sensor_id": np.concatenate(sid), "src_idx": np.concatenate(src)}


def default_config():
    """Stage switches with the reference's defaults (pp.py:165-185)."""
    return dict(
        skip_nans=True, dedup_mode=odedup.DEDUP_OPEN3D, remove_nans=True, remove_infs=True,
        transforms=(), crop=dict(min=[-60.0, -60.0, -20.0], max=[60.0, 60.0, 20.0], invert=False,
                                 mode=filters.CROP_OPEN3D),
        voxel_size=0.01,
        statistical=None,            # dict(nb_neighbors=20, std_ratio=2.0)
        radius=None,                 # dict(nb_points=5, radius=0.5)
        normals=None,                # dict(radius=0.1, max_nn=30)  (pp.py:521-530; on by default in the node)
        ground=None,                 # dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=0)
    )


def preprocess(cloud_msgs, cfg, per_sensor_transforms=None):
    """Run the pipeline on one PointCloud2 (or a list for multi-sensor concat).

    Returns a dict with the final ``positions`` / ``intensity`` and the intermediates the
    parity tests compare (``stage_mask``, ``src_idx``, ``p2v``, masks, plane, ...).
    """
    if not isinstance(cloud_msgs, (list, tuple)):
        cloud_msgs = [cloud_msgs]
    per_sensor_transforms = per_sensor_transforms or [()] * len(cloud_msgs)
    out = {}
    pos_all, int_all, stage_all, src_all = [], [], [], []
    base = 0
    for msg, sensor_T in zip(cloud_msgs, per_sensor_transforms):
        arr = pc2.read_points(msg, skip_nans=False)
        n = arr.shape[0]
        nanskip = pc2.read_points_mask(msg, skip_nans=cfg["skip_nans"])
        pos = np.vstack((arr["x"], arr["y"], arr["z"])).T.astype(np.float32)
        meta = pc2.get_pointcloud_metadata(arr.dtype.names)
        inten = (arr[meta["intensity_field_name"]].astype(np.float32) if meta["has_intensity"]
                 else np.zeros(n, np.float32))
        dd = None
        if cfg["dedup_mode"] in (odedup.DEDUP_NUMPY, odedup.DEDUP_TORCH_COMPAT):
            if len(cloud_msgs) != 1 or len(sensor_T):
                raise NotImplementedError("the sorted dedup modes are defined for one sensor without a per-sensor transform")
            fn = odedup.numpy_index if cfg["dedup_mode"] == odedup.DEDUP_NUMPY else odedup.torch_compat_index_numpy
            p, src = filters.frontend_sorted(
                pos, fn, nanskip_mask=nanskip, remove_nan=cfg["remove_nans"], remove_infinite=cfg["remove_infs"],
                transforms=list(cfg["transforms"]), crop=cfg["crop"])
            stage = np.zeros(n, dtype=np.uint8)     # per-point stage bits are not defined once rows are reordered
        else:
            if cfg["dedup_mode"] == odedup.DEDUP_OPEN3D:
                dd = odedup.open3d_mask
            p, src, stage = filters.frontend(
                pos, nanskip_mask=nanskip, dedup_mask_fn=dd, remove_nan=cfg["remove_nans"],
                remove_infinite=cfg["remove_infs"], transforms=list(sensor_T) + list(cfg["transforms"]),
                crop=cfg["crop"])
        pos_all.append(p)
        int_all.append(inten[src])
        stage_all.append(stage)
        src_all.append(src + np.uint32(base))
        base += n
    pos = np.concatenate(pos_all)
    inten = np.concatenate(int_all)
    out["stage_mask"] = np.concatenate(stage_all)
    out["src_idx"] = np.concatenate(src_all)
    out["n_filtered"] = pos.shape[0]
    out["filtered_positions"], out["filtered_intensity"] = pos, inten

    if cfg["voxel_size"] and cfg["voxel_size"] > 0.0:
        v = voxel.voxel_down_sample(pos, cfg["voxel_size"], inten, fixed=True)
        pos, inten = v["positions"], v["intensity"]
        out["p2v"], out["voxel_counts"] = v["p2v"], v["counts"]
        out["voxel_positions"], out["voxel_intensity"] = pos, inten
    if cfg.get("statistical"):
        s = cfg["statistical"]
        m, avg = outliers.statistical_mask(pos, s["nb_neighbors"], s["std_ratio"])
        out["statistical_mask"], out["statistical_avg"] = m, avg
        pos, inten = pos[m], inten[m]
    if cfg.get("radius"):
        r = cfg["radius"]
        m = outliers.radius_mask(pos, r["nb_points"], r["radius"])
        out["radius_mask"] = m
        pos, inten = pos[m], inten[m]
    nrm = ncov = None
    if cfg.get("normals"):
        nrm, out["normal_counts"], ncov = normals.estimate_normals(pos, cfg["normals"]["radius"], cfg["normals"]["max_nn"])
    if cfg.get("ground"):
        g = cfg["ground"]
        plane, inl, info = ransac.segment_plane(pos, g["distance_threshold"], g["ransac_n"],
                                                g["num_iterations"], g["probability"], g.get("seed", 0))
        out["plane"], out["ground_inliers"], out["ransac_info"] = plane, inl, info
        keep = np.ones(pos.shape[0], dtype=bool)
        keep[inl] = False
        pos, inten = pos[keep], inten[keep]
        nrm = nrm[keep] if nrm is not None else None
        ncov = ncov[keep] if ncov is not None else None
    if nrm is not None:
        out["normals"], out["normal_cov"] = nrm, ncov
    out["positions"], out["intensity"] = pos, inten
    return out

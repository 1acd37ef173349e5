"""Wall time of PointcloudPreprocessorNode.callback (message bytes in -> published message out)
on C2-size scans: fused path (xyz+intensity layouts) vs staged carrier path (ring/time layouts),
with and without normal estimation."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from autodriver_pointcloud_preprocessor_b200 import synth
from autodriver_pointcloud_preprocessor_b200.pointcloud_preprocessor import PointcloudPreprocessorNode

scan = synth.lidar_scan(seed=3, n_beams=128, n_az=2048)
for layout in ("xyzi16", "xyzirt22"):
    msg = synth.pack_cloud(scan, layout, frame_id="lidar")
    for normals in (False, True):
        for fused in ("auto", "true"):
            node = PointcloudPreprocessorNode(parameter_overrides={
                "use_gpu": True, "voxel_size": 0.1, "remove_ground": True, "remove_radius_outliers": True,
                "estimate_normals": normals, "estimate_normals.search_radius": 0.5, "fused_pipeline": fused})
            ts = []
            for it in range(12):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                node.callback(msg)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            out = node.pointcloud_pub.messages[-1]
            pt = node.processing_times
            print(f"{layout:9s} normals={normals!s:5s} fused={fused:5s}: callback p50 {np.median(ts[2:]):7.2f} ms  "
                  f"-> {out.width} pts, step {out.point_step};  preprocess {pt.get('preprocessing_time', 0) * 1e3:.2f} ms, "
                  f"parse {pt.get('pointcloud_msg_parsing', 0) * 1e3:.2f} ms, extract {pt.get('ros_to_numpy', 0) * 1e3:.2f} ms")

"""C4 stage driver for ncu: 0.05 m voxels + statistical outlier removal (k = 20) on a 1.5 M-point scan."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodriver_pointcloud_preprocessor_b200 import _capi, engine, synth
ctx = engine.Context(max_points=1_600_000)
scan = synth.lidar_scan(seed=5, n_points=1_500_000, nan_frac=0.0, dup_frac=0.0)
m = synth.pack_cloud(scan, "xyzi16")
buf = torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda()
desc = engine.make_cloud_desc(m.fields, m.point_step, m.width, buf)
pcfg = engine.make_pipeline_cfg(engine.make_filter_cfg(), voxel_size=0.05, statistical=dict(nb_neighbors=20, std_ratio=2.0))
for _ in range(2):
    out, counts, plane = ctx.pipeline_run([desc], pcfg)
    ctx.check()
    print(counts.cpu().numpy().tolist())

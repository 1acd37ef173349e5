"""The reference node's DEFAULT parameter set (pp.py:165-185: duplicate removal, non-finite filter, ROI crop,
0.01 m voxels, normal estimation radius 0.1 m / max_nn 30; no outlier or ground stage) on a 262 144-point scan
as one captured graph: p50 of the replay and the eager per-kernel CUDA-event times.  VOXEL / RADIUS / MAX_NN
override the three numbers."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from autodriver_pointcloud_preprocessor_b200 import _capi, engine
vox, rad, nn = float(os.environ.get("VOXEL", "0.01")), float(os.environ.get("RADIUS", "0.1")), int(os.environ.get("MAX_NN", "30"))
m = bench.make_frames(1, seed0=3)[0]
ctx = engine.Context(max_points=m.width)
buf = torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda()
desc = engine.make_cloud_desc(m.fields, m.point_step, m.width, buf)
fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True, crop=bench.CROP)
pcfg = engine.make_pipeline_cfg(fcfg, voxel_size=vox, normals=dict(radius=rad, max_nn=nn))
ctx.profile(True); out, counts, plane, maps = ctx.pipeline_run_maps([desc], pcfg); rep = ctx.profile_report(); ctx.profile(False)
ctx.check()
o2 = torch.zeros_like(out); c2 = torch.zeros_like(counts); p2 = torch.zeros_like(plane)
g = ctx.capture_pipeline([desc], pcfg, o2, c2, p2, maps=maps)
lat = []
for _ in range(23):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ctx.launch_graph(g); b.record(); b.synchronize(); lat.append(a.elapsed_time(b) * 1e3)
ctx.check()
c = c2.cpu().numpy()
nrm = maps["normals"][:int(c[_capi.CNT_OUTPUT])]
print(f"default config voxel {vox} normals({rad}, {nn}): p50 {np.median(lat[3:]):8.1f} us  in {c[_capi.CNT_INPUT]} -> voxels {c[_capi.CNT_VOXELS]} -> out {c[_capi.CNT_OUTPUT]}"
      f"  normals checksum {float(nrm.double().abs().sum()):.6f}")
for k, (ms, n) in sorted(rep.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"      {k:20s} {ms / n * 1e3:8.1f} us x{n}")

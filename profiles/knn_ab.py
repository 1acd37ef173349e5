"""C4 (1.5 M points, 0.05 m voxels, statistical outlier removal k = 20): p50 of the captured graph and the
per-kernel CUDA-event times, for the KNN knobs of the environment this process was started with."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodriver_pointcloud_preprocessor_b200 import _capi, engine, synth
ctx = engine.Context(max_points=1_600_000)
scan = synth.lidar_scan(seed=5, n_points=1_500_000, nan_frac=0.0, dup_frac=0.0)
m = synth.pack_cloud(scan, "xyzi16")
buf = torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda()
desc = engine.make_cloud_desc(m.fields, m.point_step, m.width, buf)
pcfg = engine.make_pipeline_cfg(engine.make_filter_cfg(), voxel_size=0.05, statistical=dict(nb_neighbors=int(os.environ.get("K", "20")), std_ratio=2.0))
out = torch.zeros((m.width, 4), device="cuda"); counts = torch.zeros(8, dtype=torch.int32, device="cuda"); plane = torch.zeros(8, dtype=torch.float64, device="cuda")
ctx.profile(True); ctx.pipeline_run([desc], pcfg, out, counts, plane); rep = ctx.profile_report(); ctx.profile(False)
g = ctx.capture_pipeline([desc], pcfg, out, counts, plane)
lat = []
for _ in range(15):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ctx.launch_graph(g); b.record(); b.synchronize(); lat.append(a.elapsed_time(b) * 1e3)
ctx.check()
c = counts.cpu().numpy()
print(f"fill={os.environ.get('APC_KNN_START_FILL','default')} margin={os.environ.get('APC_KNN_MARGIN','1')} p50 {np.median(lat[3:]):8.1f} us  out {c[_capi.CNT_OUTPUT]} of {c[_capi.CNT_VOXELS]}  checksum {int(out[:int(c[_capi.CNT_OUTPUT])].view(torch.int32).to(torch.int64).sum())}  "
      + "  ".join(f"{k} {ms / n * 1e3:.0f}" for k, (ms, n) in sorted(rep.items(), key=lambda kv: -kv[1][0])[:4]))

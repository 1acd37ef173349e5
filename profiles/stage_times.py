"""Per-kernel CUDA-event timings of the C2 pipeline (eager launches, mean over a few scans)
plus the p50 of the captured graph - a quick A/B tool for kernel changes.

    python profiles/stage_times.py [n_frames]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autodriver_pointcloud_preprocessor_b200 import _capi, replay  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2:                    # azimuth steps per scan: 2048 = C2 (262144 points), 8192 = 1 Mi points
    bench.N_AZ = int(sys.argv[2])
    bench.N_POINTS = bench.N_BEAMS * bench.N_AZ
msgs = bench.make_frames(n_frames, seed0=0)
filter_kw = dict(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                 transforms=[bench.TF], crop=bench.CROP)
pipe = replay.ScanPipeline(msgs[0].fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, bench.STAGES, lanes=1)
ln = pipe.lanes[0]
frames = [torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda() for m in msgs]
prof = {}
for rep in range(3):
    for f in frames:
        ln.d_in.copy_(f)
        for k, (ms, cnt) in pipe.stage_profile().items():
            if rep:                      # first repetition is warm-up
                p = prof.setdefault(k, [0.0, 0])
                p[0] += ms
                p[1] += cnt
lat = []
with torch.cuda.stream(ln.stream):
    for rep in range(5):
        for f in frames:
            ln.d_in.copy_(f, non_blocking=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ln.stream)
            ln.ctx.launch_graph(ln.graph)
            b.record(ln.stream)
            b.synchronize()
            lat.append(a.elapsed_time(b) * 1e3)
tot = sum(v[0] / v[1] for v in prof.values())
print(f"graph p50 {np.median(lat):.1f} us   sum of kernels {tot * 1e3:.1f} us   kernels/scan {pipe.kernels_per_scan}")
counts = ln.d_counts.cpu().numpy().astype(float)
cnt = {"N": counts[_capi.CNT_INPUT], "ps": float(bench.POINT_STEP), "M": counts[_capi.CNT_FILTERED], "V": counts[_capi.CNT_VOXELS],
       "Ps": counts[_capi.CNT_AFTER_STAT], "Pr": counts[_capi.CNT_AFTER_RADIUS], "K": counts[_capi.CNT_GROUND_INLIERS],
       "O": counts[_capi.CNT_OUTPUT]}
print(f"points per scan {bench.N_POINTS}: " + ", ".join(f"{k}={int(v)}" for k, v in cnt.items()))
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    kb = bench.kernel_bytes(k, cnt)
    ab = kb[0] if kb else None
    gbs = f"{ab / (ms / n * 1e-3) / 1e9:8.1f} GB/s algorithmic" if ab else ""
    print(f"  {k:22s} {ms / n * 1e3:8.2f} us {gbs}")

"""Small driver for ncu: a few eager (non-graph) runs of the C2 pipeline on 262144-point
scans so that every kernel of the per-scan path appears as an individual launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autodriver_pointcloud_preprocessor_b200 import _capi, engine  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
msgs = bench.make_frames(4, seed0=0)
ctx = engine.Context(max_points=bench.N_POINTS)
fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                              transforms=[bench.TF], crop=bench.CROP)
pcfg = engine.make_pipeline_cfg(fcfg, **bench.STAGES)
for f in range(n_frames):
    m = msgs[f % len(msgs)]
    data = torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda()
    desc = engine.make_cloud_desc(m.fields, m.point_step, m.width, data)
    out, counts, plane = ctx.pipeline_run([desc], pcfg)
    ctx.check()
    print(f, counts.cpu().numpy().tolist())

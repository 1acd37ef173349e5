"""CTA timelines (globaltimer stamps) of the pipeline kernels; needs a build with APC_TRACE=1:

    APC_TRACE=1 python -c "import __graft_entry__ as g; g.build()"; python profiles/cta_trace.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autodriver_pointcloud_preprocessor_b200 import _capi, engine  # noqa: E402

msgs = bench.make_frames(2, seed0=0)
ctx = engine.Context(max_points=bench.N_POINTS)
fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                              transforms=[bench.TF], crop=bench.CROP)
pcfg = engine.make_pipeline_cfg(fcfg, **bench.STAGES)
lib = ctypes.CDLL(_capi.LIB_PATH)
names = {"frontend": ["k_dedup_insert", "k_frontend", "k_select_by_mask"], "voxel": ["k_voxel_insert", "k_voxel_finalize"],
         "neighbors": ["k_radius_query"]}
buf = (ctypes.c_uint64 * (8 * 2048 * 4))()
for f in range(4):
    for tu in names:
        getattr(lib, f"apc_debug_trace_{tu}")(buf)       # clears
    m = msgs[f % 2]
    data = torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda()
    desc = engine.make_cloud_desc(m.fields, m.point_step, m.width, data)
    out, counts, plane = ctx.pipeline_run([desc], pcfg)
    ctx.check()
np.set_printoptions(linewidth=220)
for tu, kn in names.items():
    getattr(lib, f"apc_debug_trace_{tu}")(buf)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(8, 2048, 4).astype(np.int64)
    for kid, name in enumerate(kn):
        t = a[kid]
        t = t[t[:, 0] > 0]
        if not len(t):
            continue
        t0 = t[:, 0].min()
        ncol = int((t > 0).all(axis=0).sum()) if name != "k_dedup_insert" else 3
        print(f"== {name}: {len(t)} CTAs; columns = stamps relative to the first CTA start (ns)")
        for c in range(ncol):
            x = t[:, c] - t0
            print(f"   stamp{c}: min {x.min():7d}  p50 {int(np.median(x)):7d}  p90 {int(np.percentile(x, 90)):7d}  max {x.max():7d}")
        idx = np.linspace(0, len(t) - 1, 9).astype(int)
        print("   sample CTAs", idx.tolist())
        print((t[idx, :ncol] - t0).T)

"""Timeline of one captured-graph replay of the C2 pipeline: first-CTA start and last-CTA end of
every kernel (globaltimer stamps; build with APC_TRACE=1), i.e. kernel durations AND the gaps
between dependent kernels inside the graph.

    APC_TRACE=1 python -c "import __graft_entry__ as g; g.build()"; python profiles/graph_timeline.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autodriver_pointcloud_preprocessor_b200 import _capi, replay  # noqa: E402

msgs = bench.make_frames(4, seed0=0)
filter_kw = dict(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                 transforms=[bench.TF], crop=bench.CROP)
pipe = replay.ScanPipeline(msgs[0].fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, bench.STAGES, lanes=1)
ln = pipe.lanes[0]
frames = [torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda() for m in msgs]
lib = ctypes.CDLL(_capi.LIB_PATH)
names = {"ctx": ["k_begin"], "frontend": ["k_dedup_insert", "k_frontend", "k_select_by_mask"],
         "voxel": ["k_voxel_insert", "k_voxel_finalize"],
         "neighbors": ["k_radius_query", "k_grid_insert", "k_grid_assign", "k_grid_scatter", "k_grid_clean"],
         "ransac": ["k_rs_score", "k_rs_final"], "pipeline": ["k_pipeline_counts"]}
buf = (ctypes.c_uint64 * (8 * 2048 * 4))()


def read_all(clear_only=False):
    rows = []
    for tu, kn in names.items():
        getattr(lib, f"apc_debug_trace_{tu}")(buf)
        if clear_only:
            continue
        a = np.frombuffer(buf, dtype=np.uint64).reshape(8, 2048, 4).astype(np.int64)
        for kid, name in enumerate(kn):
            t = a[kid]
            t = t[t[:, 0] > 0]
            if len(t):
                rows.append((name, int(t[:, 0].min()), int(np.median(t[:, 0])), int(t.max()), len(t)))
    return rows


with torch.cuda.stream(ln.stream):
    for rep in range(6):
        ln.d_in.copy_(frames[rep % 4], non_blocking=True)
        ln.stream.synchronize()
        read_all(clear_only=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ln.stream)
        ln.ctx.launch_graph(ln.graph)
        b.record(ln.stream)
        b.synchronize()
        rows = sorted(read_all(), key=lambda r: r[1])
print(f"graph replay: {a.elapsed_time(b) * 1e3:.1f} us (CUDA events)")
t0 = rows[0][1]
prev_end = t0
print(f"{'kernel':20s} {'start':>8s} {'end':>8s} {'dur':>7s} {'gap':>6s}  CTAs   (us, relative to the first kernel)")
for name, s0, s50, e, n in rows:
    print(f"{name:20s} {(s0 - t0) / 1e3:8.2f} {(e - t0) / 1e3:8.2f} {(e - s0) / 1e3:7.2f} {(s0 - prev_end) / 1e3:6.2f}  {n}")
    prev_end = e

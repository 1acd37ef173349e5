"""Which stage stops scaling when several scans run at once?  The C2 pipeline is captured with a growing
set of stages (front end only; + voxel; + radius outliers; + RANSAC ground = the full pipeline) and
replayed over 64 resident frames on 1 / 2 / 4 / 8 lanes.  The table gives microseconds per scan; the
difference between consecutive rows is what a stage costs at that level of concurrency.  A stage that
overlaps perfectly halves with every doubling of the lanes; one bound by a shared resource stays flat.

    python profiles/stage_concurrency.py OUT.json

SHARED_STREAM=1: all lanes share one stream, so the scans run strictly one after the other while their
hash tables alternate between the lanes' contexts - the same serial execution as one lane, minus the
L2 residency of the tables.  N_CONFIGS / LANE_SET restrict the sweep.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autodriver_pointcloud_preprocessor_b200 import _capi, replay  # noqa: E402

F, REPS = 64, int(os.environ.get("REPS", "8"))
block = bench.make_c5_block(0, F, workers=8)
dev = torch.device("cuda", 0)
pool = torch.from_numpy(block.copy()).to(dev)
msg0 = bench.frame_msg(b"")
filter_kw = dict(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                 transforms=[bench.TF], crop=bench.CROP)
S = bench.STAGES
configs = [("frontend", {}),
           ("+ voxel", dict(voxel_size=S["voxel_size"])),
           ("+ radius", dict(voxel_size=S["voxel_size"], radius=S["radius"])),
           ("+ ground (full)", dict(S))]
configs = configs[:int(os.environ.get("N_CONFIGS", "4"))]
lane_set = tuple(int(x) for x in os.environ.get("LANE_SET", "1,2,4,8").split(","))
res = {}
main = torch.cuda.current_stream(dev)
for name, stages in configs:
    row = {}
    for lanes in lane_set:
        pipe = replay.ScanPipeline(msg0.fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, stages, lanes=lanes, device=0)
        if os.environ.get("SHARED_STREAM") == "1":      # L2-residency probe: the lanes' contexts (tables) alternate
            for ln in pipe.lanes[1:]:                   # frame by frame, but everything runs on ONE stream
                ln.stream = pipe.lanes[0].stream
        counts = torch.zeros((F, 8), dtype=torch.int32, device=dev)
        pipe.prepare_resident(pool, None, counts)
        ids = list(range(F))
        pipe.run_resident(ids, main)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        for _ in range(REPS):
            pipe.run_resident(ids, main)
        b.record(main)
        torch.cuda.synchronize()
        row[str(lanes)] = round(a.elapsed_time(b) * 1e3 / (REPS * F), 2)
        row["kernels"] = pipe.kernels_per_scan
        pipe.close()
    res[name] = row
    print(f"{name:18s} us/scan at {lane_set} lanes: {row}", flush=True)
if len(sys.argv) > 1:
    json.dump({"frames": F, "reps": REPS, "us_per_scan": res}, open(sys.argv[1], "w"), indent=1)

"""What does each kernel of the C2 scan cost WITH ALL LANES BUSY (the state `value` is measured in)?

The per-kernel CUDA-event times of bench.py's `kernels` table are eager single-lane launches: they are
latency chains and overlap by 2-3x once eight scans are in flight, unevenly.  Here the pipeline is
captured with only its first k kernels (APC_LAUNCH_BUDGET=k: the library skips every launch after the
k-th of a call) for k = 1 .. all, and replayed over 64 resident frames on 8 lanes and on 1 lane.  The
difference between consecutive rows is the k-th kernel's marginal cost in each regime.

    python profiles/kernel_marginal.py OUT.json
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
main = torch.cuda.current_stream(dev)


def timed(lanes):
    pipe = replay.ScanPipeline(msg0.fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, bench.STAGES, lanes=lanes, device=0)
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
    us = a.elapsed_time(b) * 1e3 / (REPS * F)
    n = pipe.kernels_per_scan
    pipe.close()
    return us, n


# kernel names in launch order (one eager profiled run with the full budget)
os.environ.pop("APC_LAUNCH_BUDGET", None)
pipe = replay.ScanPipeline(msg0.fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, bench.STAGES, lanes=1, device=0)
pipe.lanes[0].d_in.copy_(pool[0])
names = ["k_begin"] + list(pipe.stage_profile().keys())
total = pipe.kernels_per_scan
pipe.close()
print("kernels per scan:", total, names, flush=True)

rows = []
prev8 = prev1 = 0.0
for k in range(1, total + 1):
    os.environ["APC_LAUNCH_BUDGET"] = str(k)
    us8, n8 = timed(8)
    us1, _ = timed(1)
    rows.append({"k": k, "graph_kernels": n8, "us_per_scan_8_lanes": round(us8, 2), "us_per_scan_1_lane": round(us1, 2),
                 "marginal_8_lanes": round(us8 - prev8, 2), "marginal_1_lane": round(us1 - prev1, 2)})
    prev8, prev1 = us8, us1
    print(rows[-1], flush=True)
os.environ.pop("APC_LAUNCH_BUDGET", None)
if len(sys.argv) > 1:
    json.dump({"frames": F, "reps": REPS, "eager_profile_names": names, "rows": rows}, open(sys.argv[1], "w"), indent=1)

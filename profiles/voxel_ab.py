"""Hash vs sort for the voxel grid (north_star (3)): apc_voxel_downsample (open-addressing hash +
finalize) against apc_voxel_downsample_sorted (onesweep radix sort + segmented reduce) at the C2 / C3 /
C4 sizes, CUDA events, warm.  `python profiles/voxel_ab.py OUT.json`; under ncu the same script gives the
per-kernel DRAM bytes and L2 hit rates (profiles/voxel_ab.sh)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodriver_pointcloud_preprocessor_b200 import engine, synth  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(os.environ.get("VOXEL_AB_REPS", "20"))
cases = [("C2", 262_144, 0.1), ("C3", 1_048_576, 0.1), ("C4", 1_500_000, 0.05)]
res = {}
for name, n, vs in cases:
    if name == "C3":
        parts = [synth.lidar_scan(seed=70 + k, nan_frac=0.0)["positions"] for k in range(4)]
        T = synth.sensor_extrinsics(4)
        pos = np.concatenate([(p @ T[k][:3, :3].T + T[k][:3, 3]).astype(np.float32) for k, p in enumerate(parts)])
    else:
        pos = synth.lidar_scan(seed=60, n_beams=128, n_az=2048, n_points=n, nan_frac=0.0)["positions"]
    xyzi = torch.from_numpy(np.concatenate([pos, np.zeros((pos.shape[0], 1), np.float32)], 1)).cuda()
    ctx = engine.Context(max_points=xyzi.shape[0])
    row = {"points": int(xyzi.shape[0]), "voxel_size": vs}
    for how in ("hash", "sort"):
        fn = (lambda: ctx.voxel_downsample(xyzi, vs, want_counts=True)) if how == "hash" else \
             (lambda: ctx.voxel_downsample_sorted(xyzi, vs, want_counts=True))
        for _ in range(3):
            r = fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        ctx.check()
        row[how + "_us"] = round(a.elapsed_time(b) / reps * 1e3, 1)
        row[how + "_voxels"] = int(r[-1].item())
        ctx.profile(True)
        fn()
        row[how + "_kernels_us"] = {k: round(ms * 1e3, 1) for k, (ms, cnt) in ctx.profile_report().items()}
        ctx.profile(False)
    row["winner"] = "hash" if row["hash_us"] <= row["sort_us"] else "sort"
    res[name] = row
    print(name, row, flush=True)
    ctx.close()
if out_path:
    json.dump(res, open(out_path, "w"), indent=1)

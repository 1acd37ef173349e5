"""Where does the end-to-end time go?  Pure H2D, pure D2H, both directions at once, and
process_host at several lane counts (C2 frames, pinned host buffers)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from autodriver_pointcloud_preprocessor_b200 import _capi, replay

F = 64
msgs = bench.make_frames(F, seed0=0)
dev = torch.device("cuda", 0)
h_frames = [torch.frombuffer(bytearray(m.data), dtype=torch.uint8).pin_memory() for m in msgs]
d = [torch.empty_like(h, device=dev) for h in h_frames[:8]]
h_out = [torch.empty(2_700_000, dtype=torch.uint8).pin_memory() for _ in range(8)]
d_out = [torch.empty(2_700_000, dtype=torch.uint8, device=dev) for _ in range(8)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        for f in range(F):
            d[f % 8].copy_(h_frames[f], non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        for f in range(F):
            h_out[f % 8].copy_(d_out[f % 8], non_blocking=True)


def both():
    h2d(); d2h()


ms = timed(h2d); print(f"H2D  64 x 4.19 MB: {ms:.2f} ms  ({64 * 4.194 / ms:.1f} GB/s)")
ms = timed(d2h); print(f"D2H  64 x 2.70 MB: {ms:.2f} ms  ({64 * 2.7 / ms:.1f} GB/s)")
ms = timed(both); print(f"both directions  : {ms:.2f} ms")
filter_kw = dict(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True, transforms=[bench.TF], crop=bench.CROP)
for lanes in (4, 8, 12, 16):
    pipe = replay.ScanPipeline(msgs[0].fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, bench.STAGES, lanes=lanes)
    ms = timed(lambda: pipe.process_host(h_frames, keep_outputs=False))
    print(f"process_host lanes={lanes}: {ms:.2f} ms/step  ({F * bench.N_POINTS / ms / 1e3:.0f} Mpoints/s)")
    pipe.close()

import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from autodriver_pointcloud_preprocessor_b200 import _capi, replay
F = 64
msgs = bench.make_frames(F, seed0=0)
dev = torch.device("cuda", 0)
pool = torch.stack([torch.frombuffer(bytearray(m.data), dtype=torch.uint8).to(dev) for m in msgs])
filter_kw = dict(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True, transforms=[bench.TF], crop=bench.CROP)
for lanes in (1, 2, 4, 8):
    pipe = replay.ScanPipeline(msgs[0].fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, bench.STAGES, lanes=lanes)
    pipe.prepare_resident(pool)
    ids = list(range(F))
    for _ in range(3):
        pipe.run_resident(ids)
    torch.cuda.synchronize()
    cpu, gpu = [], []
    for _ in range(10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        pipe.run_resident(ids)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        cpu.append((t1 - t0) * 1e6 / F)
        gpu.append(e0.elapsed_time(e1) * 1e3 / F)
    print(f"lanes {lanes}: CPU enqueue {np.median(cpu):.1f} us/scan, GPU {np.median(gpu):.1f} us/scan")
    pipe.close()

"""Device time of the other BASELINE.json configurations through the C ABI (captured graph, p50 of
20 replays; inputs resident in HBM), next to the bench workload C2:
  C1  64 x 2048 = 131 072 points: crop + 0.1 m voxel + RANSAC ground removal
  C3  4 sensors x 262 144 points (three byte layouts), per-sensor TF, merged in one launch, 0.1 m voxel
  C4  1.5 M points: 0.05 m voxel + statistical outlier removal k = 20
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from autodriver_pointcloud_preprocessor_b200 import _capi, engine, synth


def dev_bytes(m):
    return torch.frombuffer(bytearray(m.data), dtype=torch.uint8).cuda()


def time_graph(ctx, descs, pcfg, n_total, label):
    out = torch.zeros((n_total, 4), device="cuda")
    counts = torch.zeros(8, dtype=torch.int32, device="cuda")
    plane = torch.zeros(8, dtype=torch.float64, device="cuda")
    ctx.profile(True)
    ctx.pipeline_run(descs, pcfg, out, counts, plane)
    rep = ctx.profile_report()
    ctx.profile(False)
    g = ctx.capture_pipeline(descs, pcfg, out, counts, plane)
    lat = []
    for _ in range(23):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.launch_graph(g); b.record(); b.synchronize()
        lat.append(a.elapsed_time(b) * 1e3)
    ctx.check()
    c = counts.cpu().numpy()
    p50 = float(np.median(lat[3:]))
    print(f"{label}: p50 {p50:8.1f} us  ({n_total / p50:7.1f} Mpoints/s one scan at a time)  "
          f"in {c[_capi.CNT_INPUT]} -> filtered {c[_capi.CNT_FILTERED]} -> voxels {c[_capi.CNT_VOXELS]} -> out {c[_capi.CNT_OUTPUT]}")
    for k, (ms, n) in sorted(rep.items(), key=lambda kv: -kv[1][0])[:6]:
        print(f"      {k:20s} {ms / n * 1e3:8.1f} us x{n}")


ctx = engine.Context(max_points=1_600_000)
crop = dict(min=[-60.0, -60.0, -20.0], max=[60.0, 60.0, 20.0], invert=False, mode=2)
ground = dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=7)

m = synth.pack_cloud(synth.lidar_scan(seed=1, n_beams=64, n_az=2048), "xyzi16")
fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True, crop=crop)
buf = dev_bytes(m)
time_graph(ctx, [engine.make_cloud_desc(m.fields, m.point_step, m.width, buf)],
           engine.make_pipeline_cfg(fcfg, voxel_size=0.1, ground=ground), m.width, "C1 131k crop+voxel+ground     ")

T = synth.sensor_extrinsics(4)
msgs = [synth.pack_cloud(synth.lidar_scan(seed=90 + s, nan_frac=0.0), lay)
        for s, lay in enumerate(["xyzi16", "xyzirt22", "ouster48", "xyzi16"])]
bufs = [dev_bytes(x) for x in msgs]
descs = [engine.make_cloud_desc(x.fields, x.point_step, x.width, b, transform=T[s]) for s, (x, b) in enumerate(zip(msgs, bufs))]
time_graph(ctx, descs, engine.make_pipeline_cfg(engine.make_filter_cfg(), voxel_size=0.1), 4 * 262144,
           "C3 4x262k concat+TF+voxel     ")

scan = synth.lidar_scan(seed=5, n_points=1_500_000, nan_frac=0.0, dup_frac=0.0)
m4 = synth.pack_cloud(scan, "xyzi16")
buf4 = dev_bytes(m4)
time_graph(ctx, [engine.make_cloud_desc(m4.fields, m4.point_step, m4.width, buf4)],
           engine.make_pipeline_cfg(engine.make_filter_cfg(), voxel_size=0.05, statistical=dict(nb_neighbors=20, std_ratio=2.0)),
           m4.width, "C4 1.5M voxel 0.05+statistical")

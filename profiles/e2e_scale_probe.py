"""Where does the end-to-end path stop scaling?  Run under torchrun at N = 1/2/4/8 ranks on one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 profiles/e2e_scale_probe.py OUT.json

Every rank times, between barriers, on the C2 workload of bench.py (64 frames of 4.19 MB per step):
  h2d_big / d2h_big   one pinned 268 MB upload / one 175 MB download per step (PCIe + host DRAM ceiling)
  duplex_big          both at once on two streams
  duplex_frames       the same bytes as 64 + 64 per-frame copies (what per-frame pipelining costs)
  host_memcpy         pinned -> pageable numpy copy of the 268 MB arena (host DRAM bandwidth per rank)
  kernels             device-resident graphs only (ScanPipeline.run_resident)
  e2e                 ScanPipeline.process_host (pinned bytes in, host rows out)
  e2e_batched         ScanPipeline.process_host_batched when the build has it
and rank 0 writes per-rank milliseconds plus the aggregate GB/s of every mode as one JSON file.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from autodriver_pointcloud_preprocessor_b200 import _capi, replay  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
out_path = sys.argv[1] if len(sys.argv) > 1 else None
if os.environ.get("PROBE_BIND", "1") == "1":
    bench.bind_to_gpu_numa_node(local_rank)
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

F = 64
OUT_ROWS = 170_600                                   # mean surviving rows per C2 frame
base = bench.make_frames(8, seed0=1000 * rank)        # 8 distinct scans, replayed 8 times: keeps the probe short
msgs = [base[f % 8] for f in range(F)]
h_in = torch.empty((F, bench.N_POINTS * bench.POINT_STEP), dtype=torch.uint8).pin_memory()
for f, m in enumerate(msgs):
    h_in[f] = torch.frombuffer(bytearray(m.data), dtype=torch.uint8)
h_frames = [h_in[f] for f in range(F)]
d_in = torch.empty_like(h_in, device=dev)
h_out = torch.empty((F, OUT_ROWS, 4), dtype=torch.float32).pin_memory()
d_out = torch.zeros((F, OUT_ROWS, 4), dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
IN_B, OUT_B = h_in.numel(), h_out.numel() * 4


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=4):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        return [round(float(x.item()), 3) for x in allt]
    return [round(ms, 3)]


def h2d_big():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h_big():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def duplex_big():
    h2d_big()
    d2h_big()


def duplex_frames():
    with torch.cuda.stream(s1):
        for f in range(F):
            d_in[f].copy_(h_in[f], non_blocking=True)
    with torch.cuda.stream(s2):
        for f in range(F):
            h_out[f].copy_(d_out[f], non_blocking=True)


scratch = np.empty(IN_B, dtype=np.uint8)


def host_memcpy():
    np.copyto(scratch, h_in.numpy().reshape(-1))


res = {"world": world, "cpus_rank0": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(),
       "bytes_in_per_step": IN_B, "bytes_out_per_step": OUT_B, "modes": {}}


def record(name, per_rank_ms, nbytes):
    worst = max(per_rank_ms)
    res["modes"][name] = {"ms_per_rank": per_rank_ms, "ms_max": worst,
                          "aggregate_GBps": round(world * nbytes / worst / 1e6, 1) if nbytes else None,
                          "Mpoints_per_s": round(world * F * bench.N_POINTS / worst / 1e3, 1)}
    if rank == 0:
        print(name, res["modes"][name], file=sys.stderr, flush=True)


record("h2d_big", timed(h2d_big), IN_B)
record("d2h_big", timed(d2h_big), OUT_B)
record("duplex_big", timed(duplex_big), IN_B + OUT_B)
record("duplex_frames", timed(duplex_frames), IN_B + OUT_B)
record("host_memcpy", timed(host_memcpy), 2 * IN_B)

filter_kw = dict(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                 transforms=[bench.TF], crop=bench.CROP)
pipe = replay.ScanPipeline(msgs[0].fields, bench.POINT_STEP, bench.N_POINTS, filter_kw, bench.STAGES, lanes=8,
                           device=local_rank)
pool = d_in
pool.copy_(h_in)
counts_arena = torch.zeros((F, 8), dtype=torch.int32, device=dev)
pipe.prepare_resident(pool, None, counts_arena)
main = torch.cuda.current_stream(dev)
record("kernels", timed(lambda: pipe.run_resident(list(range(F)), main)), 0)
record("e2e", timed(lambda: pipe.process_host(h_frames, keep_outputs=False)), IN_B + OUT_B)
if hasattr(pipe, "process_host_batched"):
    record("e2e_batched", timed(lambda: pipe.process_host_batched(h_in, keep_outputs=False)), IN_B + OUT_B)
    record("e2e_batched_copy", timed(lambda: pipe.process_host_batched(h_in, keep_outputs=True)), IN_B + OUT_B)

if rank == 0 and out_path:
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

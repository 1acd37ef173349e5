#!/usr/bin/env python
"""Benchmark of the per-scan preprocessing hot path (BASELINE.json: Mpoints/s of the full
preprocess pipeline; p50 per-scan latency @ 262k points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): Ouster OS1-128 shape, 128 x 2048 = 262144-point
synthetic spinning-LiDAR scans, full pipeline = read_points NaN skip + duplicate removal +
non-finite filter + TF transform + ROI crop + 0.1 m voxel downsample + radius outlier
removal (5 pts / 0.5 m) + RANSAC ground removal (0.2 m, n=5, 100 iterations, p=0.99).
A "step" is one pass over a batch of FRAMES distinct scans (64 x 4.19 MB = 268 MB of input,
larger than the 126 MB L2, so every step streams its input from HBM).  With N > 1 (torchrun)
every rank processes its own FRAMES scans (frame-parallel, weak scaling) and the per-GPU
outputs are all-gathered over NCCL once per step, overlapped with the next step's compute.

One JSON line is printed by rank 0; see the keys at the bottom of ``main``.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_BEAMS, N_AZ = 128, 2048
N_POINTS = N_BEAMS * N_AZ
LAYOUT = "xyzi16"
POINT_STEP = 16
TF = np.array([[0.9986295, -0.0523360, 0.0, 1.5], [0.0523360, 0.9986295, 0.0, -0.25],
               [0.0, 0.0, 1.0, 0.2], [0.0, 0.0, 0.0, 1.0]])
CROP = dict(min=[-60.0, -60.0, -20.0], max=[60.0, 60.0, 20.0], invert=False, mode=2)
STAGES = dict(voxel_size=0.1, radius=dict(nb_points=5, radius=0.5),
              ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=100, probability=0.99, seed=7))
WORKLOAD = ("C2: Ouster OS1-128 shape 128x2048=262144 pts/scan, xyzi16 layout; NaN skip + dedup + non-finite + TF + "
            "crop(+-60,+-60,+-20) + voxel 0.1 m + radius outliers(5, 0.5 m) + RANSAC ground(0.2, 5, 100, 0.99)")


def make_frames(n_frames: int, seed0: int):
    """``n_frames`` distinct scans: n/4 ray-cast scenes x 4 yaw-rotated copies (cheap variety)."""
    from autodriver_pointcloud_preprocessor_b200 import synth
    n_base = max(1, (n_frames + 3) // 4)
    msgs = []
    for b in range(n_base):
        scan = synth.lidar_scan(seed=seed0 + b, n_beams=N_BEAMS, n_az=N_AZ)
        for r in range(4):
            if len(msgs) == n_frames:
                break
            yaw = np.deg2rad(7.0 * r)
            R = np.array([[np.cos(yaw), -np.sin(yaw), 0.0], [np.sin(yaw), np.cos(yaw), 0.0], [0.0, 0.0, 1.0]],
                         dtype=np.float32)
            sc = dict(scan)
            sc["positions"] = (scan["positions"] @ R.T).astype(np.float32)
            msgs.append(synth.pack_cloud(sc, LAYOUT))
    return msgs


def oracle_config():
    from oracle import pipeline as opipe
    cfg = opipe.default_config()
    cfg.update(transforms=[TF], crop=CROP, voxel_size=STAGES["voxel_size"], radius=STAGES["radius"],
               ground=STAGES["ground"])
    return cfg


def cpu_pipeline_seconds(msgs):
    """The reference's CPU path (numpy verbatim expressions + Open3D semantics restated with
    numpy / scipy cKDTree on all host cores) on the given scans."""
    from oracle import pipeline as opipe
    cfg = oracle_config()
    t0 = time.perf_counter()
    n_out = 0
    for m in msgs:
        n_out += opipe.preprocess(m, cfg)["positions"].shape[0]
    return time.perf_counter() - t0, n_out


class ClockSampler(threading.Thread):
    """nvidia-smi SM clock / throttle-reason samples during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        # one long-lived nvidia-smi in loop mode (20 ms period): many samples even in a short region
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append((time.perf_counter(), parts))
        except Exception:
            pass

    def stop(self, t_begin=0.0, t_end=float('inf')):
        """Summary of the samples taken inside [t_begin, t_end] (perf_counter clock)."""
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=6)
        inside = [p for (t, p) in self.samples if t_begin <= t <= t_end]
        self.samples = inside if inside else [p for (_, p) in self.samples[-3:]]
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


_REAL_STDOUT = None
_ORIG_AFFINITY = None


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the CPUs NVML
    reports as local to its GPU: with 8 ranks on a two-socket box the end-to-end path otherwise moves
    half of its PCIe traffic across the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        global _ORIG_AFFINITY
        _ORIG_AFFINITY = set(os.sched_getaffinity(0))
        cpus &= _ORIG_AFFINITY
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:                                                 # best effort: NVML / affinity not available
        pass


def emit(line: dict):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def algorithmic_bytes(kernel: str, c) -> float | None:
    """Algorithmic HBM bytes of one launch (DESIGN.md section 'Kernels'), from the measured
    per-frame counts: N input, M filtered, V voxels, P after radius, ps = point_step."""
    N, M, V = float(c["N"]), float(c["M"]), float(c["V"])
    P = float(c["P_radius_in"])
    table = {
        "k_dedup_insert": N * POINT_STEP + 4 * N + 16 * N,          # read records, write p2slot, touch one 16 B slot
        "k_frontend": N * POINT_STEP + 4 * N + 16 * M,              # read records + p2slot, write survivors
        "k_voxel_insert": 16 * M + 4 * M + 48 * M,                  # read points, write p2slot, RMW one 48 B slot
        "k_voxel_finalize": 4 * M + 4 * M + 48 * V + 16 * V,        # p2slot + first[], read/clean slots, write centroids
        "k_grid_insert": 16 * P + 8 * P + 12 * P,                   # points, slot+rank, key/fill RMW
        "k_grid_assign": 8 * P + 8 * P,
        "k_grid_scatter": 16 * P + 12 * P + 16 * P,
        "k_radius_query": 16 * P + 1 * P,                           # every neighbour read is an on-chip re-read
        "k_grid_clean": 8 * P,
        "k_select_by_mask": 16 * P + 1 * P + 16 * P,
        "k_radius_select": 16 * P + 1 * P + 8 * P + 16 * float(c["P_ground_in"]),   # points + mask + slot/rank, write survivors
        "k_rs_score": 16 * float(c["P_ground_in"]),                 # one pass over the points for all hypotheses
        "k_rs_final": 16 * float(c["P_ground_in"]) + float(c["P_ground_in"]),
    }
    return table.get(kernel)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: numpy
    verbatim + Open3D restated; Open3D itself is not installable) on the host cores."""
    if rank != 0:
        return
    import torch
    frames_per_step = 2
    msgs = make_frames(frames_per_step, seed0=0)
    for _ in range(args.warmup):
        cpu_pipeline_seconds(msgs[:1])
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_pipeline_seconds(msgs)
        t += dt
    ms = t / args.steps * 1e3
    value = frames_per_step * N_POINTS / (ms * 1e-3) / 1e6
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "Mpoints/s full preprocess pipeline", "value": round(value, 4),
        "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": frames_per_step},
        "cpu_baseline": {"value": round(value, 4), "unit": "Mpoints/s", "cores": cores, "kind": "port",
                         "sample": f"{frames_per_step} scans of the workload per step; numpy single-threaded + "
                                   f"scipy cKDTree workers=-1 ({cores} cores, torch threads {torch.get_num_threads()})"},
        "e2e": {"value": round(value, 4), "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="distinct scans per step per GPU")
    ap.add_argument("--lanes", type=int, default=8, help="concurrent stream lanes per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # exactly ONE line on stdout: libraries that print to fd 1 (NCCL's version banner when the
    # box sets NCCL_DEBUG) are sent to stderr; the JSON line goes to the real stdout at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from autodriver_pointcloud_preprocessor_b200 import _capi, replay

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    F = args.frames
    msgs = make_frames(F, seed0=1000 * rank)
    h_frames = [torch.frombuffer(bytearray(m.data), dtype=torch.uint8).pin_memory() for m in msgs]
    pool = torch.stack([h.to(dev, non_blocking=True) for h in h_frames])
    filter_kw = dict(skip_nans=True, dedup_mode=_capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                     transforms=[TF], crop=CROP)
    pipe = replay.ScanPipeline(msgs[0].fields, POINT_STEP, N_POINTS, filter_kw, STAGES, lanes=args.lanes,
                               device=local_rank)
    arena = counts_arena = None
    if os.environ.get("APC_FORCE_ARENA"):                          # diagnostic: one output buffer per frame
        arena = torch.zeros((F, N_POINTS, 4), dtype=torch.float32, device=dev)
    counts_arena = torch.zeros((F, 8), dtype=torch.int32, device=dev)
    pipe.prepare_resident(pool, arena, counts_arena)
    frame_ids = list(range(F))
    main_stream = torch.cuda.current_stream(dev)

    # multi-GPU: every lane stages a frame's output rows into the send slab of the step's parity right
    # behind the frame's graph; the all-gather of step k's slab runs on a side stream underneath step
    # k+1, which fills the other parity
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    pad_rows, send2 = 0, None
    send_counts = torch.zeros(F, dtype=torch.int32, device=dev) if world > 1 else None
    all_counts = torch.zeros((world, F), dtype=torch.int32, device=dev) if world > 1 else None
    peer_gather, gather_how = None, "NCCL all-gather"
    gather_done = [None, None]
    step_no = 0
    no_gather = os.environ.get("APC_GATHER") == "none"               # diagnostic: attribute the gather's cost

    def step():
        nonlocal step_no
        if world == 1 or no_gather or send2 is None:
            pipe.run_resident(frame_ids, main_stream)
            return
        parity = step_no & 1
        step_no += 1
        if gather_done[parity] is not None:
            main_stream.wait_event(gather_done[parity])          # the gather two steps back has sent this slab
        slab = peer_gather.slot(parity) if peer_gather is not None else send2[parity]
        pipe.run_resident(frame_ids, main_stream, stage_to=slab)
        send_counts.copy_(counts_arena[:, _capi.CNT_OUTPUT])
        done = torch.cuda.Event()
        done.record(main_stream)
        comm_stream.wait_event(done)
        with torch.cuda.stream(comm_stream):
            if peer_gather is not None:
                peer_gather.gather(parity)
                dist.all_gather_into_tensor(all_counts, send_counts)
            else:
                replay.gather_outputs(slab, send_counts)
            gather_done[parity] = torch.cuda.Event()
            gather_done[parity].record(comm_stream)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- warm-up (also sizes the gather slabs) -----------------------------------------------------
    pipe.run_resident(frame_ids, main_stream)
    torch.cuda.synchronize(dev)
    pipe.check()
    counts0 = counts_arena.cpu().numpy()
    assert (counts0[:, _capi.CNT_STATUS] == 0).all()
    if world > 1:
        mx = torch.tensor([int(counts0[:, _capi.CNT_OUTPUT].max())], device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        pad_rows = min(N_POINTS, (int(mx.item()) * 9 // 8 + 1023) // 1024 * 1024)
        if os.environ.get("APC_GATHER", "peer") == "peer":
            try:
                peer_gather = replay.PeerGather((F, pad_rows, 4), torch.float32, dev)
                gather_how = "peer-to-peer copy-engine writes over NVLink into symmetric buffers + NCCL counts"
            except Exception as e:                                               # no symmetric memory / peer access
                print(f"PeerGather unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
        send2 = True if peer_gather is not None else torch.zeros((2, F, pad_rows, 4), dtype=torch.float32, device=dev)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)

    # ---- timed region: K steps, device events, barrier + synchronize on both sides ------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_stream)
    for _ in range(args.steps):
        step()
    if world > 1:
        main_stream.wait_stream(comm_stream)
    e1.record(main_stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, time.perf_counter())
    if world > 1 and pad_rows:                    # the exchanged slabs hold every frame's output in full
        assert int(counts_arena[:, _capi.CNT_OUTPUT].max().item()) <= pad_rows, "gather slab too small"
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * F * N_POINTS / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API: pinned host bytes in, host points out --------------------
    pipe.process_host(h_frames[:8], keep_outputs=False)                      # warm-up
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    d2h = 0
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        _, _, b = pipe.process_host(h_frames, keep_outputs=False)
        d2h += b
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * F * N_POINTS / (e2e_ms * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0 only: latency, per-kernel roofline, CPU baseline ---------------------------------------
    ln = pipe.lanes[0]
    lat = []
    with torch.cuda.stream(ln.stream):
        for rep in range(3):
            for f in range(0, F, len(pipe.lanes)):                             # lane 0's own frames
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(ln.stream)
                ln.ctx.launch_graph(ln.resident_graphs[f])
                b.record(ln.stream)
                b.synchronize()
                lat.append(a.elapsed_time(b))
    lat_e2e = []
    for f in range(min(F, 32)):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        pipe.process_host([h_frames[f]], keep_outputs=True)
        lat_e2e.append((time.perf_counter() - t0) * 1e3)

    # per-kernel CUDA-event timings over several frames (eager launches on lane 0's stream)
    prof = {}
    n_prof = 8
    for f in range(n_prof):
        with torch.cuda.stream(ln.stream):
            ln.d_in.copy_(pool[f], non_blocking=True)
        for k, (ms, cnt) in pipe.stage_profile().items():
            p = prof.setdefault(k, [0.0, 0])
            p[0] += ms
            p[1] += cnt
    c = counts0.astype(np.float64).mean(axis=0)
    cnt = {"N": c[_capi.CNT_INPUT], "M": c[_capi.CNT_FILTERED], "V": c[_capi.CNT_VOXELS],
           "P_radius_in": c[_capi.CNT_VOXELS], "P_ground_in": c[_capi.CNT_AFTER_RADIUS], "out": c[_capi.CNT_OUTPUT]}
    tot_ms = sum(v[0] for v in prof.values())
    kernels = []
    for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        per_launch_ms = ms / max(n, 1)
        ab = algorithmic_bytes(k, cnt)
        kernels.append({"kernel": k, "share": round(ms / tot_ms, 4), "us_per_launch": round(per_launch_ms * 1e3, 2),
                        "algorithmic_MB": round(ab / 1e6, 3) if ab else None,
                        "GBps": round(ab / (per_launch_ms * 1e-3) / 1e9, 1) if ab else None})
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # measured DRAM bytes per launch (one `ncu --set full` capture of an eager scan, summarised by
    # profiles/summarise_ncu.py into the newest profiles/*_traffic.json that is committed)
    traffic, traffic_src = {}, None
    tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
    if tfiles:
        tj = json.load(open(os.path.join(ROOT, "profiles", tfiles[-1])))
        traffic, traffic_src = tj.get("dram_bytes_per_launch", {}), f"profiles/{tfiles[-1]}: {tj.get('source', '')}"
    for k in kernels:
        k["dram_traffic_MB"] = round(traffic[k["kernel"]] / 1e6, 3) if k["kernel"] in traffic else None
    dom = kernels[0]
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                "frac": round(dom["GBps"] / peak, 5) if dom["GBps"] else None,
                "traffic": traffic.get(dom["kernel"]), "traffic_source": traffic_src, "peak_source": peak_src,
                "share_of_step": dom["share"], "us_per_launch": dom["us_per_launch"],
                "note": "algorithmic bytes / CUDA-event duration of the dominant kernel, eager launch on its own "
                        "stream, mean over 8 frames; per-kernel table under 'kernels'"}

    cpu = None
    if not args.no_cpu_baseline:
        if _ORIG_AFFINITY:                                                      # the CPU baseline gets every host core back
            os.sched_setaffinity(0, _ORIG_AFFINITY)
        n_cpu = 3
        cpu_pipeline_seconds(msgs[:1])                                           # warm-up (thread pools, imports)
        secs, _ = cpu_pipeline_seconds(msgs[:n_cpu])
        cpu = {"value": round(n_cpu * N_POINTS / secs / 1e6, 4), "unit": "Mpoints/s", "cores": len(os.sched_getaffinity(0)),
               "kind": "port", "ms_per_scan": round(secs / n_cpu * 1e3, 1),
               "sample": f"{n_cpu} scans of the same workload through oracle/pipeline.py (numpy + scipy cKDTree "
                         f"workers=-1) after 1 warm-up scan"}

    line = {
        "metric": "Mpoints/s full preprocess pipeline", "value": round(value, 2), "unit": "Mpoints/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "lanes": args.lanes,
                   "input_bytes_per_step_per_gpu": F * N_POINTS * POINT_STEP,
                   "l2": "inputs larger than L2 (268 MB of distinct scans per step vs 126 MB L2)",
                   "multi_gpu": f"frame-parallel, no per-scan collective; per-step all-gather of outputs ({gather_how}) "
                                "overlapped with the next step" if world > 1 else "single GPU"},
        "p50_latency_ms": round(float(np.median(lat)), 4), "p99_latency_ms": round(float(np.percentile(lat, 99)), 4),
        "p50_latency_e2e_ms": round(float(np.median(lat_e2e)), 4),
        "points_per_scan": {k: round(float(v), 1) for k, v in cnt.items()},
        "e2e": {"value": round(e2e_value, 2), "unit": "Mpoints/s", "h2d_bytes_per_step": F * N_POINTS * POINT_STEP,
                "d2h_bytes_per_step": int(d2h / e2e_steps), "ms_per_step": round(e2e_ms, 3)},
        "gpu_launches": int(pipe.kernels_per_scan * F * args.steps),
        "kernels_per_scan": int(pipe.kernels_per_scan),
        "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Streaming / batched replay engine: the public per-GPU API for processing many scans.

``ScanPipeline`` owns ``lanes`` independent lanes (CUDA stream + apc context + static device
buffers + one captured CUDA graph of the whole preprocess() chain).  Frames are dealt to the
lanes round-robin so that the copy engines and the SMs of one B200 stay busy: while one
lane's kernels run, another lane's H2D / D2H copies and small latency-bound kernels overlap.

Scans are independent units (the reference keeps no cross-frame state besides the cached
static TF, pp.py:706), so multi-GPU is frame-parallel with no collective on the per-scan
path; ``shard_frames`` deals contiguous blocks of frames to ranks.  The batched PCAP-replay
configuration ends with every GPU holding every GPU's output: ``PeerSlabs`` fuses that exchange
into the pipeline's final kernel (peer stores over NVLink while the kernels run);
``gather_outputs`` is the plain all-gather (NCCL on the GPU box when symmetric memory is not
available, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi, engine


def shard_frames(n_frames: int, world_size: int, rank: int) -> range:
    """Contiguous block of frame indices owned by ``rank``."""
    base, rem = divmod(n_frames, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_outputs(send: torch.Tensor, counts: torch.Tensor, group=None):
    """All-gather per-rank outputs.

    ``send``   [F, pad_rows, 4] float32 - each frame's surviving points, zero padded;
    ``counts`` [F] int32 - valid rows per frame.
    Returns ``(recv [G, F, pad_rows, 4], all_counts [G, F])``.  Works on CUDA tensors with
    the NCCL backend and on CPU tensors with gloo.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    all_counts = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
    if send.is_cuda:
        dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=group)
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    else:
        dist.all_gather(list(all_counts.unbind(0)), counts.contiguous(), group=group)
        dist.all_gather(list(recv.unbind(0)), send.contiguous(), group=group)
    return recv, all_counts


class PeerSlabs:
    """Output slabs of the exchange that is FUSED into the pipeline's final stage (``apc_out_mirror``).

    Every rank owns ``[parity][source rank][frame][rows][4]`` float32 plus ``[...][8]`` int32 counter
    slabs in symmetric memory.  A frame's graph writes its surviving rows into this rank's own block
    of the LOCAL slab and, with the same store loop, into the same block of every PEER's slab over
    NVLink (peer-mapped pointers, or one NVLS multicast address when the fabric offers it): no
    gather pass, no staging copy, no padding on the wire - only the rows that exist travel, while
    the kernels run.  ``barrier()`` (signal pads, on the current stream) separates "all ranks have
    finished writing this parity" from its readers / its next reuse.

    Needs ``torch.distributed._symmetric_memory`` and peer access between the GPUs of the node."""

    def __init__(self, n_frames: int, rows: int, device, group=None, parities: int = 2, multicast: bool | None = None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_frames, self.rows, self.parities = int(n_frames), int(rows), int(parities)
        shape = (self.parities, self.world, self.n_frames, self.rows, 4)
        cshape = (self.parities, self.world, self.n_frames, 8)
        self.buf = symm.empty(shape, dtype=torch.float32, device=device)
        self.cnt = symm.empty(cshape, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.cnt.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        self.hdl_c = symm.rendezvous(self.cnt, group)
        self.peer_buf = [self.hdl.get_buffer(p, shape, torch.float32) for p in range(self.world)]
        self.peer_cnt = [self.hdl_c.get_buffer(p, cshape, torch.int32) for p in range(self.world)]
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.multicast_base = mc if (multicast is None or multicast) else 0
        if multicast and not mc:
            raise RuntimeError("NVLS multicast requested but the symmetric-memory handle has no multicast mapping")

    def local_out(self, parity: int, f: int):
        """(rows [rows, 4], counters [8]) of frame ``f`` in this rank's own block of the local slab."""
        return self.buf[parity][self.rank][f], self.cnt[parity][self.rank][f]

    def mirror(self, parity: int, f: int):
        """``apc_out_mirror`` addressing the same block in every peer's slab."""
        peers = [(self.rank + k) % self.world for k in range(1, self.world)]     # rank+1, rank+2, ...: no hot spot
        cnts = [self.peer_cnt[p][parity][self.rank][f].data_ptr() for p in peers]
        if self.multicast_base:
            # the multicast mapping covers every rank's buffer at the same offset (this rank's included:
            # its own copy is then written twice, which is harmless)
            off = self.buf[parity][self.rank][f].data_ptr() - self.buf.data_ptr()
            return engine.make_out_mirror([self.multicast_base + off], cnts, multicast=True)
        return engine.make_out_mirror([self.peer_buf[p][parity][self.rank][f].data_ptr() for p in peers], cnts)

    def barrier(self, channel: int = 0):
        self.hdl.barrier(channel=channel)

    def frame(self, parity: int, src_rank: int, f: int):
        """Received (or own) rows and counters of frame ``f`` of ``src_rank``."""
        return self.buf[parity][src_rank][f], self.cnt[parity][src_rank][f]


class _Lane:
    __slots__ = ("stream", "ctx", "d_in", "d_out", "d_counts", "d_plane", "h_counts", "h_out", "graph",
                 "resident_graphs", "pending", "event", "copy_event", "desc")


class ScanPipeline:
    """preprocess() for a stream of same-layout scans on one GPU.

    Parameters mirror the node's (``pointcloud_preprocessor.PointcloudPreprocessorNode``):
    ``fields`` / ``point_step`` / ``n_points`` describe the PointCloud2 layout, ``filter_kw``
    is passed to :func:`engine.make_filter_cfg`, ``stages`` to :func:`engine.make_pipeline_cfg`.
    ``low_latency``: for ``lanes=1`` (a node processing one scan at a time) - the kernels of a scan
    are launched as programmatic dependents (``apc_ctx_set_low_latency``).
    """

    def __init__(self, fields, point_step: int, n_points: int, filter_kw: dict, stages: dict,
                 lanes: int = 4, device: int | None = None, low_latency: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required: this package has no CPU path")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        self.fields, self.point_step, self.n_points = list(fields), int(point_step), int(n_points)
        self.frame_bytes = self.point_step * self.n_points
        self.pcfg = engine.make_pipeline_cfg(engine.make_filter_cfg(**filter_kw), **stages)
        self.lanes = []
        for _ in range(lanes):
            ln = _Lane()
            ln.stream = torch.cuda.Stream(device=self.device)
            ln.ctx = engine.Context(max_points=self.n_points, device=self.device_index)
            if low_latency:                     # one scan in flight per lane: programmatic dependent launches
                ln.ctx.set_low_latency(True)
            ln.d_in = torch.zeros(self.frame_bytes, dtype=torch.uint8, device=self.device)
            ln.d_out = torch.zeros((self.n_points, 4), dtype=torch.float32, device=self.device)
            ln.d_counts = torch.zeros(8, dtype=torch.int32, device=self.device)
            ln.d_plane = torch.zeros(8, dtype=torch.float64, device=self.device)
            ln.h_counts = torch.zeros(8, dtype=torch.int32).pin_memory()
            ln.h_out = torch.zeros((self.n_points, 4), dtype=torch.float32).pin_memory()
            ln.desc = engine.make_cloud_desc(self.fields, self.point_step, self.n_points, ln.d_in)
            ln.graph = ln.ctx.capture_pipeline([ln.desc], self.pcfg, ln.d_out, ln.d_counts, ln.d_plane)
            ln.resident_graphs = {}
            ln.pending = None
            ln.event = torch.cuda.Event()
            ln.copy_event = torch.cuda.Event()
            self.lanes.append(ln)
        torch.cuda.synchronize(self.device)
        #: kernels launched per scan by one graph replay (our own kernels; counted at capture)
        self.kernels_per_scan = self._count_kernels()

    def _count_kernels(self) -> int:
        n = _capi.lib.apc_graph_kernel_count(self.lanes[0].graph)
        if n < 0:
            raise RuntimeError("apc_graph_kernel_count failed")
        return n

    def stage_profile(self) -> dict:
        """One eager (non-graph) run on lane 0 with every kernel bracketed by CUDA events:
        ``{kernel: (ms, launches)}`` for the frame currently in the lane's input buffer."""
        ln = self.lanes[0]
        with torch.cuda.stream(ln.stream):
            ln.ctx.profile(True)
            ln.ctx.pipeline_run([ln.desc], self.pcfg, ln.d_out, ln.d_counts, ln.d_plane)
            rep = ln.ctx.profile_report()
            ln.ctx.profile(False)
        return rep

    def close(self):
        for ln in self.lanes:
            ln.ctx.close()
        self.lanes = []

    # ---- end to end: host bytes in, host points out ----------------------------------------------
    def process_host(self, frames, keep_outputs: bool = True):
        """``frames``: pinned uint8 host tensors (one PointCloud2 ``data`` buffer each).
        Returns ``(outputs, counts, d2h_bytes)``: per frame a float32 [n_out, 4] numpy array
        (x, y, z, intensity) and the 8 pipeline counters.  ``keep_outputs``: ``True`` = every output is an
        owned copy; ``"view"`` = a view of the lane's pinned output buffer, valid until that lane's next
        frame (at most ``lanes`` frames per call: what a node publishing one scan at a time needs - the
        message constructor copies the records anyway, pp.py:769); ``False`` = outputs stay in the
        pinned buffers and ``None`` is returned for them.

        Software-pipelined per lane so that the host thread never waits for a copy it has just
        issued: visiting a lane (1) harvests the payload copy issued two visits ago, (2) reads the
        counters of the previous frame (its graph has had a whole round of the other lanes to
        finish) and issues the device->host copy of exactly the surviving rows, (3) issues the
        next frame's upload + graph + counters read-back.  H2D, kernels and D2H of different
        frames overlap across lanes."""
        if keep_outputs == "view" and len(frames) > len(self.lanes):
            raise ValueError("keep_outputs='view': at most one frame per lane per call")
        results = [None] * len(frames)
        counts = np.zeros((len(frames), 8), dtype=np.int32)
        self._d2h_bytes = 0
        S = len(self.lanes)
        for ln in self.lanes:
            ln.pending = [None, None]                   # [frame whose graph is in flight, (frame, n) whose payload copy is]
        for f, h_in in enumerate(frames):
            ln = self.lanes[f % S]
            self._harvest(ln, results, keep_outputs)
            self._issue_payload_copy(ln, counts)
            with torch.cuda.stream(ln.stream):
                ln.d_in.copy_(h_in, non_blocking=True)
                ln.ctx.launch_graph(ln.graph)
                ln.h_counts.copy_(ln.d_counts, non_blocking=True)
                ln.event.record(ln.stream)
            ln.pending[0] = f
        for ln in self.lanes:
            self._harvest(ln, results, keep_outputs)
            self._issue_payload_copy(ln, counts)
        for ln in self.lanes:
            self._harvest(ln, results, keep_outputs)
        return results, counts, self._d2h_bytes

    def _issue_payload_copy(self, ln, counts):
        f = ln.pending[0]
        if f is None:
            return
        ln.event.synchronize()
        counts[f] = ln.h_counts.numpy()
        if counts[f, _capi.CNT_STATUS] != 0:
            ln.ctx.check()
        n = int(counts[f, _capi.CNT_OUTPUT])
        with torch.cuda.stream(ln.stream):
            ln.h_out[:n].copy_(ln.d_out[:n], non_blocking=True)
            ln.copy_event.record(ln.stream)
        ln.pending = [None, (f, n)]
        self._d2h_bytes += n * 16 + 32

    def _harvest(self, ln, results, keep_outputs):
        if ln.pending[1] is None:
            return
        f, n = ln.pending[1]
        ln.copy_event.synchronize()
        if keep_outputs == "view":
            results[f] = ln.h_out[:n].numpy()
        elif keep_outputs:
            results[f] = ln.h_out[:n].numpy().copy()
        ln.pending[1] = None

    # ---- device resident: inputs already in HBM -----------------------------------------------------
    def prepare_resident(self, pool: torch.Tensor, arena: torch.Tensor | None = None,
                         counts_arena: torch.Tensor | None = None, slabs: "PeerSlabs | None" = None):
        """Capture one graph per pool frame (``pool`` [F, frame_bytes] uint8 on the device) so
        replaying reads each frame in place.  With ``arena`` [F, n_points, 4] / ``counts_arena``
        [F, 8] every frame writes its own output slot.  With ``slabs`` (multi-GPU) one graph per
        frame and slab parity is captured whose final stage writes the frame's rows and counters
        into this rank's block of the local slab AND of every peer's slab (the fused exchange)."""
        self._pool = pool
        self._resident_out = {}
        S = len(self.lanes)
        for f in range(pool.shape[0]):
            ln = self.lanes[f % S]
            desc = engine.make_cloud_desc(self.fields, self.point_step, self.n_points, pool[f])
            if slabs is not None:
                for parity in range(slabs.parities):
                    out, cnt = slabs.local_out(parity, f)
                    ln.resident_graphs[(f, parity)] = ln.ctx.capture_pipeline([desc], self.pcfg, out, cnt, ln.d_plane,
                                                                              mirror=slabs.mirror(parity, f))
                continue
            out = arena[f] if arena is not None else ln.d_out
            cnt = counts_arena[f] if counts_arena is not None else ln.d_counts
            self._resident_out[f] = out
            ln.resident_graphs[f] = ln.ctx.capture_pipeline([desc], self.pcfg, out, cnt, ln.d_plane)
        torch.cuda.synchronize(self.device)

    def run_resident(self, frame_ids, main_stream=None, parity: int | None = None,
                     stage_to: torch.Tensor | None = None):
        """Replay the captured graphs for ``frame_ids`` across the lanes.  The caller's current
        stream is the fork/join point, so CUDA events recorded on it bracket the whole batch.
        ``parity``: which slab parity's graphs to replay (graphs captured with ``slabs``).
        ``stage_to`` [F, rows, 4] (NCCL fall-back only): each lane copies the first ``rows`` rows of
        the frame's output into ``stage_to[f]`` right behind the frame's graph."""
        main = torch.cuda.current_stream(self.device) if main_stream is None else main_stream
        S = len(self.lanes)
        fork = torch.cuda.Event()
        fork.record(main)
        for ln in self.lanes:
            ln.stream.wait_event(fork)
        for f in frame_ids:
            ln = self.lanes[f % S]
            with torch.cuda.stream(ln.stream):
                ln.ctx.launch_graph(ln.resident_graphs[f if parity is None else (f, parity)])
                if stage_to is not None:
                    stage_to[f].copy_(self._resident_out[f][:stage_to.shape[1]], non_blocking=True)
        for ln in self.lanes:
            ln.event.record(ln.stream)
            main.wait_event(ln.event)

    def check(self):
        for ln in self.lanes:
            ln.ctx.check()

"""The multi-GPU exchange fused into the pipeline's final stage (``apc_out_mirror``): mirrored rows and
counters are bit-identical to the primary output (one GPU: the mirrors are ordinary local buffers;
two GPUs: ``replay.PeerSlabs`` over NVLink, checked against an NCCL all-gather of the same outputs)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def dev_bytes(msg):
    return torch.frombuffer(bytearray(msg.data), dtype=torch.uint8).cuda()


STAGE_SETS = {
    "ground_last": dict(voxel_size=0.1, radius=dict(nb_points=5, radius=0.5),
                        ground=dict(distance_threshold=0.2, ransac_n=5, num_iterations=60, probability=0.99, seed=3)),
    "radius_last": dict(voxel_size=0.1, radius=dict(nb_points=4, radius=0.4)),
    "statistical_last": dict(voxel_size=0.15, statistical=dict(nb_neighbors=12, std_ratio=1.5)),
}


@pytest.mark.parametrize("case", sorted(STAGE_SETS))
@pytest.mark.parametrize("graph", [False, True])
def test_mirrored_output_equals_primary(case, graph):
    from autodriver_pointcloud_preprocessor_b200 import _capi as capi, engine, synth
    msg = synth.pack_cloud(synth.lidar_scan(seed=21, n_beams=32, n_az=1024), "xyzirt22")
    n = msg.width
    ctx = engine.Context(max_points=n)
    try:
        data = dev_bytes(msg)
        desc = engine.make_cloud_desc(msg.fields, msg.point_step, n, data)
        fcfg = engine.make_filter_cfg(skip_nans=True, dedup_mode=capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                                      crop=dict(min=[-50, -50, -10], max=[50, 50, 10], invert=False, mode=capi.CROP_OPEN3D))
        pcfg = engine.make_pipeline_cfg(fcfg, **STAGE_SETS[case])
        want, wc, _ = ctx.pipeline_run([desc], pcfg)
        ctx.check()
        wc = wc.cpu().numpy()
        n_out = int(wc[capi.CNT_OUTPUT])
        assert 0 < n_out < n
        out = torch.zeros((n, 4), device="cuda")
        cnt = torch.zeros(8, dtype=torch.int32, device="cuda")
        plane = torch.zeros(8, dtype=torch.float64, device="cuda")
        mirrors = [torch.full((n, 4), -7.0, device="cuda") for _ in range(3)]
        cmirrors = [torch.full((8,), -1, dtype=torch.int32, device="cuda") for _ in range(2)]
        m = engine.make_out_mirror([t.data_ptr() for t in mirrors], [t.data_ptr() for t in cmirrors])
        if graph:
            g = ctx.capture_pipeline([desc], pcfg, out, cnt, plane, mirror=m)
            for t in mirrors:
                t.fill_(-7.0)
            ctx.launch_graph(g)
        else:
            ctx.pipeline_run_mirrored([desc], pcfg, out, cnt, plane, m)
        ctx.check()
        assert np.array_equal(cnt.cpu().numpy(), wc)
        ref = want[:n_out].cpu().numpy().view(np.uint32)
        assert np.array_equal(out[:n_out].cpu().numpy().view(np.uint32), ref)
        for t in mirrors:
            got = t.cpu().numpy()
            assert np.array_equal(got[:n_out].view(np.uint32), ref)
            assert (got[n_out:] == -7.0).all()                      # nothing but the real rows is written
        for t in cmirrors:
            assert np.array_equal(t.cpu().numpy(), wc)
    finally:
        ctx.close()


def test_mirror_needs_a_selection_stage():
    from autodriver_pointcloud_preprocessor_b200 import _capi as capi, engine, synth
    msg = synth.pack_cloud(synth.lidar_scan(seed=2, n_beams=16, n_az=256), "xyzi16")
    ctx = engine.Context(max_points=msg.width)
    try:
        desc = engine.make_cloud_desc(msg.fields, msg.point_step, msg.width, dev_bytes(msg))
        pcfg = engine.make_pipeline_cfg(engine.make_filter_cfg(), voxel_size=0.2)
        out = torch.zeros((msg.width, 4), device="cuda")
        m = engine.make_out_mirror([torch.zeros((msg.width, 4), device="cuda").data_ptr()], [])
        with pytest.raises(capi.ApcError) as e:
            ctx.pipeline_run_mirrored([desc], pcfg, out, torch.zeros(8, dtype=torch.int32, device="cuda"),
                                      torch.zeros(8, dtype=torch.float64, device="cuda"), m)
        assert e.value.code == capi.APC_ERR_BAD_ARG
    finally:
        ctx.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _peer_worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    import bench
    from autodriver_pointcloud_preprocessor_b200 import _capi as capi, replay, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        n_beams, n_az = 32, 1024
        n = n_beams * n_az
        msgs = [synth.pack_cloud(synth.lidar_scan(seed=100 * rank + f, n_beams=n_beams, n_az=n_az), "xyzi16")
                for f in range(n_frames)]
        pool = torch.stack([torch.frombuffer(bytearray(m.data), dtype=torch.uint8).to(dev) for m in msgs])
        filter_kw = dict(skip_nans=True, dedup_mode=capi.DEDUP_OPEN3D, remove_nan=True, remove_inf=True,
                         transforms=[bench.TF], crop=bench.CROP)
        pipe = replay.ScanPipeline(msgs[0].fields, 16, n, filter_kw, bench.STAGES, lanes=2, device=rank)
        # reference outputs: plain per-frame arena, no exchange
        arena = torch.zeros((n_frames, n, 4), device=dev)
        carena = torch.zeros((n_frames, 8), dtype=torch.int32, device=dev)
        pipe.prepare_resident(pool, arena, carena)
        pipe.run_resident(list(range(n_frames)))
        torch.cuda.synchronize(dev)
        pipe.check()
        all_arena = torch.zeros((world,) + tuple(arena.shape), device=dev)
        all_cnt = torch.zeros((world,) + tuple(carena.shape), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_arena, arena)
        dist.all_gather_into_tensor(all_cnt, carena)
        # fused exchange
        slabs = replay.PeerSlabs(n_frames, n, dev, multicast=False)
        pipe.prepare_resident(pool, slabs=slabs)
        for parity in (0, 1, 0):
            slabs.buf[parity].fill_(-3.0)
            torch.cuda.synchronize(dev)
            dist.barrier()
            pipe.run_resident(list(range(n_frames)), parity=parity)
            slabs.barrier()
            torch.cuda.synchronize(dev)
            for src in range(world):
                for f in range(n_frames):
                    rows, cnt = slabs.frame(parity, src, f)
                    want_c = all_cnt[src, f].cpu().numpy()
                    assert np.array_equal(cnt.cpu().numpy(), want_c), (parity, src, f)
                    k = int(want_c[capi.CNT_OUTPUT])
                    assert np.array_equal(rows[:k].cpu().numpy().view(np.uint32),
                                          all_arena[src, f, :k].cpu().numpy().view(np.uint32)), (parity, src, f)
                    assert (rows[k:] == -3.0).all()
        pipe.close()
        q.put((rank, "ok"))
    except Exception as e:                                            # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_peer_slabs_two_gpus():
    """Every rank ends with every rank's rows and counters, bit-identical to an NCCL all-gather."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == [0, 1]
    for r, msg in res:
        assert msg == "ok", msg

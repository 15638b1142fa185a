"""Two-rank end-to-end check of both shard points, one process per rank:
chunk-sharded meshing on each rank -> packed shards all-gathered on the device -> vx_mesh_batch_assemble_shards ->
every rank renders its (work-balanced) screen stripe and its raster kernel stores the rows straight into the composed frame
in rank 0's memory (CUDA IPC peer mapping, arrival flags, no collective) -> the composed frame, its depth and the draw order
equal the oracle's, bit for bit.

With two or more GPUs the ranks use GPU 0 and 1 and NCCL; on a single-GPU box both ranks share GPU 0 (CUDA IPC works
between processes on one device) and a gloo group does the plumbing -- same library calls, same kernels, same checks."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, n_dev):
    import torch
    import torch.distributed as dist

    import vx_scenes
    from differential_projection_voxel_renderer_b200 import api, multigpu, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    di = rank % n_dev
    torch.cuda.set_device(di)
    dev = torch.device("cuda", di)
    if n_dev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, vd = 640, 360, 5
        pos, world_obj, p, v, nb = vx_scenes.terrain_scene(vd)
        n = p.shape[0]
        ctx = api.Context(di)
        dv, dn, dp = torch.from_numpy(v).to(dev), torch.from_numpy(nb).to(dev), torch.from_numpy(p).to(dev)
        ex = multigpu.MeshShardExchange(ctx, n, rank, world, dev)
        batch = ex.sweep(dv.data_ptr(), dp.data_ptr(), dn.data_ptr(), 0)
        batch = ex.sweep(dv.data_ptr(), dp.data_ptr(), dn.data_ptr(), 0)  # steady state: re-filled in place
        cam = vx_scenes.path_camera(1, w, h)
        vp = cam.view_projection()
        cfg = api.default_frame_config(w, h)
        from oracle import binding as ob
        ref = ob.mesh_chunks(v, nb, None, p)
        got = batch.download()
        assert np.array_equal(got["quad_count"], ref.quad_count) and np.array_equal(got["has_mesh"], ref.has_mesh)
        assert np.array_equal(got["slice_offsets"].reshape(n, -1), ref.slice_offsets.reshape(n, -1))
        assert np.array_equal(got["face_aabb"].reshape(n, -1), ref.face_aabb.reshape(n, -1))
        for i in range(n):
            assert np.array_equal(batch.chunk_quads(i).reshape(-1), ref.chunk_quads(i).reshape(-1)), f"chunk {i}"
        vis = ob.cull_chunks(p, vp, cam.position, vd)
        mesh_ids = np.flatnonzero((vis != 0) & (ref.has_mesh != 0)).astype(np.int32)
        oc, od, osurv = ob.render_frame(ref, mesh_ids, vp, cam.position, ob.default_frame_config(w, h, n_threads=4), ob.default_atlas())

        # full frame once on every rank: scratch sizing + the per-band work the balanced split is derived from
        _, _, order = api.render_frame(batch, vp, cam.position, cfg, mesh_ids=None, view_distance=vd, ctx=ctx)
        assert np.array_equal(order, osurv)
        band = sharding.stripe_band_cost(api.frame_bin_counts(ctx), api.frame_bin_tasks(ctx))
        assert band.shape == ((h + 7) // 8,) and band.sum() > 0
        comp = multigpu.StripeCompositor(ctx, w, h, rank, world, want_depth=True, timeout_us=20_000_000)
        for layout in (None, sharding.balanced_stripes(band, h, world), [(0, 8), (8, h - 8)], [(0, h), (h, 0)]):
            comp.set_stripes(layout)
            for k in range(3):  # more frames than buffers: exercises the acknowledgement path
                fno = getattr(comp, "_next", 0)
                comp._next = fno + 1
                comp.render(batch, vp, cam.position, cfg, vd, fno)
                if rank == 0:
                    comp.complete(fno)
                    ctx.synchronize()
                    comp.check()
                    c = comp.frame_tensor(fno, dev).cpu().numpy().view(np.uint32)
                    d = comp.depth_tensor(fno, dev).cpu().numpy().view(np.uint32)
                    assert np.array_equal(c, oc), f"colour differs, stripes {comp.stripes}"
                    assert np.array_equal(d, od.view(np.uint32)), f"depth differs, stripes {comp.stripes}"
                    comp.release(fno)
            ctx.synchronize()
            dist.barrier()
        # steady state without a host round trip per frame: arrival wait + acknowledgement in one kernel on rank 0
        comp.set_stripes(sharding.balanced_stripes(band, h, world))
        for k in range(5):
            fno = comp._next
            comp._next = fno + 1
            comp.render(batch, vp, cam.position, cfg, vd, fno)
            if rank == 0:
                comp.complete_and_release(fno)
        # ... and with the composing GPU's bookkeeping folded into its raster kernel (no hand-off kernel at all)
        for k in range(5):
            fno = comp._next
            comp._next = fno + 1
            fused = comp.render(batch, vp, cam.position, cfg, vd, fno, compose_release=fno if rank == 0 else None)
            assert fused == (rank == 0)
        ctx.synchronize()
        dist.barrier()
        comp.check()
        api.frame_stats(ctx)  # a hand-off timeout inside the raster kernel would surface here
        if rank == 0:
            last = comp._next - 1
            assert np.array_equal(comp.frame_tensor(last, dev).cpu().numpy().view(np.uint32), oc)
            assert np.array_equal(comp.depth_tensor(last, dev).cpu().numpy().view(np.uint32), od.view(np.uint32))
        dist.barrier()
        # frames in flight: two lanes per rank (a context each), one compositor -- composite buffers + flag words -- per lane;
        # frames alternate between the lanes and overlap on every GPU
        lanes = api.FrameLanes(di, 2, first=ctx)
        api.render_frame_device(batch, vp, cam.position, cfg, vd, lanes[1])  # sizes the second lane's scratch
        comps = [comp, multigpu.StripeCompositor(lanes[1], w, h, rank, world, want_depth=True, timeout_us=20_000_000)]
        comps[1].set_stripes(comp.stripes)
        nos = [comp._next, 0]
        for k in range(10):
            l = k % 2
            fused = comps[l].render(batch, vp, cam.position, cfg, vd, nos[l], compose_release=nos[l] if rank == 0 else None)
            assert fused == (rank == 0)
            nos[l] += 1
        lanes.synchronize()
        dist.barrier()
        for l in range(2):
            comps[l].check()
            api.frame_stats(lanes[l])
            if rank == 0:
                assert np.array_equal(comps[l].frame_tensor(nos[l] - 1, dev).cpu().numpy().view(np.uint32), oc), f"lane {l}"
                assert np.array_equal(comps[l].depth_tensor(nos[l] - 1, dev).cpu().numpy().view(np.uint32), od.view(np.uint32)), f"lane {l}"
        dist.barrier()
        comps[1].close()
        comp.close()
        lanes.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
        ex.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_sharded_mesh_exchange_and_peer_store_composite_two_ranks(tmp_path):
    import torch
    import torch.multiprocessing as mp
    n_dev = torch.cuda.device_count()
    assert n_dev >= 1
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), n_dev), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

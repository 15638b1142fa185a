"""Two-GPU end-to-end check of both shard points (needs >= 2 devices, skipped otherwise): chunk-sharded meshing on the
ranks' own GPUs -> all-gather of the mesh shards -> every rank renders its screen stripe -> NCCL gather to GPU0 ->
the composed frame equals the oracle's, bit for bit."""
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


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    import vx_scenes
    from differential_projection_voxel_renderer_b200 import api, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        w, h, vd = 640, 360, 5
        pos, world_obj, p, v, nb = vx_scenes.terrain_scene(vd)
        n = p.shape[0]
        ctx = api.Context(rank)
        dv, dn, dp = torch.from_numpy(v).to(dev), torch.from_numpy(nb).to(dev), torch.from_numpy(p).to(dev)
        ids = sharding.chunk_shard(n, rank, world)
        dids = torch.from_numpy(ids).to(dev)
        shard = api.BinaryGreedyMesher.mesh_batch_subset(dv.data_ptr(), dp.data_ptr(), dn.data_ptr(), 0, n, dids.data_ptr(), ids.size, ctx)
        merged = sharding.all_gather_mesh_shards(shard.download(), n)
        batch = api.upload_mesh_batch(ctx, merged["quads"], merged["quad_base"], merged["quad_count"], merged["slice_offsets"],
                                      merged["face_aabb"], merged["has_mesh"], p)
        cam = vx_scenes.path_camera(1, w, h)
        vp = cam.view_projection()
        cfg = api.default_frame_config(w, h)
        cfg.stripe_y0, cfg.stripe_rows = sharding.stripe_of(h, rank, world)
        color, depth, order = api.render_frame(batch, vp, cam.position, cfg, mesh_ids=None, view_distance=vd, ctx=ctx)
        frame = sharding.gather_stripes(torch.from_numpy(color.view(np.int32)).to(dev), h, w, dst=0)
        dframe = sharding.gather_stripes(torch.from_numpy(depth).to(dev), h, w, dst=0)
        if rank == 0:
            from oracle import binding as ob
            ref = ob.mesh_chunks(v, nb, None, p)
            vis = ob.cull_chunks(p, vp, cam.position, vd)
            mesh_ids = np.flatnonzero((vis != 0) & (ref.has_mesh != 0)).astype(np.int32)
            oc, od, osurv = ob.render_frame(ref, mesh_ids, vp, cam.position, ob.default_frame_config(w, h, n_threads=4), ob.default_atlas())
            assert np.array_equal(order, osurv)
            assert np.array_equal(frame.cpu().numpy().view(np.uint32), oc)
            assert np.array_equal(dframe.cpu().numpy().view(np.uint32), od.view(np.uint32))
        dist.barrier()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
        batch.release()
        shard.release()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_sharded_mesh_and_stripe_frame_two_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

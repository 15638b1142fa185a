"""The streaming world on the device (vx_world_batch_*, world.py) against the literal restatement of the
reference loop (tests/vx_refloop.py: host chunks + oracle mesher + oracle renderer), frame by frame along a camera walk:
loaded set, every cached mesh (stale ones included) and the rendered frame are bit-identical."""
import numpy as np
import pytest

import vx_refloop
from differential_projection_voxel_renderer_b200 import api, camera, world as vxw

pytestmark = pytest.mark.gpu

WALK = [(0.0, 10.0, 20.0), (0.0, 10.0, 20.0), (10.0, 12.0, 5.0), (40.0, 14.0, -20.0), (75.0, 20.0, -40.0), (75.0, 20.0, -40.0),
        (140.0, 30.0, -40.0), (140.0, 30.0, -40.0), (140.0, 30.0, -40.0), (20.0, 5.0, 0.0), (20.0, 5.0, 0.0), (20.0, 5.0, 0.0)]


def _oracle_batch(ob, ref):
    """The reference's mesh cache as an oracle mesh batch, chunks in (x, y, z) order."""
    allp = sorted(ref.mesh_cache)
    n = len(allp)
    quads, base, count = [], np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
    so = np.zeros((n, 6, 33), dtype=np.uint32)
    ab = np.zeros((n, 6, 6), dtype=np.int32)
    hm = np.zeros(n, dtype=np.uint8)
    total = 0
    for i, p in enumerate(allp):
        m = ref.mesh_cache[p]
        base[i] = total
        if m is not None:
            q, s, a = m
            quads.append(q.reshape(-1, 3))
            count[i] = q.reshape(-1, 3).shape[0]
            so[i], ab[i], hm[i] = s, a, 1
            total += int(count[i])
    flat = np.concatenate(quads).reshape(-1).astype(np.uint8) if quads else np.zeros(0, dtype=np.uint8)
    return allp, ob.MeshBatch(flat, base, count, so, ab, hm, np.array(allp, dtype=np.int32).reshape(-1, 3))


@pytest.mark.parametrize("vd,cap", [(2, 6), (3, 1000)])
def test_streaming_world_on_device_matches_the_reference_loop(ctx, ob, vd, cap):
    ref = vx_refloop.RefLoop(ob, vd, cap)
    w = vxw.World(vxw.WorldConfig(view_distance=vd, max_chunks_per_frame=cap), ctx=ctx)
    cache = vxw.MeshCache(w)
    W, H = 320, 180
    cfg = api.default_frame_config(W, H)
    ocfg, atlas = ob.default_frame_config(W, H, n_threads=2), ob.default_atlas()
    drawn_frames = 0
    for step, pos in enumerate(WALK):
        cam = camera.Camera(pos, W / H, yaw=0.4 * step, pitch=-0.2)
        vp = cam.view_projection()
        vis_ref = ref.frame(cam.position, vp)
        color, depth, surv = vxw.frame(w, cache, cam.position, vp, cfg, ctx)
        assert sorted(w.chunks) == sorted(ref.chunks), f"step {step}: loaded set differs"
        assert w.generated_last_update == ref.generated and cache.meshed_last_frame == ref.meshed
        # device-side filter A over the slot positions gives the reference's visible list
        assert w.get_visible_chunks_frustum(cam.position, vp) == vis_ref
        # every cached mesh on the device equals the reference's cache entry
        got = w.batch.download()
        for p, m in ref.mesh_cache.items():
            s = w.chunks[p]
            if m is None:
                assert got["has_mesh"][s] == 0, f"step {step}: {p} should have no mesh"
            else:
                assert got["has_mesh"][s] == 1
                b, c = int(got["quad_base"][s]), int(got["quad_count"][s])
                assert np.array_equal(got["quads"][b:b + c], m[0].reshape(-1, 3)), f"step {step}: mesh of {p} differs"
                assert np.array_equal(got["slice_offsets"][s], m[1]) and np.array_equal(got["face_aabb"][s], m[2])
        # chunks of the world without a cache entry, and empty slots, have no mesh
        cached_slots = {w.chunks[p] for p in ref.mesh_cache}
        assert all(got["has_mesh"][s] == 0 for s in range(w.capacity) if s not in cached_slots)
        # the frame: oracle renderer over the reference's cache, same list order
        allp, omb = _oracle_batch(ob, ref)
        oidx = {p: i for i, p in enumerate(allp)}
        oids = np.array([oidx[p] for p in sorted(vis_ref) if ref.mesh_cache.get(p) is not None], dtype=np.int32)  # Some(Some(mesh)), main.rs:281
        oc, od, osurv = ob.render_frame(omb, oids, vp, cam.position, ocfg, atlas)
        assert [allp[i] for i in osurv.tolist()] == [next(p for p, s in w.chunks.items() if s == sl) for sl in surv.tolist()]
        assert np.array_equal(depth.view(np.uint32), od.view(np.uint32)) and np.array_equal(color, oc), f"step {step}: frame differs"
        drawn_frames += int((color != cfg.clear_color).any())
    assert drawn_frames >= len(WALK) // 2
    w.batch.release()


def test_world_batch_compaction_keeps_meshes(ctx, ob):
    """Re-meshing the same chunks over and over fills the quad stream with dead space; compaction (a copy, never a
    re-mesh) must leave every mesh as it was."""
    w = vxw.World(vxw.WorldConfig(view_distance=2, max_chunks_per_frame=1000), ctx=ctx)
    w.update((0.0, 10.0, 20.0))
    cache = vxw.MeshCache(w)
    cache.update(w.get_all_chunks())
    before = w.batch.download()
    slots = np.array(sorted(w.chunks.values()), dtype=np.int32)
    live = int(before["quad_count"].sum())
    assert live > 0
    cap0 = None
    for _ in range(40):
        w.device.remesh(slots)
    after = w.batch.download()
    assert int(w.batch.info().total_quads) < 40 * live, "the stream was never compacted"
    assert np.array_equal(after["has_mesh"], before["has_mesh"]) and np.array_equal(after["quad_count"], before["quad_count"])
    for s in slots.tolist():
        a0, c0 = int(before["quad_base"][s]), int(before["quad_count"][s])
        a1 = int(after["quad_base"][s])
        assert np.array_equal(before["quads"][a0:a0 + c0], after["quads"][a1:a1 + c0])
    with pytest.raises(api.VxError):
        w.device.remesh(np.array([w.capacity], dtype=np.int32))
    w.batch.release()


def test_world_batch_grows_and_keeps_everything(ctx, ob):
    """vx_world_batch_grow (the reference's chunk map grows without bound under a moving camera, world.rs:84-87): a world
    started in a batch that is far too small ends up identical -- loaded set, meshes, rendered frame -- to one that had
    room from the start, and a moving camera follows the reference loop frame by frame across several growth steps."""
    vd = 2
    big = vxw.World(vxw.WorldConfig(view_distance=vd, max_chunks_per_frame=1000), ctx=ctx)
    small = vxw.World(vxw.WorldConfig(view_distance=vd, max_chunks_per_frame=1000), ctx=ctx, capacity=5)
    W, H = 320, 180
    cfg = api.default_frame_config(W, H)
    cb, cs = vxw.MeshCache(big), vxw.MeshCache(small)
    for step, pos in enumerate([(0.0, 10.0, 20.0), (40.0, 14.0, -20.0), (75.0, 20.0, -40.0)]):
        cam = camera.Camera(pos, W / H, yaw=0.4 * step, pitch=-0.2)
        vp = cam.view_projection()
        c0, d0, s0 = vxw.frame(big, cb, cam.position, vp, cfg, ctx)
        c1, d1, s1 = vxw.frame(small, cs, cam.position, vp, cfg, ctx)
        assert sorted(big.chunks) == sorted(small.chunks) and small.capacity >= small.chunk_count()
        assert np.array_equal(c0, c1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
        inv0 = {s: p for p, s in big.chunks.items()}
        inv1 = {s: p for p, s in small.chunks.items()}
        assert [inv0[s] for s in s0.tolist()] == [inv1[s] for s in s1.tolist()]
    assert small.capacity > 5
    big.batch.release()
    small.batch.release()
    # moving camera, never unloading: chunk for chunk like the reference loop
    ref = vx_refloop.RefLoop(ob, vd, 4)
    w = vxw.World(vxw.WorldConfig(view_distance=vd, max_chunks_per_frame=4), ctx=ctx)
    cap0 = w.capacity
    for step in range(110):
        pos = (16.0 * step, 10.0, 20.0)
        assert ref.update(pos) == w.update(pos)
        assert sorted(w.chunks) == sorted(ref.chunks)
    assert w.capacity > cap0 and w.chunk_count() > cap0
    cache = vxw.MeshCache(w)  # every chunk loaded before or after the growth meshes like the reference's
    vis = sorted(w.chunks)
    ref.update_cache(vis)
    cache.update(vis)
    got = w.batch.download()
    for p, m in ref.mesh_cache.items():
        s = w.chunks[p]
        assert (m is None) == (got["has_mesh"][s] == 0)
        if m is not None:
            b, c = int(got["quad_base"][s]), int(got["quad_count"][s])
            assert np.array_equal(got["quads"][b:b + c], m[0].reshape(-1, 3))
    w.batch.release()

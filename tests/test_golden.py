"""Committed fixtures (tests/golden/golden_v1.json, generated from the CPU oracle by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU test) and the CUDA path must reproduce them through the C ABI (GPU test)."""
import hashlib
import json
import os

import numpy as np
import pytest

import vx_kat as kat
import vx_scenes

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "golden_v1.json")))
VD, W, H = G["vd"], G["width"], G["height"]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_reproduces_golden(ob):
    for name, e in G["slices"].items():
        q = ob.greedy_mesh_slice(np.asarray(e["mask"], dtype=np.uint32))
        assert q.reshape(-1, 4).astype(int).tolist() == e["quads"], name
    pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
    ref = ob.mesh_chunks(v, nb, None, p)
    w = G["world"]
    assert p.shape[0] == w["chunks"] and int(ref.quad_count.sum()) == w["total_quads"]
    assert sha(ref.quad_count.astype(np.uint32)) == w["quad_count_sha"]
    assert sha(ref.slice_offsets.astype(np.uint32)) == w["slice_offsets_sha"]
    assert sha(ref.face_aabb.astype(np.int32)) == w["face_aabb_sha"]
    assert [sha(ref.chunk_quads(i)) for i in range(p.shape[0])] == w["chunk_quads_sha"]
    for name, vox in (("slab", kat.chunk_slab()), ("checker3d", kat.chunk_checker3d())):
        m = ob.mesh_chunks(vox.reshape(1, -1))
        assert int(m.quad_count[0]) == G["shapes"][name]["quads"] and sha(m.chunk_quads(0)) == G["shapes"][name]["sha"]
    for ci, e in G["frames"].items():
        vp = np.asarray(e["vp_bits"], dtype=np.uint32).view(np.float32)  # the stored matrix, not a recomputed one
        campos = np.asarray(e["cam"], dtype=np.float32)
        assert sha(ob.cull_chunks(pos, vp, campos, VD).astype(np.uint8)) == e["visible_sha"]
        visv = ob.cull_chunks(p, vp, campos, VD)
        ids = np.flatnonzero((visv != 0) & (ref.has_mesh != 0)).astype(np.int32)
        c, d, s = ob.render_frame(ref, ids, vp, campos, ob.default_frame_config(W, H, n_threads=3), ob.default_atlas())
        assert s.astype(int).tolist() == e["order"] and sha(c) == e["color_sha"] and sha(d) == e["depth_sha"]


@pytest.mark.gpu
def test_cuda_reproduces_golden(ctx):
    from differential_projection_voxel_renderer_b200 import api
    for name, e in G["slices"].items():
        q = api.BinaryGreedyMesher.greedy_mesh_slice(np.asarray(e["mask"], dtype=np.uint32), ctx)
        assert q.reshape(-1, 4).astype(int).tolist() == e["quads"], name
    pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
    batch = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
    got = batch.download()
    w = G["world"]
    assert int(got["quad_count"].sum()) == w["total_quads"]
    assert sha(got["quad_count"].astype(np.uint32)) == w["quad_count_sha"]
    assert sha(got["slice_offsets"].astype(np.uint32)) == w["slice_offsets_sha"]
    assert sha(got["face_aabb"].astype(np.int32)) == w["face_aabb_sha"]
    assert [sha(batch.chunk_quads(i)) for i in range(p.shape[0])] == w["chunk_quads_sha"]
    for name, vox in (("slab", kat.chunk_slab()), ("checker3d", kat.chunk_checker3d())):
        b = api.BinaryGreedyMesher.mesh_batch(vox.reshape(1, -1), None, None, None, ctx)
        b.download()
        assert b.chunk_quads(0).shape[0] == G["shapes"][name]["quads"] and sha(b.chunk_quads(0)) == G["shapes"][name]["sha"]
        b.release()
    for ci, e in G["frames"].items():
        vp = np.asarray(e["vp_bits"], dtype=np.uint32).view(np.float32)
        campos = np.asarray(e["cam"], dtype=np.float32)
        vis = api.get_visible_chunks_frustum(pos, campos, vp, VD, True, ctx)
        assert sha(vis.astype(np.uint8)) == e["visible_sha"] and int(vis.sum()) == e["visible"]
        c, d, s = api.render_frame(batch, vp, campos, api.default_frame_config(W, H), mesh_ids=None, view_distance=VD, ctx=ctx)
        assert s.astype(int).tolist() == e["order"]
        assert sha(c) == e["color_sha"] and sha(d) == e["depth_sha"]
        assert int((c != 0xFF87CEEB).sum()) == e["covered"]
    batch.release()


def test_oracle_reproduces_golden_v2(ob):
    """Regression fixtures of the later restatements (macrotile renderer, occlusion pass, barycentric mesh path, span
    walker, Hyper-Pipeline composition): tests/golden/golden_v2.json, generator tests/golden/make_golden_v2.py."""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_v2", os.path.join(here, "make_golden_v2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(here, "golden_v2.json")))
    got = mod.build()
    assert got == want
    assert all(e["barycentric"]["covered"] > 0 for e in want["frames"].values()) and want["span_walker"]["covered"] > 0

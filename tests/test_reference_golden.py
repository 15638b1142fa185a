"""Vectors produced by the REFERENCE crate itself (tools/ref_dump/ref_dump.rs, run with cargo on a machine that has a
Rust toolchain) against the CPU oracle and against the CUDA path.  This is what turns "bit-exact to the restatement" into
"bit-exact to the reference": terrain heights (noise 0.9 Perlin), voxels, quad lists, the glam view-projection matrix and
rendered frames.  Without tests/golden/ref_vectors.json the tests are reported as xfail, never as a silent pass."""
import json
import os

import numpy as np
import pytest

from differential_projection_voxel_renderer_b200 import camera, worldgen

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.environ.get("VX_REF_VECTORS") or os.path.join(HERE, "golden", "ref_vectors.json")  # the override is for checking the loader itself
FNV_OFF, FNV_PRIME = 0xcbf29ce484222325, 0x100000001b3


def fnv1a(data: bytes) -> str:
    h = FNV_OFF
    for b in data:
        h = ((h ^ b) * FNV_PRIME) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def vectors():
    if not os.path.exists(PATH):
        pytest.xfail("reference vectors not generated: no cargo in this image (tools/ref_dump/README.md has the one command)")
    return json.load(open(PATH))


def world_of(V):
    pos = np.asarray([c["pos"] for c in V["chunks"]], dtype=np.int32)
    order = np.lexsort((pos[:, 2], pos[:, 1], pos[:, 0]))
    w = worldgen.generate_world(pos[order])
    back = np.empty_like(order)
    back[order] = np.arange(order.size)
    return pos, w, back  # back[i] = index of V["chunks"][i] in the sorted world


def quads_of(entry):
    """reference dump -> (n, 6) rows of (face, slice, u, v, w, h, type) in mesher order"""
    out = []
    for face, sl, quads in entry or []:
        for q in quads:
            out.append([face, sl] + list(q))
    return np.asarray(out, dtype=np.int64).reshape(-1, 7)


def mesh_rows(slice_offsets, quads3, unpack):
    so = np.asarray(slice_offsets, dtype=np.int64).reshape(6, 33)
    q = unpack(np.asarray(quads3, dtype=np.uint8).reshape(-1, 3))  # (n, 5): u, v, w, h, type
    rows = []
    for f in range(6):
        for s in range(32):
            for k in range(int(so[f, s]), int(so[f, s + 1])):
                rows.append([f, s] + [int(x) for x in q[k]])
    return np.asarray(rows, dtype=np.int64).reshape(-1, 7)


def test_oracle_matches_reference_vectors(ob):
    V = vectors()
    # terrain heights: top solid voxel of a column == sample_terrain_height (chunk.rs:139-165)
    for e in V["heights"]:
        cx, cz = e["chunk_xz"]
        assert np.array_equal(np.asarray(e["top"], dtype=np.int32), ob.terrain_heights(cx * 32, cz * 32, 32, 32)), (cx, cz)
    # voxels + quads
    pos, w, back = world_of(V)
    nb = w.neighbor_table()
    ref = ob.mesh_chunks(w.voxels, nb, w.uniform_flags, w.positions)
    for i, c in enumerate(V["chunks"]):
        j = int(back[i])
        assert int(w.uniform_flags[j]) == c["uniform"], c["pos"]
        if c["uniform"] == 0:
            assert fnv1a(w.voxels[j].tobytes()) == c["voxels_fnv"], c["pos"]
        want = quads_of(c["quads_in_world"])
        got = mesh_rows(ref.slice_offsets[j], ref.chunk_quads(j), ob.unpack_quads) if ref.has_mesh[j] else np.zeros((0, 7), np.int64)
        assert np.array_equal(got, want), c["pos"]
        if c["uniform"] == 0:
            alone = ob.mesh_chunks(w.voxels[j:j + 1])
            got_a = mesh_rows(alone.slice_offsets[0], alone.chunk_quads(0), ob.unpack_quads) if alone.has_mesh[0] else np.zeros((0, 7), np.int64)
            assert np.array_equal(got_a, quads_of(c["quads_alone"])), c["pos"]
    # glam: view-projection bits
    cam = camera.Camera((0.0, 10.0, 20.0), 1280 / 720)
    assert cam.view_projection().reshape(16).view(np.uint32).tolist() == V["camera"]["vp_bits"]
    # frames: meshes drawn in list order
    for fr in V["frames"]:
        wd, ht = fr["width"], fr["height"]
        vp = np.asarray(fr["vp_bits"], dtype=np.uint32).view(np.float32)
        cfg = ob.default_frame_config(wd, ht)
        color = np.full((ht, wd), cfg.clear_color, dtype=np.uint32)
        depth = np.full((ht, wd), np.inf, dtype=np.float32)
        for i in fr["meshes"]:
            ob.render_mesh(ref, int(back[i]), vp, cfg, ob.default_atlas(), (0, 0, wd, ht), color, depth)
        assert int((color != cfg.clear_color).sum()) == fr["covered"]
        y0 = fr["row0"]
        assert np.array_equal(color[y0:y0 + 4], np.asarray(fr["color_rows"], dtype=np.uint32))
        assert np.array_equal(depth[y0:y0 + 4].view(np.uint32), np.asarray(fr["depth_bits_rows"], dtype=np.uint32))
        assert fnv1a(color.tobytes()) == fr["color_fnv"] and fnv1a(depth.tobytes()) == fr["depth_fnv"]


@pytest.mark.gpu
def test_cuda_matches_reference_vectors(ctx):
    from differential_projection_voxel_renderer_b200 import api
    V = vectors()
    pos, w, back = world_of(V)
    nb = w.neighbor_table()
    batch = api.BinaryGreedyMesher.mesh_batch(w.voxels, w.positions, nb, w.uniform_flags, ctx)
    got = batch.download()
    for i, c in enumerate(V["chunks"]):
        j = int(back[i])
        want = quads_of(c["quads_in_world"])
        rows = mesh_rows(got["slice_offsets"][j], batch.chunk_quads(j), api.unpack_quads) if got["has_mesh"][j] else np.zeros((0, 7), np.int64)
        assert np.array_equal(rows, want), c["pos"]
    for fr in V["frames"]:
        wd, ht = fr["width"], fr["height"]
        vp = np.asarray(fr["vp_bits"], dtype=np.uint32).view(np.float32)
        fb = api.Framebuffer(wd, ht)
        fb.clear(0xFF87CEEB)
        r = api.Rasterizer(ctx)
        for i in fr["meshes"]:
            r.render_mesh(batch, int(back[i]), vp, fb)
        color = fb.color_buffer.reshape(ht, wd)
        depth = fb.depth_buffer.reshape(ht, wd)
        assert fnv1a(color.tobytes()) == fr["color_fnv"] and fnv1a(depth.tobytes()) == fr["depth_fnv"]
    batch.release()

"""Pins the CPU oracle against the known-answer tests the reference's own test-suite holds for the hot
path.  Runs on CPU."""
import numpy as np
import pytest

import vx_kat as kat

FACES = range(6)


def mesh_one(ob, vox, nb_vox=None):
    return ob.mesh_chunks(vox.reshape(1, -1))


# ---- greedy slice KATs: src/meshing/binary_greedy.rs:814-855 -------------------------------------------
def test_greedy_empty(ob):
    assert ob.greedy_mesh_slice(np.zeros(32, np.uint32)).shape[0] == 0


def test_greedy_single(ob):
    m = np.zeros(32, np.uint32)
    m[0] = 1
    q = ob.greedy_mesh_slice(m)
    assert q.tolist() == [[0, 0, 1, 1]]


def test_greedy_vertical_line(ob):
    m = np.zeros(32, np.uint32)
    m[0] = 0b1111
    q = ob.greedy_mesh_slice(m)
    assert q.shape[0] == 1 and q[0, 2] == 1 and q[0, 3] == 4


def test_greedy_rectangle(ob):
    m = np.zeros(32, np.uint32)
    m[:3] = 0b1111
    q = ob.greedy_mesh_slice(m)
    assert q.shape[0] == 1 and q[0, 2] == 3 and q[0, 3] == 4


def test_greedy_microbench_masks(ob):  # benches/microbench.rs:21-38 workloads, answers by construction
    masks = kat.slice_masks()
    assert ob.greedy_mesh_slice(masks["full"]).tolist() == [[0, 0, 32, 32]]
    q = ob.greedy_mesh_slice(masks["checker_rows"])
    assert q.shape[0] == 16 and all(r[2] == 1 and r[3] == 32 for r in q.tolist())
    q = ob.greedy_mesh_slice(masks["sparse"])  # two full-height columns
    assert q.tolist() == [[0, 0, 32, 1], [0, 31, 32, 1]]
    assert ob.greedy_mesh_slice(masks["alt_bits"]).shape[0] == 512


def test_greedy_covers_mask_exactly(ob):
    rng = np.random.default_rng(7)
    for _ in range(50):
        m = rng.integers(0, 2 ** 32, size=32, dtype=np.uint64).astype(np.uint32)
        m &= rng.integers(0, 2 ** 32, size=32, dtype=np.uint64).astype(np.uint32)
        q = ob.greedy_mesh_slice(m)
        cover = np.zeros((32, 32), dtype=np.int32)
        for r, c, w, h in q.tolist():
            cover[r:r + w, c:c + h] += 1
        bits = ((m[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.int32)
        assert np.array_equal(cover, bits)


# ---- TinyQuad KATs: src/meshing/mesh.rs:694-728 ---------------------------------------------------------
@pytest.mark.parametrize("u,v,w,h,bt", [(0, 0, 1, 1, 0), (31, 31, 32, 32, 3), (5, 10, 3, 7, 2), (16, 16, 16, 16, 1)])
def test_tinyquad_roundtrip(ob, u, v, w, h, bt):
    p = ob.tinyquad_pack(u, v, w, h, bt)
    assert p.shape == (3,)  # size_of::<TinyQuad>() == 3
    assert ob.unpack_quads(p).tolist() == [[u, v, w, h, bt]]


# ---- chunk KATs: tests/meshing_tests.rs ------------------------------------------------------------------
def test_single_voxel_six_faces(ob):  # :55-86
    mb = mesh_one(ob, kat.chunk_single_voxel())
    assert mb.quad_count[0] == 6 and mb.has_mesh[0] == 1
    uq = ob.unpack_quads(mb.chunk_quads(0))
    for f in FACES:
        q = kat.quads_of_face(uq, mb.slice_offsets[0], f)
        assert q.shape[0] == 1 and q[0, 2] == 1 and q[0, 3] == 1 and q[0, 4] == kat.STONE


def test_face_positions(ob):  # :88-138  voxel at origin: faces on x/y/z = 1 (positive) or 0 (negative)
    mb = mesh_one(ob, kat.chunk_single_voxel(0, 0, 0))
    for f in FACES:
        s = kat.slice_of_quad(mb.slice_offsets[0], f, 0)
        slice_pos = s + 1 if f % 2 == 0 else s  # rasterizer.rs:896-900
        assert slice_pos == (1 if f % 2 == 0 else 0)


def test_top_and_bottom_face_height(ob):  # :140-191
    mb = mesh_one(ob, kat.chunk_single_voxel(5, 10, 5, kat.GRASS))
    assert kat.slice_of_quad(mb.slice_offsets[0], 2, 0) + 1 == 11  # +Y face plane at y = 11
    assert kat.slice_of_quad(mb.slice_offsets[0], 3, 0) == 10      # -Y face plane at y = 10


def test_internal_faces_culled(ob):  # :193-220
    mb = mesh_one(ob, kat.chunk_two_adjacent())
    assert mb.quad_count[0] == 6
    for f in (0, 1):
        s = kat.slice_of_quad(mb.slice_offsets[0], f, 0)
        assert (s + 1 if f == 0 else s) != 11


def test_2x2_merges(ob):  # :257-281
    mb = mesh_one(ob, kat.chunk_2x2_plane())
    uq = ob.unpack_quads(mb.chunk_quads(0))
    top = kat.quads_of_face(uq, mb.slice_offsets[0], 2)
    assert top.shape[0] == 1 and top[0, 2] == 2 and top[0, 3] == 2


def test_uniform_chunks_give_no_mesh(ob):  # :284-309
    for t in (kat.AIR, kat.STONE):
        vox = np.full((1, 32768), t, dtype=np.uint8)
        mb = ob.mesh_chunks(vox, None, np.array([1 + t], dtype=np.uint8))
        assert mb.has_mesh[0] == 0 and mb.quad_count[0] == 0


def test_types_not_merged(ob):  # :418-470
    mb = mesh_one(ob, kat.chunk_two_types())
    uq = ob.unpack_quads(mb.chunk_quads(0))
    top = kat.quads_of_face(uq, mb.slice_offsets[0], 2)
    assert top.shape[0] == 2 and sorted(top[:, 4].tolist()) == [kat.GRASS, kat.STONE]


def test_cross_chunk_boundary_culling(ob):  # :530-625  solid voxels touching across the +X border
    a = kat.empty_chunk(); kat.set_block(a, 31, 5, 5, kat.STONE)
    b = kat.empty_chunk(); kat.set_block(b, 0, 5, 5, kat.STONE)
    vox = np.stack([a.reshape(-1), b.reshape(-1)])
    nb = np.full((2, 6), -1, dtype=np.int32)
    nb[0, 0] = 1  # +X of chunk 0 is chunk 1
    nb[1, 1] = 0  # -X of chunk 1 is chunk 0
    mb = ob.mesh_chunks(vox, nb)
    assert mb.quad_count.tolist() == [5, 5]
    so = mb.slice_offsets
    assert so[0, 0, 32] - so[0, 0, 0] == 0  # no +X face on chunk 0
    assert so[1, 1, 32] - so[1, 1, 0] == 0  # no -X face on chunk 1
    alone = ob.mesh_chunks(vox)  # without neighbour info both keep 6
    assert alone.quad_count.tolist() == [6, 6]


def test_uniform_solid_neighbour_hides_border(ob):  # binary_greedy.rs:305-313
    a = kat.empty_chunk(); kat.set_block(a, 5, 31, 5, kat.STONE)
    nb = np.full((1, 6), -1, dtype=np.int32)
    nb[0, 2] = -3
    assert ob.mesh_chunks(a.reshape(1, -1), nb).quad_count[0] == 5
    nb[0, 2] = -2
    assert ob.mesh_chunks(a.reshape(1, -1), nb).quad_count[0] == 6


def test_dense_solid_six_big_quads(ob):  # benches/meshing.rs:28-41
    mb = mesh_one(ob, kat.chunk_dense_solid())
    uq = ob.unpack_quads(mb.chunk_quads(0))
    assert uq.tolist() == [[0, 0, 32, 32, kat.STONE]] * 6
    assert mb.face_aabb[0, 0].tolist() == [32, 0, 0, 32, 32, 32]  # +X face plane at x = 32
    assert mb.face_aabb[0, 1].tolist() == [0, 0, 0, 0, 32, 32]


def test_winding_matches_face_normal(ob):  # tests/meshing_tests.rs:311-373, mesh.rs:753-889
    from differential_projection_voxel_renderer_b200 import camera
    ident = np.eye(4, dtype=np.float32).reshape(16)
    normals = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float32)
    for f in FACES:
        clip = ob.quad_clip_vertices(f, 10 + (1 if f % 2 == 0 else 0), 3, 4, 2, 5, (0, 0, 0), ident)
        p = clip[:, :3]
        n = np.cross(p[1] - p[0], p[2] - p[0])
        n /= np.linalg.norm(n)
        assert float(np.dot(n, normals[f])) > 0.9


# ---- frustum KAT: src/camera/mod.rs:186-212 --------------------------------------------------------------
def test_frustum_culls_box_behind_camera(ob):
    from differential_projection_voxel_renderer_b200 import camera
    cam = camera.Camera((0, 0, 0), 16.0 / 9.0)
    planes = ob.frustum_from_vp(cam.view_projection())
    assert ob.frustum_intersects_aabb(planes, (-1, -1, -10), (1, 1, -8))
    assert not ob.frustum_intersects_aabb(planes, (-1, -1, 8), (1, 1, 10))


# ---- differential projection: tests/differential_projection_tests.rs:78-176, :436-453 ---------------------
def test_basis_projection_matches_full_mvp(ob):
    from differential_projection_voxel_renderer_b200 import camera
    vp = camera.mat4_mul(camera.perspective_rh(np.radians(70.0), 16 / 9, 0.1, 1000.0),
                         camera.look_at_rh((64, 50, 100), (64, 32, 64), (0, 1, 0))).reshape(16)
    m = vp.reshape(4, 4)
    rng = np.random.default_rng(3)
    for f in FACES:
        basis = ob.face_basis(f, (1, 0, 2), 7, vp)
        # tangent/bitangent world directions of the reference's (mirrored) bases: differential_projection.rs:231-290
        t = [(0, 1, 0), (0, 1, 0), (1, 0, 0), (1, 0, 0), (1, 0, 0), (-1, 0, 0)][f]
        b = [(0, 0, 1), (0, 0, -1), (0, 0, 1), (0, 0, -1), (0, 1, 0), (0, 1, 0)][f]
        o = np.array([32, 0, 64], dtype=np.float64)
        o[f // 2] += 7
        for _ in range(20):
            u, v = float(rng.integers(0, 33)), float(rng.integers(0, 33))
            w = o + u * np.array(t) + v * np.array(b)
            full = m.T.astype(np.float64) @ np.array([w[0], w[1], w[2], 1.0])
            got = ob.basis_project_point(basis, u, v)
            assert np.allclose(got, full, atol=1e-2)  # reference tolerance :137-176
    origin = ob.face_basis(2, (0, 0, 0), 5, vp)[0]  # slice -> origin :436-453
    assert np.allclose(origin, m.T @ np.array([0, 5, 0, 1], dtype=np.float32), atol=1e-4)


# ---- texture / shading constants: texture.rs:60-123, shading.rs:90-110, SURVEY 8 a16/a17 ------------------
def test_shading_constants(ob):
    cfg = ob.default_frame_config(64, 64)
    fps = [int(ob.lib().vxo_face_light(__import__("ctypes").byref(cfg), f) * 256.0) for f in range(6)]
    assert fps == [148, 89, 237, 89, 134, 89]
    # tests/shading_tests.rs:8-37 (test_shading_brighter_when_facing_light): with the default light a face looking up
    # (PosY, face 2) is brighter than one looking down (NegY, face 3) -- and nothing falls below the ambient floor
    assert fps[2] > fps[3] and min(fps) == int(cfg.ambient * 256.0)
    assert ob.lib().vxo_shade_color_u32(0xFFFFFFFF, 1.0) == 0xFFFFFFFF
    assert ob.lib().vxo_shade_color_u32(0xFF808080, 0.5) == 0xFF404040


def test_atlas_noise_texture(ob):
    a = ob.default_atlas()
    seed, idx = 12345, []
    for _ in range(32):
        seed = (seed * 1103515245 + 12345) & 0xFFFFFFFF
        idx.append((seed >> 16) & 0xFF)
    for t in (1, 2, 3):
        assert list(a.indices[t]) == idx  # same LCG stream for all three noise textures
    assert a.palette[1][0] == 0xFF007D00 and a.palette[1][1] == 0xFF005D00  # rgb565 0x03E0 / 0x02E0 expanded


# ---- rasterizer pixel-centre rule: tests/rasterizer_{gap,x_gap}_test.rs -----------------------------------
def _one_quad_batch(ob, vox):
    return ob.mesh_chunks(vox.reshape(1, -1))


def test_span_renderer_draws_single_voxel(ob):  # tests/rendering_pipeline_tests.rs:17-73 (320x180, > 0 px)
    from differential_projection_voxel_renderer_b200 import camera
    mb = _one_quad_batch(ob, kat.chunk_single_voxel(16, 16, 16))
    cam = camera.Camera((16.5, 16.5, 30.0), 320 / 180)
    cfg = ob.default_frame_config(320, 180)
    color = np.full((180, 320), cfg.clear_color, dtype=np.uint32)
    depth = np.full((180, 320), np.inf, dtype=np.float32)
    ob.render_mesh(mb, 0, cam.view_projection(), cfg, ob.default_atlas(), (0, 0, 320, 180), color, depth)
    assert int((color != cfg.clear_color).sum()) > 0
    assert np.isfinite(depth[color != cfg.clear_color]).all()


def test_stripes_leave_no_gaps(ob):  # tests/rasterizer_slice_gap_test.rs:1-79: stripe union == full-frame render
    from differential_projection_voxel_renderer_b200 import camera
    mb = _one_quad_batch(ob, kat.chunk_slab())
    cam = camera.Camera((16, 40, 80), 256 / 192)
    cfg = ob.default_frame_config(256, 192)
    atlas = ob.default_atlas()
    full_c = np.full((192, 256), cfg.clear_color, dtype=np.uint32); full_d = np.full((192, 256), np.inf, dtype=np.float32)
    ob.render_mesh(mb, 0, cam.view_projection(), cfg, atlas, (0, 0, 256, 192), full_c, full_d)
    st_c = np.full((192, 256), cfg.clear_color, dtype=np.uint32); st_d = np.full((192, 256), np.inf, dtype=np.float32)
    for y0 in range(0, 192, 7):
        ob.render_mesh(mb, 0, cam.view_projection(), cfg, atlas, (0, y0, 256, min(7, 192 - y0)), st_c, st_d)
    assert np.array_equal(full_c, st_c) and np.array_equal(full_d.view(np.uint32), st_d.view(np.uint32))
    assert int((full_c != cfg.clear_color).sum()) > 1000  # near slab covers a large area (:314-360)


def test_near_plane_clip(ob):  # src/rendering/rasterizer.rs:166-247: camera inside the slab still draws, no NaN depth
    from differential_projection_voxel_renderer_b200 import camera
    mb = _one_quad_batch(ob, kat.chunk_slab())
    cam = camera.Camera((16, 12.5, 16), 320 / 180)
    cfg = ob.default_frame_config(320, 180)
    color = np.full((180, 320), cfg.clear_color, dtype=np.uint32); depth = np.full((180, 320), np.inf, dtype=np.float32)
    ob.render_mesh(mb, 0, cam.view_projection(), cfg, ob.default_atlas(), (0, 0, 320, 180), color, depth)
    assert int((color != cfg.clear_color).sum()) > 0
    assert not np.isnan(depth).any()


def test_face_packets_single_voxel_and_split(ob):  # face_packets.rs:184-228
    c = kat.empty_chunk().reshape(32, 32, 32)
    c[16, 16, 16] = kat.STONE
    m = ob.mesh_chunks(c.reshape(1, -1))
    pk = ob.face_packets(m, 0)
    for f in range(6):
        assert len(pk[f]) == 1 and pk[f][0]["len"] == 1 and pk[f][0]["block_type"][0] == kat.STONE
        assert pk[f][0]["axis_pos"][0] == (17 if f % 2 == 0 else 16)
    # 2 * 32 + 5 isolated voxels in one y-slice: the +Y face splits into packets of 32, 32, 5, filled sequentially
    c = kat.empty_chunk().reshape(32, 32, 32)
    cells = [(2 * (i % 16), 2 * (i // 16)) for i in range(69)]
    for x, z in cells:
        c[z, 3, x] = kat.GRASS
    m = ob.mesh_chunks(c.reshape(1, -1))
    pk = ob.face_packets(m, 0)
    assert [p["len"] for p in pk[2]] == [32, 32, 5]
    assert all(int(p["axis_pos"][0]) == 4 for p in pk[2]) and all(int(p["axis_pos"][0]) == 3 for p in pk[3])
    assert pk[2][2]["u_len"][5:].sum() == 0  # unused lanes stay zero


def _ortho_vp(w, h, x0, x1, y0, y1, voxel=(4, 4)):
    """Column-major VP (w == 1 everywhere) that maps the unit +Z face of voxel `voxel` to the screen rectangle
    [x0, x1] x [y0, y1] (pixels, y down)."""
    vx, vy = voxel
    ax = 2.0 * (x1 - x0) / w                  # ndc_x = ax * (X - vx) + (2 x0 / w - 1)
    ay = 2.0 * (y1 - y0) / h                  # world y up -> screen y down: the top edge (Y = vy + 1) lands on y0
    m = np.zeros((4, 4), dtype=np.float64)    # row-major here
    m[0, 0], m[0, 3] = ax, (2.0 * x0 / w - 1.0) - ax * vx
    m[1, 1], m[1, 3] = ay, (1.0 - 2.0 * y1 / h) - ay * vy
    m[2, 2], m[2, 3] = 0.001, 0.5
    m[3, 3] = 1.0
    return np.ascontiguousarray(m.T.reshape(16).astype(np.float32))


@pytest.mark.parametrize("lo,hi,n", [(10.1, 10.9, 1), (10.0, 10.5, 1), (10.6, 11.6, 1), (10.1, 11.9, 2), (10.4, 10.6, 1), (10.0, 11.0, 1),
                                      (10.0, 10.4, 0), (10.6, 11.0, 0)])
def test_pixel_centre_coverage_rules(ob, lo, hi, n):
    """tests/rasterizer_gap_test.rs:6-106 and tests/rasterizer_x_gap_test.rs:3-80: a pixel is drawn iff its centre lies in
    the span, in x (`ceil(x0 - 0.5) ..= floor(x1 - 0.5)`, rasterizer.rs:1408-1409) and in y (half-open edge test at
    `y + 0.5`, :1357-1390).  The reference checks the arithmetic; here the restated rasterizer itself is driven with a
    quad that spans exactly [lo, hi] in one axis and a comfortable [20.25, 29.75] in the other."""
    w, h = 64, 48
    c = kat.empty_chunk()
    kat.set_block(c, 4, 4, 4, kat.STONE)
    mb = ob.mesh_chunks(c.reshape(1, -1))
    cfg = ob.default_frame_config(w, h)
    cfg.backface_culling = 0
    for axis in ("x", "y"):
        if axis == "y" and hi == 10.5:
            continue  # rows use the half-open edge test (:1363-1390): a quad ending exactly ON a centre is left to rounding
        vp = _ortho_vp(w, h, lo, hi, 20.25, 29.75) if axis == "x" else _ortho_vp(w, h, 20.25, 29.75, lo, hi)
        color = np.full((h, w), cfg.clear_color, dtype=np.uint32)
        depth = np.full((h, w), np.inf, dtype=np.float32)
        ob.render_mesh(mb, 0, vp, cfg, ob.default_atlas(), (0, 0, w, h), color, depth)
        cov = color != cfg.clear_color
        cols, rows = np.flatnonzero(cov.any(axis=0)), np.flatnonzero(cov.any(axis=1))
        narrow, wide = (cols, rows) if axis == "x" else (rows, cols)
        assert narrow.size == n, (axis, lo, hi, narrow)
        if n:
            first = int(np.ceil(np.float32(lo) - np.float32(0.5)))
            assert narrow.tolist() == list(range(first, first + n))
            assert wide.tolist() == list(range(20, 30))  # centres 20.5 .. 29.5


def _draw_all(ob, mb, cam, w, h, clear):
    cfg = ob.default_frame_config(w, h)
    cfg.clear_color = clear
    color = np.full((h, w), clear, dtype=np.uint32)
    depth = np.full((h, w), np.inf, dtype=np.float32)
    for m in np.flatnonzero(mb.has_mesh).tolist():  # render_mesh per mesh, in chunk order, like the reference tests
        ob.render_mesh(mb, m, cam.view_projection(), cfg, ob.default_atlas(), (0, 0, w, h), color, depth)
    return int((color != clear).sum())


def test_pipeline_pixel_count_thresholds(ob):
    """tests/rendering_pipeline_tests.rs: the reference's own thresholds on drawn pixels, with the restated camera
    (camera/mod.rs:20-61) and rasterizer."""
    from differential_projection_voxel_renderer_b200 import camera, worldgen
    # :17-57 render_single_voxel_writes_pixels: voxel (0,0,0), camera (32,32,80), 320x180, clear 0xFF000000 -> > 0
    mb = ob.mesh_chunks(kat.chunk_single_voxel(0, 0, 0, kat.GRASS).reshape(1, -1), None, None, np.zeros((1, 3), np.int32))
    assert _draw_all(ob, mb, camera.Camera((32.0, 32.0, 80.0), 320 / 180), 320, 180, 0xFF000000) > 0
    # :185-260 near voxel (chunk 0,0,0) + voxel in chunk (0,0,6), camera (16,16,50), 640x360 -> >= 50
    vox = np.stack([kat.chunk_single_voxel(16, 16, 16, kat.GRASS), kat.chunk_single_voxel(16, 16, 16, kat.STONE)])
    pos = np.array([[0, 0, 0], [0, 0, 6]], dtype=np.int32)
    mb = ob.mesh_chunks(vox, None, None, pos)
    assert _draw_all(ob, mb, camera.Camera((16.0, 16.0, 50.0), 640 / 360), 640, 360, 0xFF000000) >= 50
    # :263-311 a voxel in chunk (0,0,30) only (behind the -Z looking camera / sub-pixel) -> < 10
    mb = ob.mesh_chunks(kat.chunk_single_voxel(16, 16, 16, kat.STONE).reshape(1, -1), None, None, np.array([[0, 0, 30]], np.int32))
    assert _draw_all(ob, mb, camera.Camera((16.0, 16.0, 50.0), 640 / 360), 640, 360, 0xFF000000) < 10
    # :314-360 close geometry: camera (16,16,20), voxel (16,16,16) -> > 1000
    mb = ob.mesh_chunks(kat.chunk_single_voxel(16, 16, 16, kat.GRASS).reshape(1, -1), None, None, np.zeros((1, 3), np.int32))
    assert _draw_all(ob, mb, camera.Camera((16.0, 16.0, 20.0), 640 / 360), 640, 360, 0xFF000000) > 1000
    # :129-182 render_small_world_smoke_test: 3x3x3 terrain chunks, camera (32,32,80), 640x360 -> > 0
    grid = np.array([[x, y, z] for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)], dtype=np.int32)
    world = worldgen.generate_world(grid, store_uniform_voxels=True)
    mb = ob.mesh_chunks(world.voxels, world.neighbor_table(), world.uniform_flags, world.positions)
    assert _draw_all(ob, mb, camera.Camera((32.0, 32.0, 80.0), 640 / 360), 640, 360, 0xFF87CEEB) > 0


def test_face_basis_backface_sign_and_slices(ob):
    """tests/differential_projection_tests.rs:406-453: +Z / -Z basis normals have opposite z signs for a camera on +Z
    looking at the origin; with the identity matrix the origin of a +Y basis moves with the slice, tangent and bitangent
    do not."""
    from differential_projection_voxel_renderer_b200 import camera
    vp = camera.mat4_mul(camera.perspective_rh(np.radians(70.0), 16 / 9, 0.1, 1000.0),
                         camera.look_at_rh((0, 0, 10), (0, 0, 0), (0, 1, 0))).reshape(16)
    front, back = ob.face_basis(4, (0, 0, 0), 0, vp), ob.face_basis(5, (0, 0, 0), 0, vp)
    assert np.sign(front[3][2]) != np.sign(back[3][2])
    ident = np.eye(4, dtype=np.float32).reshape(16)
    b0, b15, b31 = (ob.face_basis(2, (0, 0, 0), s, ident) for s in (0, 15, 31))
    assert abs(b0[0][1] - 0.0) < 1e-3 and abs(b15[0][1] - 15.0) < 1e-3 and abs(b31[0][1] - 31.0) < 1e-3
    assert np.array_equal(b0[1], b15[1]) and np.array_equal(b0[2], b31[2])


# ---- span walker KATs: src/rendering/span_walker.rs:615-678, tests/span_walker_{differential_tests,bug_reproduction}.rs -----
def _blank(w, h):
    return np.zeros((h, w), dtype=np.uint32), np.full((h, w), np.inf, dtype=np.float32)  # Framebuffer::new


def test_fill_span_basic_and_depth_test(ob):  # span_walker.rs:615-656
    c, d = _blank(64, 64)
    ob.fill_span(c, d, 32, 10, 50, 0.5, 0xFF0000FF)
    assert (c[32, 10:50] == 0xFF0000FF).all() and (d[32, 10:50] == np.float32(0.5)).all()
    assert c[32, 9] == 0 and c[32, 50] == 0 and int((c != 0).sum()) == 40
    ob.fill_span(c, d, 32, 10, 50, 0.7, 0x00FF00FF)  # farther: rejected
    assert c[32, 25] == 0xFF0000FF and d[32, 25] == np.float32(0.5)
    ob.fill_span(c, d, 32, 10, 50, 0.3, 0x0000FFFF)  # nearer: wins
    assert c[32, 25] == 0x0000FFFF and d[32, 25] == np.float32(0.3)
    ob.fill_span(c, d, 32, 10, 50, 0.3, 0x12345678)  # equal depth: `<` keeps the first
    assert c[32, 25] == 0x0000FFFF
    # clamps (:422-428): x_start into [0, W-1], x_end into [0, W]; empty after the clamp -> nothing
    c, d = _blank(64, 4)
    ob.fill_span(c, d, 1, -20, 5, 0.5, 7)
    ob.fill_span(c, d, 2, 60, 500, 0.5, 7)
    ob.fill_span(c, d, 3, 70, 90, 0.5, 7)   # x_start clamps to 63, x_end to 64: the last pixel IS written
    ob.fill_span(c, d, 0, 30, 30, 0.5, 7)
    assert (c[1] == 7).sum() == 5 and (c[2] == 7).sum() == 4 and (c[3] == 7).sum() == 1 and c[3, 63] == 7 and (c[0] == 7).sum() == 0


def test_fill_span_partial_occlusion(ob):  # span_walker.rs:876-906
    c, d = _blank(128, 128)
    d[64, 0::2] = 0.3; c[64, 0::2] = 0xAAAAAA00
    d[64, 1::2] = 0.7; c[64, 1::2] = 0xBBBBBB00
    ob.fill_span(c, d, 64, 0, 128, 0.5, 0xFF00FF00)
    assert (c[64, 0::2] == 0xAAAAAA00).all() and (d[64, 0::2] == np.float32(0.3)).all()
    assert (c[64, 1::2] == 0xFF00FF00).all() and (d[64, 1::2] == np.float32(0.5)).all()


def test_span_walker_simple_quad(ob):  # span_walker.rs:658-681
    c, d = _blank(128, 128)
    ob.span_walk_quads(c, d, [-0.5], [-0.5], [0.5], [0.5], [0.5], [1])
    assert c[64, 64] == 0x00FF00FF and d[64, 64] == np.float32(0.5)
    assert int((c != 0).sum()) == 64 * 64 and (c[32:96, 32:96] != 0).all()


def test_span_walker_block_colours(ob):  # span_walker.rs:386-396
    assert [ob.span_walker_block_color(b) for b in (0, 1, 2, 3, 4, 255)] == [0, 0x00FF00FF, 0x8B4513FF, 0x808080FF, 0, 0]


def test_span_walker_single_quad_fills_about_100_px(ob):  # tests/span_walker_differential_tests.rs:11-56
    c, d = _blank(100, 100)
    ob.span_walk_quads(c, d, [-0.6], [-0.2], [-0.4], [0.0], [0.5], [1])
    assert 80 <= int((c != 0).sum()) <= 120


def test_span_walker_depth_testing_across_packets(ob):  # tests/span_walker_differential_tests.rs:58-112
    c, d = _blank(100, 100)
    ob.span_walk_quads(c, d, [-0.5], [-0.5], [0.5], [0.5], [0.7], [1])
    ob.span_walk_quads(c, d, [-0.3], [-0.3], [0.3], [0.3], [0.3], [2])
    assert abs(float(d[50, 50]) - 0.3) < 0.1 and c[50, 50] == 0x8B4513FF
    assert c[30, 30] == 0x00FF00FF and d[30, 30] == np.float32(0.7)


def test_span_walker_visibility_mask_and_two_quads(ob):  # tests/span_walker_differential_tests.rs:114-160 (+ mask)
    c, d = _blank(200, 200)
    ob.span_walk_quads(c, d, [-0.8, 0.2], [-0.4, -0.4], [-0.2, 0.8], [0.4, 0.4], [0.5, 0.5], [1, 2])
    assert int((c == 0x00FF00FF).sum()) > 0 and int((c == 0x8B4513FF).sum()) > 0
    c2, d2 = _blank(200, 200)
    ob.span_walk_quads(c2, d2, [-0.8, 0.2], [-0.4, -0.4], [-0.2, 0.8], [0.4, 0.4], [0.5, 0.5], [1, 2], visible=[0, 1])
    assert int((c2 == 0x00FF00FF).sum()) == 0 and np.array_equal(c2 == 0x8B4513FF, c == 0x8B4513FF)


def test_span_walker_fractional_start_and_vertical_gaps(ob):  # tests/span_walker_bug_reproduction.rs:10-140
    c, d = _blank(200, 200)
    ob.span_walk_quads(c, d, [-0.5], [-0.892], [0.5], [-0.85], [0.5], [1])
    assert int((c != 0).sum()) > 0
    c, d = _blank(200, 200)
    ob.span_walk_quads(c, d, [-0.5, -0.5], [-0.9, -0.8], [0.5, 0.5], [-0.85, -0.75], [0.5, 0.5], [1, 2])
    assert int((c != 0).sum()) >= 500 and int((c == 0x00FF00FF).sum()) > 0 and int((c == 0x8B4513FF).sum()) > 0
    c, d = _blank(200, 200)
    k = np.arange(3, dtype=np.float32) * np.float32(0.15)
    ob.span_walk_quads(c, d, [-0.4] * 3, np.float32(-0.9) + k, [0.4] * 3, np.float32(-0.87) + k, [0.5] * 3, [1, 2, 3])
    assert int((c != 0).sum()) >= 100


def test_span_walker_rows_are_sampled_at_pixel_centres(ob):  # span_walker.rs:237-247 (y + 0.5 in [start_y, end_y))
    # 200 px tall: NDC y in [0.795, 0.9] -> screen rows [10.0, 20.5(+eps)): rows 10 .. 20 (row 20's centre 20.5 < 20.501)
    c, d = _blank(200, 200)
    ob.span_walk_quads(c, d, [-0.5], [0.795], [0.5], [0.9], [0.5], [3])
    rows = np.flatnonzero((c != 0).any(axis=1))
    ymin = (np.float32(1.0) - np.float32(0.9)) * np.float32(0.5) * np.float32(200.0)
    ymax = (np.float32(1.0) - np.float32(0.795)) * np.float32(0.5) * np.float32(200.0) + np.float32(0.001)
    want = [y for y in range(200) if np.float32(y + 0.5) >= ymin and np.float32(y + 0.5) < ymax]
    assert list(rows) == want and len(want) in (10, 11)
    cols = np.flatnonzero((c != 0).any(axis=0))  # columns: round() of the screen box, [50, 150)
    assert cols[0] == 50 and cols[-1] == 149


def test_span_walker_more_than_eight_quads_batches(ob):  # span_walker.rs:181-192: batches of 8 do not change the result
    rng = np.random.default_rng(5)
    n = 70
    x0 = rng.uniform(-1.2, 1.0, n).astype(np.float32); x1 = x0 + rng.uniform(0.0, 0.6, n).astype(np.float32)
    y0 = rng.uniform(-1.2, 1.0, n).astype(np.float32); y1 = y0 + rng.uniform(0.0, 0.6, n).astype(np.float32)
    z = rng.choice(np.linspace(0.1, 0.9, 7).astype(np.float32), n)
    bt = rng.integers(0, 4, n).astype(np.uint8)
    c, d = _blank(160, 120)
    ob.span_walk_quads(c, d, x0, y0, x1, y1, z, bt)
    # serial restatement of the definition: quads in order, rows by pixel centre, columns by round(), depth test `<`
    c2, d2 = _blank(160, 120)
    W, H = np.float32(160), np.float32(120)
    for i in range(n):
        sx0 = max((x0[i] + np.float32(1)) * np.float32(0.5) * W, np.float32(0)); sy0 = max((np.float32(1) - y1[i]) * np.float32(0.5) * H, np.float32(0))
        sx1 = min((x1[i] + np.float32(1)) * np.float32(0.5) * W + np.float32(0.001), W); sy1 = min((np.float32(1) - y0[i]) * np.float32(0.5) * H + np.float32(0.001), H)
        if sx0 >= W or sy0 >= H or sx1 <= 0 or sy1 <= 0:
            continue
        xs = int(np.floor(sx0 + np.float32(0.5))); xe = int(np.floor(sx1 + np.float32(0.5)))  # round(), positive values
        xs = min(max(xs, 0), 159); xe = min(max(xe, 0), 160)
        for y in range(120):
            if np.float32(y + 0.5) >= sy0 and np.float32(y + 0.5) < sy1 and xs < xe:
                m = z[i] < d2[y, xs:xe]
                d2[y, xs:xe][m] = z[i]
                c2[y, xs:xe][m] = ob.span_walker_block_color(int(bt[i]))
    assert np.array_equal(c, c2) and np.array_equal(d.view(np.uint32), d2.view(np.uint32))


# ---- macrotile KATs: src/rendering/macrotile.rs:393-440 ----------------------------------------------------------
def test_macrotile_binning(ob):
    assert ob.macrotile_bin(0, 0, 100, 100, 1280, 720) == (1, (0, 0, 0, 0))          # :404-418 tile (0,0) only
    assert ob.macrotile_bin(64, 64, 192, 192, 1280, 720) == (1, (0, 0, 1, 1))        # :420-433 four tiles
    assert ob.macrotile_bin(0, 0, 1279, 719, 1280, 720)[0] == 2                       # :435-445 large primitive
    assert ob.macrotile_bin(150, 0, 250, 100, 1280, 720) == (1, (1, 0, 1, 0))        # :447-465 second mesh -> tile (1,0)
    assert ob.macrotile_bin(-50, -50, -1, -1, 1280, 720)[0] == 0                     # off-screen :197-199
    assert ob.macrotile_bin(0, 0, 639, 359, 1280, 720)[0] == 1                       # exactly 25 %: not "more than"
    assert ob.macrotile_bin(0, 0, 640, 359, 1280, 720)[0] == 2
    assert ob.macrotile_bin(1200, 700, 5000, 5000, 1280, 720) == (1, (9, 5, 9, 5))   # clamped to the last tile


def test_macrotile_frame_covers_the_same_pixels_as_the_stripe_frame(ob):
    """render_frame_macrotile (macrotile_renderer.rs:51-170) draws the same meshes through the same span rasterizer,
    tile by tile: the covered pixel set equals the stripe renderer's (main.rs), colours differ only where the per-tile
    restart of the span interpolation or the draw order (list order, no near-depth sort) flips a depth test."""
    import vx_scenes
    pos, world, p, v, nb = vx_scenes.terrain_scene(3)
    mb = ob.mesh_chunks(v, nb, None, p)
    w, h = 400, 300
    cam = vx_scenes.main_camera(w, h)
    vp = cam.view_projection()
    vis = ob.cull_chunks(p, vp, cam.position, 3)
    ids = np.flatnonzero((vis != 0) & (mb.has_mesh != 0)).astype(np.int32)
    cfg = ob.default_frame_config(w, h, n_threads=2)
    atlas = ob.default_atlas()
    c_s, d_s, surv = ob.render_frame(mb, ids, vp, cam.position, cfg, atlas)
    c_m, d_m, proj = ob.render_frame_macrotile(mb, ids, vp, cfg, atlas)
    assert sorted(proj.tolist()) == sorted(surv.tolist()) and proj.tolist() == [i for i in ids.tolist() if i in set(surv.tolist())]
    assert np.array_equal(c_m != cfg.clear_color, c_s != cfg.clear_color)
    assert np.array_equal(np.isfinite(d_m), np.isfinite(d_s))
    assert float((c_m != c_s).mean()) < 0.01 and float(np.abs(d_m[np.isfinite(d_m)] - d_s[np.isfinite(d_s)]).max()) < 1e-3
    # one tile (a 100 x 90 frame) is exactly the full-frame target of render_mesh, meshes drawn in list order
    cfg1 = ob.default_frame_config(100, 90)
    cam1 = vx_scenes.main_camera(100, 90)
    vp1 = cam1.view_projection()
    c1, d1, proj1 = ob.render_frame_macrotile(mb, ids, vp1, cfg1, atlas)
    c2 = np.full((90, 100), cfg1.clear_color, dtype=np.uint32); d2 = np.full((90, 100), np.inf, dtype=np.float32)
    for i in proj1:
        ob.render_mesh(mb, int(i), vp1, cfg1, atlas, (0, 0, 100, 90), c2, d2)
    assert proj1.size > 0
    kinds1 = ob.render_frame_macrotile(mb, ids, vp1, cfg1, atlas, want_kinds=True)[3]
    if not (kinds1 == 2).any():  # large primitives are drawn last, so the order differs when there are any
        assert np.array_equal(c1, c2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    else:
        assert np.array_equal(c1 != cfg1.clear_color, c2 != cfg1.clear_color)
        c3 = np.full((90, 100), cfg1.clear_color, dtype=np.uint32); d3 = np.full((90, 100), np.inf, dtype=np.float32)
        for i in list(proj1[kinds1 == 1]) + list(proj1[kinds1 == 2]):
            ob.render_mesh(mb, int(i), vp1, cfg1, atlas, (0, 0, 100, 90), c3, d3)
        assert np.array_equal(c1, c3) and np.array_equal(d1.view(np.uint32), d3.view(np.uint32))


# ---- barycentric mesh path: tests/rendering_pipeline_tests.rs:75-127 --------------------------------------------------
def _row_coverage(color, clear):
    return (color != clear).any(axis=1)


def test_span_renderer_matches_barycentric_row_coverage(ob):
    """A 32 x 32 Grass floor (y = 0) seen from (16, 40, 80) at 256 x 192: render_mesh (span renderer) and
    render_mesh_with_up with the non-level up vector (0, 0, 1) (barycentric renderer) cover the same scanlines."""
    from differential_projection_voxel_renderer_b200 import camera
    vox = np.zeros((32, 32, 32), dtype=np.uint8)  # [z][y][x]
    vox[:, 0, :] = 1
    mb = ob.mesh_chunks(vox.reshape(1, -1))
    w, h, clear = 256, 192, 0xFF000000
    cam = camera.Camera((16.0, 40.0, 80.0), w / h)
    vp = cam.view_projection()
    cfg = ob.default_frame_config(w, h)
    atlas = ob.default_atlas()
    span_c = np.full((h, w), clear, dtype=np.uint32); span_d = np.full((h, w), np.inf, dtype=np.float32)
    ref_c = span_c.copy(); ref_d = span_d.copy()
    ob.render_mesh(mb, 0, vp, cfg, atlas, (0, 0, w, h), span_c, span_d)
    ob.render_mesh_with_up(mb, 0, vp, cfg, atlas, (0.0, 0.0, 1.0), ref_c, ref_d)
    assert _row_coverage(span_c, clear).any()
    assert np.array_equal(_row_coverage(span_c, clear), _row_coverage(ref_c, clear))
    # a level up vector takes the span renderer: bit-identical to render_mesh (is_camera_level, rasterizer.rs:377-382)
    lvl_c = np.full((h, w), clear, dtype=np.uint32); lvl_d = np.full((h, w), np.inf, dtype=np.float32)
    ob.render_mesh_with_up(mb, 0, vp, cfg, atlas, (0.05, 0.998, 0.0), lvl_c, lvl_d)
    assert np.array_equal(lvl_c, span_c) and np.array_equal(lvl_d.view(np.uint32), span_d.view(np.uint32))
    # and the two renderers agree on nearly every pixel (same triangles, different interpolation arithmetic)
    assert float((ref_c != span_c).mean()) < 0.01
    both = np.isfinite(ref_d) & np.isfinite(span_d)
    assert float(np.abs(ref_d[both] - span_d[both]).max()) < 1e-4


def test_barycentric_target_rects_partition_the_frame_coverage(ob):
    """Unlike the span renderer the barycentric one restarts its edge-function chains at every target's box corner, so
    tiles are not bit-identical to the full frame -- but the covered pixel set is the same up to the rounding of those
    chains on exact edges, and every tile only writes inside its rect."""
    from differential_projection_voxel_renderer_b200 import camera
    mb = _one_quad_batch(ob, kat.chunk_slab())
    w, h = 200, 150
    cam = camera.Camera((16, 40, 80), w / h)
    vp = cam.view_projection()
    cfg = ob.default_frame_config(w, h)
    atlas = ob.default_atlas()
    full_c = np.full((h, w), cfg.clear_color, dtype=np.uint32); full_d = np.full((h, w), np.inf, dtype=np.float32)
    ob.render_mesh_tiny_quads(mb, 0, vp, cfg, atlas, (0, 0, w, h), False, full_c, full_d)
    tile_c = np.full((h, w), cfg.clear_color, dtype=np.uint32); tile_d = np.full((h, w), np.inf, dtype=np.float32)
    for (x0, y0, tw, th) in ((0, 0, 100, 75), (100, 0, 100, 75), (0, 75, 100, 75), (100, 75, 100, 75)):
        before = tile_c.copy()
        ob.render_mesh_tiny_quads(mb, 0, vp, cfg, atlas, (x0, y0, tw, th), False, tile_c, tile_d)
        changed = tile_c != before
        changed[y0:y0 + th, x0:x0 + tw] = False
        assert not changed.any()
    assert int((full_c != cfg.clear_color).sum()) > 1000
    assert float(((tile_c != cfg.clear_color) != (full_c != cfg.clear_color)).mean()) < 0.002


# ---- chunk-level occlusion pass: main.rs:473-478, :501-526 over OcclusionBuffer (occlusion.rs:60-153) ----------------
def test_occlusion_pass_only_drops_far_meshes_and_keeps_the_order(ob):
    """No reference test exercises the pass (it is off in the default run, main.rs:112); these are the properties its code
    guarantees: survivors are a subsequence of the un-occluded draw order, meshes nearer than two chunks are never
    dropped, the first mesh drawn is never dropped, and with the pass off the frame is unchanged."""
    import vx_scenes
    pos, world, p, v, nb = vx_scenes.terrain_scene(5)
    mb = ob.mesh_chunks(v, nb, None, p)
    w, h = 320, 180
    atlas = ob.default_atlas()
    dropped_any = False
    for cam_i in range(len(vx_scenes.CAMERA_PATH)):
        cam = vx_scenes.path_camera(cam_i, w, h)
        vp = cam.view_projection()
        vis = ob.cull_chunks(p, vp, cam.position, 5)
        ids = np.flatnonzero((vis != 0) & (mb.has_mesh != 0)).astype(np.int32)
        cfg = ob.default_frame_config(w, h)
        assert (cfg.occlusion_culling, cfg.occlusion_grid_w, cfg.occlusion_grid_h) == (0, 128, 72)
        c0, d0, s0 = ob.render_frame(mb, ids, vp, cam.position, cfg, atlas)
        cfg.occlusion_culling = 1
        c1, d1, s1 = ob.render_frame(mb, ids, vp, cam.position, cfg, atlas)
        it = iter(s0.tolist())
        assert all(m in it for m in s1.tolist()), "survivors must be a subsequence of the draw order"
        if s0.size:
            assert s1.size and s1[0] == s0[0]
        centers = (p[s0].astype(np.float32) * np.float32(32) + np.float32(16))
        dist_sq = ((centers - cam.position.astype(np.float32)) ** 2).sum(axis=1)
        near = set(s0[dist_sq < 64.0 * 64.0 * 0.999].tolist())
        assert near <= set(s1.tolist())
        dropped_any |= s1.size < s0.size
        # pixels can only change where a dropped mesh would have drawn
        if s1.size == s0.size:
            assert np.array_equal(c0, c1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    assert dropped_any


# ---- terrain: noise 0.9.0 Perlin::new(12345) restated in the oracle, reproduced by the host generator -----------------------
def test_noise_permutation_table_is_a_permutation_and_seed_dependent(ob):
    t = ob.noise_permutation_table(12345)
    assert sorted(t.tolist()) == list(range(256))
    assert not np.array_equal(t, ob.noise_permutation_table(12346))
    assert not np.array_equal(t, np.arange(256))
    from differential_projection_voxel_renderer_b200 import worldgen
    assert np.array_equal(worldgen._perm_table(12345)[:256], t) and np.array_equal(worldgen._perm_table(12345)[256:], t)
    assert np.array_equal(worldgen._perm_table(7)[:256], ob.noise_permutation_table(7))


def test_perlin_lattice_zeros_range_and_gradient_kat(ob):
    """Properties that follow from perlin_2d itself (core/perlin.rs): the noise vanishes on the integer lattice (every
    corner distance is 0 there), stays inside [-1, 1], and next to a lattice point it is the corner's gradient dotted with
    the offset, scaled by 2/sqrt(2), up to the quintic fade's O(t^3) blend."""
    for x, y in ((0.0, 0.0), (3.0, -7.0), (-12.0, 40.0), (255.0, 256.0)):
        assert ob.perlin2(x, y) == 0.0
    perm = ob.noise_permutation_table(12345).astype(np.int64)
    rng = np.random.default_rng(5)
    for _ in range(200):
        cx, cy = (int(v) for v in rng.integers(-300, 300, size=2))
        ex, ey = 1e-4 * rng.random(2)
        h = int(perm[perm[cx & 255] ^ (cy & 255)]) & 3
        g = (ex + ey, -ex + ey, ex - ey, -ex - ey)[h]
        got = ob.perlin2(cx + ex, cy + ey)
        assert abs(got - g * (2.0 / 1.4142135623730951)) < 1e-9
    pts = rng.uniform(-50, 50, size=(2000, 2))
    vals = np.array([ob.perlin2(a, b) for a, b in pts])
    assert vals.min() >= -1.0 and vals.max() <= 1.0 and vals.std() > 0.1


def test_host_terrain_generator_reproduces_the_oracle(ob):
    from differential_projection_voxel_renderer_b200 import worldgen
    for x0, z0 in ((0, 0), (-384, -384), (352, -96), (-1024, 992)):
        assert np.array_equal(worldgen.terrain_heights(x0, z0, 32, 32), ob.terrain_heights(x0, z0, 32, 32))
    pos = [(0, 0, 0), (0, -1, 0), (0, 1, 0), (3, 0, -5), (-12, 0, 11), (2, -2, 2), (7, -1, -7)]
    w = worldgen.generate_world(np.asarray(pos, dtype=np.int32))
    kinds = set()
    for i, p in enumerate(pos):
        flag, vox = ob.generate_terrain(p)
        assert flag == int(w.uniform_flags[i]), p
        kinds.add(flag)
        if flag == 0:
            assert np.array_equal(vox, w.voxels[i]), p
    assert kinds == {0, 1, 4}
    h = ob.terrain_heights(-64, -64, 128, 128)
    assert -20 <= int(h.min()) < 0 < int(h.max()) <= 20  # (n * 20) as i32

"""GPU parity of the flat-colour span walker (span_walker.rs) against the CPU oracle, through the
C ABI (vx_span_walk_quads, vx_span_walk_quads_device, vx_fill_spans).  Bit-exact: integer / byte work plus f32
screen mapping that must round like the reference."""
import ctypes as C

import numpy as np
import pytest

from differential_projection_voxel_renderer_b200 import api

pytestmark = pytest.mark.gpu


def _blank(w, h):
    return np.zeros((h, w), dtype=np.uint32), np.full((h, w), np.inf, dtype=np.float32)


def _same(fb, c, d):
    return np.array_equal(fb.color_buffer, c) and np.array_equal(fb.depth_buffer.view(np.uint32), d.view(np.uint32))


def test_fill_span_kats(ctx):  # span_walker.rs:615-656, :876-906
    fb = api.Framebuffer(64, 64)
    fb.fill_span(32, 10, 50, 0.5, 0xFF0000FF, ctx)
    assert (fb.color_buffer[32, 10:50] == 0xFF0000FF).all() and (fb.depth_buffer[32, 10:50] == np.float32(0.5)).all()
    assert fb.color_buffer[32, 9] == 0 and fb.color_buffer[32, 50] == 0 and int((fb.color_buffer != 0).sum()) == 40
    fb.fill_span(32, 10, 50, 0.7, 0x00FF00FF, ctx)
    assert fb.color_buffer[32, 25] == 0xFF0000FF and fb.depth_buffer[32, 25] == np.float32(0.5)
    fb.fill_span(32, 10, 50, 0.3, 0x0000FFFF, ctx)
    assert fb.color_buffer[32, 25] == 0x0000FFFF and fb.depth_buffer[32, 25] == np.float32(0.3)
    fb.fill_span(32, 10, 50, 0.3, 0x12345678, ctx)  # equal depth keeps the first
    assert fb.color_buffer[32, 25] == 0x0000FFFF
    fb = api.Framebuffer(128, 128)
    fb.depth_buffer[64, 0::2] = 0.3; fb.color_buffer[64, 0::2] = 0xAAAAAA00
    fb.depth_buffer[64, 1::2] = 0.7; fb.color_buffer[64, 1::2] = 0xBBBBBB00
    fb.fill_span(64, 0, 128, 0.5, 0xFF00FF00, ctx)
    assert (fb.color_buffer[64, 0::2] == 0xAAAAAA00).all() and (fb.color_buffer[64, 1::2] == 0xFF00FF00).all()
    assert (fb.depth_buffer[64, 1::2] == np.float32(0.5)).all() and (fb.depth_buffer[64, 0::2] == np.float32(0.3)).all()
    with pytest.raises(api.VxError):
        fb.fill_span(128, 0, 10, 0.5, 1, ctx)  # row outside the framebuffer


def test_fill_spans_random_bit_exact(ctx, ob):
    rng = np.random.default_rng(11)
    w, h, n = 200, 90, 3000
    y = rng.integers(0, h, n).astype(np.int32)
    xs = rng.integers(-40, w + 40, n).astype(np.int32)
    xe = xs + rng.integers(-5, 120, n).astype(np.int32)
    d = rng.choice(np.linspace(0.0, 1.0, 9).astype(np.float32), n)  # many exact ties: order decides
    d[rng.integers(0, n, 20)] = np.nan
    d[rng.integers(0, n, 20)] = -0.0
    col = rng.integers(1, 2**32, n, dtype=np.uint64).astype(np.uint32)
    c0, d0 = _blank(w, h)
    d0[::7, ::3] = 0.5      # existing contents take part in the depth test
    c0[::7, ::3] = 0xDEADBEEF
    d0[5, 5:50] = np.nan    # a stored NaN rejects everything
    fb = api.Framebuffer(w, h)
    fb.color_buffer[...] = c0; fb.depth_buffer[...] = d0
    fb.fill_spans(y, xs, xe, d, col, ctx)
    for i in range(n):
        ob.fill_span(c0, d0, int(y[i]), int(xs[i]), int(xe[i]), float(d[i]), int(col[i]))
    assert _same(fb, c0, d0)


def test_span_walker_reference_kats(ctx):  # span_walker.rs:658-681, tests/span_walker_{differential_tests,bug_reproduction}.rs
    fb = api.Framebuffer(128, 128)
    sw = api.SpanWalkerRasterizer(128, 128, ctx)
    sw.rasterize_projected_packet([-0.5], [-0.5], [0.5], [0.5], [0.5], [1], fb)
    assert fb.color_buffer[64, 64] == 0x00FF00FF and fb.depth_buffer[64, 64] == np.float32(0.5)
    assert int((fb.color_buffer != 0).sum()) == 64 * 64
    fb = api.Framebuffer(100, 100)
    sw = api.SpanWalkerRasterizer(100, 100, ctx)
    sw.rasterize_projected_packet([-0.6], [-0.2], [-0.4], [0.0], [0.5], [1], fb)
    assert 80 <= int((fb.color_buffer != 0).sum()) <= 120
    fb = api.Framebuffer(100, 100)
    sw.rasterize_projected_packet([-0.5], [-0.5], [0.5], [0.5], [0.7], [1], fb)
    sw.rasterize_projected_packet([-0.3], [-0.3], [0.3], [0.3], [0.3], [2], fb)
    assert abs(float(fb.depth_buffer[50, 50]) - 0.3) < 0.1 and fb.color_buffer[50, 50] == 0x8B4513FF
    fb = api.Framebuffer(200, 200)
    sw = api.SpanWalkerRasterizer(200, 200, ctx)
    sw.rasterize_projected_packet([-0.5], [-0.892], [0.5], [-0.85], [0.5], [1], fb)  # fractional start row
    assert int((fb.color_buffer != 0).sum()) > 0
    fb = api.Framebuffer(200, 200)
    sw.rasterize_projected_packet([-0.5, -0.5], [-0.9, -0.8], [0.5, 0.5], [-0.85, -0.75], [0.5, 0.5], [1, 2], fb)  # vertical gap
    assert int((fb.color_buffer == 0x00FF00FF).sum()) > 0 and int((fb.color_buffer == 0x8B4513FF).sum()) > 0
    with pytest.raises(ValueError):
        sw.rasterize_projected_packet([0], [0], [0], [0], [0], [1], api.Framebuffer(64, 64))


def _random_quads(rng, n, spread, size):
    x0 = rng.uniform(-spread, spread, n).astype(np.float32); x1 = x0 + rng.uniform(0.0, size, n).astype(np.float32)
    y0 = rng.uniform(-spread, spread, n).astype(np.float32); y1 = y0 + rng.uniform(0.0, size, n).astype(np.float32)
    z = rng.choice(np.linspace(0.05, 0.95, 13).astype(np.float32), n)
    bt = rng.integers(0, 5, n).astype(np.uint8)  # 0 = Air (colour 0 but still drawn), 4 = invalid -> Air
    return x0, y0, x1, y1, z, bt


@pytest.mark.parametrize("w,h,n,spread,size,seed", [
    (160, 120, 70, 1.2, 0.6, 5),        # the oracle KAT's case
    (1920, 1080, 32, 1.1, 1.5, 6),      # benches/span_walker.rs: 1080p, one packet of large quads
    (1920, 1080, 4000, 1.05, 0.08, 7),  # many small quads
    (333, 77, 500, 1.5, 0.9, 8),        # odd size, much clipping
])
def test_span_walker_random_quads_bit_exact(ctx, ob, w, h, n, spread, size, seed):
    rng = np.random.default_rng(seed)
    x0, y0, x1, y1, z, bt = _random_quads(rng, n, spread, size)
    vis = (rng.random(n) > 0.1).astype(np.uint8)
    c, d = _blank(w, h)
    ob.span_walk_quads(c, d, x0, y0, x1, y1, z, bt, vis)
    fb = api.Framebuffer(w, h)
    api.SpanWalkerRasterizer(w, h, ctx).rasterize_projected_packet(x0, y0, x1, y1, z, bt, fb, visible=vis)
    assert _same(fb, c, d)
    assert int((c != 0).sum()) > 0


def test_span_walker_special_values_and_existing_contents(ctx, ob):
    w, h = 96, 64
    inf, nan = np.float32(np.inf), np.float32(np.nan)
    x0 = np.array([-0.5, nan, -inf, 0.2, -0.9, 2.0, -0.5, -0.5, -1.0, 0.999], dtype=np.float32)
    x1 = np.array([0.5, 0.5, 0.0, nan, inf, 3.0, 0.5, 0.5, 1.0, 1.0], dtype=np.float32)
    y0 = np.array([-0.5, -0.2, -0.3, -0.4, nan, -0.5, -0.5, -0.5, -1.0, -1.0], dtype=np.float32)
    y1 = np.array([0.5, 0.2, 0.3, 0.4, 0.1, 0.5, 0.5, 0.5, 1.0, -0.999], dtype=np.float32)
    z = np.array([0.5, 0.4, 0.3, 0.2, 0.6, 0.1, nan, -0.0, 0.9, 0.05], dtype=np.float32)
    bt = np.array([1, 2, 3, 1, 2, 3, 1, 2, 3, 1], dtype=np.uint8)
    c, d = _blank(w, h)
    d[10:20, 10:40] = 0.0; c[10:20, 10:40] = 0x11111111     # +0.0 stored: a -0.0 fragment does not pass `<`
    d[30:34, :] = np.nan                                       # stored NaN rejects every fragment
    d[40:44, 20:60] = -np.inf; c[40:44, 20:60] = 0x22222222
    fb = api.Framebuffer(w, h)
    fb.color_buffer[...] = c; fb.depth_buffer[...] = d
    ob.span_walk_quads(c, d, x0, y0, x1, y1, z, bt)
    api.SpanWalkerRasterizer(w, h, ctx).rasterize_projected_packet(x0, y0, x1, y1, z, bt, fb)
    assert _same(fb, c, d)
    assert (np.signbit(d[np.isfinite(d) & (d == 0)]).any())  # the -0.0 quad landed somewhere and kept its sign


def test_span_walker_device_entry_matches_the_host_entry(ctx, ob):
    import torch
    rng = np.random.default_rng(21)
    w, h, n = 640, 360, 2000
    x0, y0, x1, y1, z, bt = _random_quads(rng, n, 1.1, 0.2)
    c, d = _blank(w, h)
    ob.span_walk_quads(c, d, x0, y0, x1, y1, z, bt)
    dev = torch.device("cuda", ctx.device)
    boxes = torch.from_numpy(np.concatenate([x0, y0, x1, y1, z])).to(dev)
    types = torch.from_numpy(np.concatenate([bt, np.ones(n, dtype=np.uint8)])).to(dev)
    col = torch.zeros((h, w), dtype=torch.int32, device=dev)
    dep = torch.full((h, w), float("inf"), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    before = ctx.launch_count
    ctx.check(ctx.lib.vx_span_walk_quads_device(ctx.handle, C.c_void_p(boxes.data_ptr()), C.c_void_p(types.data_ptr()), n, w, h,
                                                C.c_void_p(col.data_ptr()), C.c_void_p(dep.data_ptr())))
    ctx.synchronize()
    assert ctx.launch_count - before == 5
    assert np.array_equal(col.cpu().numpy().view(np.uint32), c) and np.array_equal(dep.cpu().numpy().view(np.uint32), d.view(np.uint32))
    # n == 0 and bad arguments
    ctx.check(ctx.lib.vx_span_walk_quads_device(ctx.handle, None, None, 0, w, h, C.c_void_p(col.data_ptr()), C.c_void_p(dep.data_ptr())))
    assert ctx.lib.vx_span_walk_quads_device(ctx.handle, None, None, 5, w, h, C.c_void_p(col.data_ptr()), C.c_void_p(dep.data_ptr())) != 0


# ---- the Hyper-Pipeline end to end on the device -------------------------------------------------------------
def test_hyper_pipeline_reference_bench_scene(ctx, ob):
    """benches/differential_projection.rs:8-36: the y < 8 + ((x + z) % 4) Stone chunk under
    persp(70 deg, 16/9, 0.1, 1000) * look_at((64,50,100) -> (64,32,64)), 1280 x 720: face packets -> packet pipeline -> span
    walker on the device equals the composition of the restated pieces, bit for bit."""
    import vx_kat as kat
    from differential_projection_voxel_renderer_b200 import camera
    vox = kat.chunk_slab().reshape(1, -1)
    batch = api.BinaryGreedyMesher.mesh_batch(vox, [(0, 0, 0)], None, None, ctx)
    ref = ob.mesh_chunks(vox)
    w, h = 1280, 720
    proj = camera.perspective_rh(np.radians(np.float32(70.0)), 16 / 9, 0.1, 1000.0)
    vp = camera.mat4_mul(proj, camera.look_at_rh((64.0, 50.0, 100.0), (64.0, 32.0, 64.0), (0.0, 1.0, 0.0))).reshape(16)
    c, d = _blank(w, h)
    n_packets, n_quads = ob.hyper_pipeline_render(ref, [0], vp, c, d)
    fb = api.Framebuffer(w, h)
    got = api.hyper_pipeline_render(batch, [0], vp, fb, ctx)
    assert got == n_quads and n_quads > 0
    assert _same(fb, c, d) and int((c != 0).sum()) > 1000
    batch.release()


@pytest.mark.parametrize("cam_i", [0, 1, 3, 5])
def test_hyper_pipeline_terrain_world_bit_exact(ctx, ob, cam_i):
    import vx_scenes
    pos, world, p, v, nb = vx_scenes.terrain_scene(5)
    batch = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
    ref = ob.mesh_chunks(v, nb, None, p)
    w, h = 640, 360
    vp = vx_scenes.path_camera(cam_i, w, h).view_projection()
    ids = np.flatnonzero(ref.has_mesh != 0).astype(np.int32)
    if cam_i == 1:   # list order is the caller's, duplicates and meshless chunks allowed
        ids = np.concatenate([ids[::-1], ids[:5], np.flatnonzero(ref.has_mesh == 0)[:3].astype(np.int32)])
    c, d = _blank(w, h)
    if cam_i == 3:   # existing contents take part in the depth test
        d[100:200, 100:400] = 0.2
        c[100:200, 100:400] = 0xFF445566
    fb = api.Framebuffer(w, h)
    fb.color_buffer[...] = c; fb.depth_buffer[...] = d
    n_packets, n_quads = ob.hyper_pipeline_render(ref, ids, vp, c, d)
    got = api.hyper_pipeline_render(batch, ids, vp, fb, ctx)
    assert got == n_quads
    assert _same(fb, c, d)
    assert api.hyper_pipeline_render(batch, np.zeros(0, dtype=np.int32), vp, fb, ctx) == 0 and _same(fb, c, d)
    with pytest.raises(api.VxError):
        api.hyper_pipeline_render(batch, [10**6], vp, fb, ctx)
    batch.release()

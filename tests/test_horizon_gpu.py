"""vx_horizon_cull (culling::apply_horizon_culling, culling.rs:40-119) against the oracle, and the reference's own
invariant: horizon culling never turns a covered pixel into sky (tests/horizon_culling_pipeline_movement_tests.rs:178-271)."""
import numpy as np
import pytest

import vx_scenes
from differential_projection_voxel_renderer_b200 import api, camera

pytestmark = pytest.mark.gpu


def test_horizon_cull_matches_oracle(ctx, ob):
    """Exact equality, kept set AND order: the angle of every candidate is evaluated with the platform's atan2f on both
    sides (the call culling.rs:86 makes), so even meshes that sit on an angular bin boundary -- lattice-aligned chunk
    centres on the diagonals do -- land in the same bin."""
    pos, world, p, v, nb = vx_scenes.terrain_scene(12)
    centers = (pos.astype(np.float32) * 32.0 + 16.0)  # main.rs:286-290
    rng = np.random.default_rng(5)
    n_boundary = 0
    for i in range(len(vx_scenes.CAMERA_PATH)):
        cam = vx_scenes.path_camera(i, 1280, 720)
        vis = ob.cull_chunks(pos, cam.view_projection(), cam.position, 12)
        ids = np.flatnonzero(vis).astype(np.int32)
        ids = ids[rng.permutation(ids.size)]  # caller order must not matter beyond distance ties
        want = ob.horizon_cull(cam.position, centers, ids)
        got = api.apply_horizon_culling(cam.position, centers, ids, ctx=ctx)
        assert np.array_equal(got, want), f"camera {i}"
        assert 0 < got.size <= ids.size
        c = centers[ids] - np.asarray(cam.position, np.float32)
        f = (np.arctan2(c[:, 2].astype(np.float64), c[:, 0].astype(np.float64)) + np.pi) / (2 * np.pi) * 128
        n_boundary += int((np.abs(f - np.round(f)) < 1e-5).sum())
    # a camera on the lattice: chunk centres exactly on the axes and diagonals, i.e. exactly on bin boundaries
    ids = np.arange(pos.shape[0], dtype=np.int32)
    cam_pos = (16.0, 40.0, 16.0)
    c = centers - np.asarray(cam_pos, np.float32)
    f = (np.arctan2(c[:, 2].astype(np.float64), c[:, 0].astype(np.float64)) + np.pi) / (2 * np.pi) * 128
    n_boundary += int((np.abs(f - np.round(f)) < 1e-5).sum())
    assert n_boundary > 100  # the boundary cases the exact comparison is about are really there
    for bins in (128, 64, 8):
        assert np.array_equal(api.apply_horizon_culling(cam_pos, centers, ids, bins=bins, ctx=ctx), ob.horizon_cull(cam_pos, centers, ids, bins=bins))
    # degenerate inputs: empty list, everything closer than min_dist_chunks
    assert api.apply_horizon_culling((0, 0, 0), centers, np.zeros(0, np.int32), ctx=ctx).size == 0
    near = np.array([[1.0, 2.0, 3.0], [10.0, -5.0, 4.0], [0.0, 50.0, 0.0]], np.float32)
    assert api.apply_horizon_culling((0, 0, 0), near, ctx=ctx).tolist() == ob.horizon_cull((0, 0, 0), near).tolist()
    # random clouds: other configs
    for trial in range(4):
        pts = rng.uniform(-800, 800, size=(3000, 3)).astype(np.float32)
        pts[:, 1] = rng.uniform(-60, 120, size=3000)
        kw = dict(bins=(128, 64, 360, 1)[trial], base_margin=0.1 * (trial + 1), margin_dist_factor=0.05, min_dist_chunks=2.0 + trial)
        want = ob.horizon_cull((3.0, 20.0, -7.0), pts, None, **kw)
        got = api.apply_horizon_culling((3.0, 20.0, -7.0), pts, None, ctx=ctx, **kw)
        assert np.array_equal(got, want)


def test_horizon_culling_does_not_remove_visible_pixels_during_movement(ctx, ob):
    """The reference's invariant on its own camera path (640x360, view distance 8, camera y = 32 moving along the
    +X/-Z diagonal), rendered by the CUDA frame path with and without the horizon-culled list."""
    w, h, vd = 640, 360, 8
    pos, world, p, v, nb = vx_scenes.terrain_scene(vd)
    batch = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
    has = batch.download()["has_mesh"] != 0
    centers = p.astype(np.float32) * 32.0 + 16.0
    cfg = api.default_frame_config(w, h)
    for campos in ((0.0, 32.0, 80.0), (8.0, 32.0, 72.0), (16.0, 32.0, 64.0), (24.0, 32.0, 56.0), (32.0, 32.0, 48.0)):
        cam = camera.Camera(campos, w / h)
        vp = cam.view_projection()
        vis = api.get_visible_chunks_frustum(p, cam.position, vp, vd, True, ctx)
        ids = np.flatnonzero((vis != 0) & has).astype(np.int32)
        kept = api.apply_horizon_culling(cam.position, centers, ids, ctx=ctx)
        assert kept.size <= ids.size
        base, _, _ = api.render_frame(batch, vp, cam.position, cfg, mesh_ids=ids, ctx=ctx, want_depth=False)
        hz, _, _ = api.render_frame(batch, vp, cam.position, cfg, mesh_ids=kept, ctx=ctx, want_depth=False)
        missing = int(((base != cfg.clear_color) & (hz == cfg.clear_color)).sum())
        assert missing == 0, f"horizon culling removed {missing} visible pixels at {campos}"
    batch.release()

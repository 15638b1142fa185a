"""Small end-to-end run for compute-sanitizer (tools/sanitize.sh): meshing with neighbours, one full frame, a stripe frame,
a read-modify-write mesh render, the occlusion pass, two pipelined frames -- every kernel of the frame path at least once,
checked against the oracle so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vx_scenes  # noqa: E402
from differential_projection_voxel_renderer_b200 import api  # noqa: E402
from oracle import binding as ob  # noqa: E402

W, H, VD = 320, 184, 3
pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
ctx = api.Context(0)
batch = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
ref = ob.mesh_chunks(v, nb, None, p)
for i in range(p.shape[0]):
    assert np.array_equal(batch.chunk_quads(i).reshape(-1), ref.chunk_quads(i).reshape(-1))
cam = vx_scenes.path_camera(1, W, H)
vp = cam.view_projection()
vis = ob.cull_chunks(p, vp, cam.position, VD)
ids = np.flatnonzero((vis != 0) & (ref.has_mesh != 0)).astype(np.int32)
oc, od, osurv = ob.render_frame(ref, ids, vp, cam.position, ob.default_frame_config(W, H, n_threads=2), ob.default_atlas())
cfg = api.default_frame_config(W, H)
for _ in range(2):  # the second frame uses the plan made from the first one's counters
    c, d, s = api.render_frame(batch, vp, cam.position, cfg, mesh_ids=None, view_distance=VD, ctx=ctx)
    assert np.array_equal(c, oc) and np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(s, osurv)
cs = api.VxFrameConfig.from_buffer_copy(cfg)
cs.stripe_y0, cs.stripe_rows = 96, 88
c, d, _ = api.render_frame(batch, vp, cam.position, cs, mesh_ids=ids, ctx=ctx)
assert np.array_equal(c, oc[96:]) and np.array_equal(d.view(np.uint32), od[96:].view(np.uint32))
co = api.VxFrameConfig.from_buffer_copy(cfg)
co.occlusion_culling = 1
api.render_frame(batch, vp, cam.position, co, mesh_ids=ids, ctx=ctx)
fb = api.Framebuffer(W, H)
fb.clear(cfg.clear_color)
api.Rasterizer(ctx).render_mesh(batch, int(ids[0]), vp, fb)
loop = api.FrameLoop(batch, cfg, view_distance=VD, want_depth=True, ctx=ctx)
t0 = loop.submit(vp, cam.position)
t1 = loop.submit(vp, cam.position)
for t in (t0, t1):
    c, d, s = loop.wait(t)
    assert np.array_equal(c, oc) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
print("sanitize target ok:", p.shape[0], "chunks,", ids.size, "meshes,", ctx.launch_count, "launches")
batch.release()
ctx.close()

"""Short program for ncu: one world remesh (818 chunks) + a few 1280x720 vd12 frames.  Not a benchmark."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vx_scenes  # noqa: E402
from differential_projection_voxel_renderer_b200 import api  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pos, world, p, v, nb = vx_scenes.terrain_scene(12)
cam = vx_scenes.main_camera(1280, 720)
ctx = api.Context(0)
# the bench's batch: every lattice chunk (Uniform ones carry a flag), so filter A runs over all 7,153 of them
batch = api.BinaryGreedyMesher.mesh_batch(world.voxels, pos, world.neighbor_table(), world.uniform_flags, ctx, validate=False)
print("quads", batch.info().total_quads)
vb = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)  # and one re-mesh of the Varied chunks for the mesher's line in the launch list
vb.release()
cfg = api.default_frame_config(1280, 720)
cfg.frames_in_flight = int(os.environ.get("VX_FRAMES_IN_FLIGHT", "0"))  # > 1: the coarser work items of a frame that shares the GPU
for _ in range(frames):
    api.render_frame_device(batch, cam.view_projection(), cam.position, cfg, 12, ctx)
ctx.synchronize()
st = api.frame_stats(ctx)
print("survivors", st.n_survivors, "tris", st.n_triangles, "entries", st.n_bin_entries, "max_bin", st.reserved[0])
bc = api.frame_bin_counts(ctx)
np.set_printoptions(linewidth=250)
print("tile grid", bc.shape, "rows: max per tile row")
print(bc.max(axis=1))
print("per tile-row sum", bc.sum(axis=1))
print("hist", np.histogram(bc, bins=[0, 1, 8, 64, 256, 1024, 4096, 16384, 1 << 20])[0])
batch.release()
ctx.close()

"""Short program for ncu: mesh a batch of N copies of a terrain chunk twice (large-batch regime).  Not a benchmark."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vx_scenes  # noqa: E402
from differential_projection_voxel_renderer_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pos, world, p, v, nb = vx_scenes.terrain_scene(3)
ctx = api.Context(0)
b0 = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
qc = b0.download()["quad_count"]
best = int(np.argsort(qc)[len(qc) // 2]) if len(sys.argv) > 2 and sys.argv[2] == "median" else int(np.argmax(qc))
dev = torch.device("cuda", 0)
d_big = torch.from_numpy(v[best]).to(dev).repeat(n, 1).contiguous()
h = C.c_void_p()
ctx.check(ctx.lib.vx_mesh_chunks_device(ctx.handle, C.c_void_p(d_big.data_ptr()), None, None, None, n, C.byref(h)))
big = api.MeshBatch(ctx, h)
ctx.check(ctx.lib.vx_remesh_chunks_device(ctx.handle, C.c_void_p(d_big.data_ptr()), None, None, big.handle))
ctx.synchronize()
print("chunks", n, "quads", big.info().total_quads)

"""Quick device timing of the mesher on a large batch (development tool).  VX_B200_LIB=path python tools/mesh_variant_bench.py [n]"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vx_scenes
from differential_projection_voxel_renderer_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
pos, world, p, v, nb = vx_scenes.terrain_scene(12)
ctx = api.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
b0 = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
qc = b0.download()["quad_count"]
res = []
for name, idx in (("max-quad chunk", int(np.argmax(qc))), ("median chunk", int(np.argsort(qc)[len(qc) // 2]))):
    d_big = torch.from_numpy(v[idx]).to(dev).repeat(n, 1).contiguous()
    h = C.c_void_p()
    ctx.check(ctx.lib.vx_mesh_chunks_device(ctx.handle, C.c_void_p(d_big.data_ptr()), None, None, None, n, C.byref(h)))
    big = api.MeshBatch(ctx, h)
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        ctx.check(ctx.lib.vx_remesh_chunks_device(ctx.handle, C.c_void_p(d_big.data_ptr()), None, None, big.handle))
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts[1:]))
    res.append(f"{name} ({int(qc[idx])} quads): {n / ms / 1e3:.2f} M chunks/s")
    big.release(); del d_big
# mixed batch: the whole vd12 world tiled 16x (with neighbours)
print(os.environ.get("VX_B200_LIB", "default"), " | ".join(res))

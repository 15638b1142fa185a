"""Quick device-time comparison of library variants (development tool): frames of the bench scene, L2 flushed.
   VX_B200_LIB=path python tools/variant_bench.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vx_scenes
from differential_projection_voxel_renderer_b200 import api
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
H = int(sys.argv[2]) if len(sys.argv) > 2 else 720
VD = int(sys.argv[3]) if len(sys.argv) > 3 else 12
pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
cam = vx_scenes.main_camera(W, H)
ctx = api.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
batch = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
cfg = api.default_frame_config(W, H)
cfga = api.VxFrameConfig.from_buffer_copy(cfg); cfga.async_submit = 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
vp = cam.view_projection()
api.render_frame_device(batch, vp, cam.position, cfg, VD, ctx)
for _ in range(5):
    api.render_frame_device(batch, vp, cam.position, cfga, VD, ctx)
K = 100
s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
for i in range(K):
    with torch.cuda.stream(stream):
        flush.fill_(1)
    s[i].record(stream); api.render_frame_device(batch, vp, cam.position, cfga, VD, ctx); e[i].record(stream)
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in zip(s, e)])
cfgp = api.VxFrameConfig.from_buffer_copy(cfg); cfgp.profile_kernels = 1
ks = np.zeros(4)
for _ in range(20):
    with torch.cuda.stream(stream):
        flush.fill_(1)
    api.render_frame_device(batch, vp, cam.position, cfgp, VD, ctx); ks += api.frame_kernel_times(ctx)
import time
color_host = ctx.host_array((H, W), np.uint32)
surv_host = np.empty(p.shape[0], dtype=np.int32)
for _ in range(5):
    api.render_frame(batch, vp, cam.position, cfg, mesh_ids=None, view_distance=VD, color_out=color_host, want_depth=False, ctx=ctx, survivors_out=surv_host)
t0 = time.perf_counter()
for _ in range(200):
    api.render_frame(batch, vp, cam.position, cfg, mesh_ids=None, view_distance=VD, color_out=color_host, want_depth=False, ctx=ctx, survivors_out=surv_host)
e2e_us = (time.perf_counter() - t0) / 200 * 1e6
t0 = time.perf_counter()
for _ in range(200):
    api.render_frame_device(batch, vp, cam.position, cfg, VD, ctx)   # synchronous, nothing copied back
sync_us = (time.perf_counter() - t0) / 200 * 1e6
t0 = time.perf_counter()
for _ in range(200):
    api.render_frame_device(batch, vp, cam.position, cfga, VD, ctx)  # asynchronous submit only
ctx.synchronize()
async_us = (time.perf_counter() - t0) / 200 * 1e6
print("wall us per frame: e2e (mapped host colour) %.1f | synchronous device frame %.1f | async submit, back to back %.1f" % (e2e_us, sync_us, async_us))
print(os.environ.get("VX_B200_LIB", "default"), f"{W}x{H} vd{VD}", "e2e us %.1f" % e2e_us, "frame ms mean %.4f median %.4f min %.4f | kernels us" % (ms.mean(), np.median(ms), ms.min()), np.round(ks / 20 * 1000, 1))

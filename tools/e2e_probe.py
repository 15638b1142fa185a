"""Where the pipelined e2e frame period goes (development probe): host submit cost, D2H bandwidth, steady-state period."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vx_scenes
from differential_projection_voxel_renderer_b200 import api
W, H, VD = 1280, 720, 12
pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
cam = vx_scenes.main_camera(W, H)
ctx = api.Context(0)
batch = api.BinaryGreedyMesher.mesh_batch(world.voxels, pos, world.neighbor_table(), world.uniform_flags, ctx, validate=False)
cfg = api.default_frame_config(W, H)
vp = cam.view_projection()
loop = api.FrameLoop(batch, cfg, view_distance=VD, want_depth=False, ctx=ctx)
for _ in range(3):
    loop.wait(loop.submit(vp, cam.position))
N = 300
# (a) steady state, two in flight
t0 = time.perf_counter(); prev = loop.submit(vp, cam.position)
ts, tw = 0.0, 0.0
for _ in range(1, N):
    a = time.perf_counter(); nxt = loop.submit(vp, cam.position); b = time.perf_counter(); loop.wait(prev); c = time.perf_counter()
    ts += b - a; tw += c - b; prev = nxt
loop.wait(prev)
el = time.perf_counter() - t0
print("pipelined: %.1f us/frame  (host: submit %.1f us, wait %.1f us)" % (el / N * 1e6, ts / N * 1e6, tw / N * 1e6))
# (b) D2H copy alone
d = torch.empty((H, W), dtype=torch.int32, device="cuda"); h = torch.empty((H, W), dtype=torch.int32).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("D2H 3.7 MB: %.1f us per copy" % ((time.perf_counter() - t0) / 100 * 1e6))
hm = torch.from_numpy(loop.color.view(np.int32))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): hm.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("D2H 3.7 MB into vx_host_alloc memory: %.1f us per copy" % ((time.perf_counter() - t0) / 100 * 1e6))
# (c) synchronous zero-copy
t0 = time.perf_counter()
for _ in range(N): loop.render(vp, cam.position)
print("synchronous zero-copy: %.1f us/frame" % ((time.perf_counter() - t0) / N * 1e6))
# (d) device-only async submit rate
cfga = api.VxFrameConfig.from_buffer_copy(cfg); cfga.async_submit = 1
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(N): api.render_frame_device(batch, vp, cam.position, cfga, VD, ctx)
t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
print("device frames async: host submit %.1f us/frame, total %.1f us/frame" % ((t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
# (e) begin/end without the frame copy (color_out = NULL): the period the kernels + small read-backs allow
import ctypes as C
lib, h = ctx.lib, ctx.handle
tk = C.c_int32(0); ns = C.c_int32(0)
vpp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16); camp = np.ascontiguousarray(cam.position, dtype=np.float32)
def begin(color):
    ctx.check(lib.vx_render_frame_begin(h, batch.handle, None, -1, vpp.ctypes.data, camp.ctypes.data, VD, C.byref(cfg), color, None, C.byref(tk)))
    return tk.value
def end(t):
    ctx.check(lib.vx_render_frame_end(h, t, None, C.byref(ns)))
for label, col in (("no frame copy", None), ("with frame copy", loop.color.ctypes.data)):
    end(begin(col)); end(begin(col))
    t0 = time.perf_counter(); prev = begin(col)
    for _ in range(1, N):
        nxt = begin(col); end(prev); prev = nxt
    end(prev)
    print("begin/end %s: %.1f us/frame" % (label, (time.perf_counter() - t0) / N * 1e6))
# (f) do kernels and D2H copies overlap at all?  N async frames on the context stream + N independent copies on another
side = torch.cuda.Stream()
ctx.synchronize(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N):
    api.render_frame_device(batch, vp, cam.position, cfga, VD, ctx)
    with torch.cuda.stream(side):
        hm.copy_(d, non_blocking=True)
ctx.synchronize(); torch.cuda.synchronize()
print("independent: frames + copies concurrently %.1f us/frame" % ((time.perf_counter() - t0) / N * 1e6))

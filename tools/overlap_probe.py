"""Frames in flight on the device (development probe): the same batch rendered through L independent contexts (stream + frame
scratch each), frames dealt round-robin, asynchronous submits.  Prints device throughput for L = 1, 2, 3.
   python tools/overlap_probe.py [W H VD]"""
import os, sys, time
import ctypes as C
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
ctxs = [api.Context(0) for _ in range(3)]
batch = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctxs[0])
cfg = api.default_frame_config(W, H)
cfga = api.VxFrameConfig.from_buffer_copy(cfg); cfga.async_submit = 1
vp = cam.view_projection()
for c in ctxs:
    api.render_frame_device(batch, vp, cam.position, cfg, VD, c)
    for _ in range(3):
        api.render_frame_device(batch, vp, cam.position, cfga, VD, c)
    c.synchronize()
ref = api.framebuffer_device(ctxs[0])
K = 600
def cfg_for(L):
    c = api.VxFrameConfig.from_buffer_copy(cfga)
    c.frames_in_flight = L
    return c
for L in (1, 2, 3):
    torch.cuda.synchronize()
    cl = cfg_for(L)
    t0 = time.perf_counter()
    for i in range(K):
        api.render_frame_device(batch, vp, cam.position, cl, VD, ctxs[i % L])
    t_sub = time.perf_counter() - t0
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"{W}x{H} vd{VD} lanes {L}: {t / K * 1e6:.1f} us per frame ({K / t:.0f} frames/s), host submit {t_sub / K * 1e6:.1f} us per frame")

# group mode: L frames start together right after an L2 flush (every frame of a group reads cold inputs); device time from the
# end of the flush to the last lane's end, per frame
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
streams = [torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", 0)) for c in ctxs]
for L in (1, 2, 3, 4):
    if L > len(ctxs):
        ctxs.append(api.Context(0))
        api.render_frame_device(batch, vp, cam.position, cfg, VD, ctxs[-1])
        streams.append(torch.cuda.ExternalStream(ctxs[-1].stream, device=torch.device("cuda", 0)))
    cl = cfg_for(L)
    for l in range(L):
        for _ in range(3):
            api.render_frame_device(batch, vp, cam.position, cl, VD, ctxs[l])
    torch.cuda.synchronize()
    G = 200 // L
    tot = 0.0
    evs = []
    for g in range(G):
        with torch.cuda.stream(streams[0]):
            flush.fill_(g & 1)
        f_ev = torch.cuda.Event(enable_timing=True); f_ev.record(streams[0])
        ends = []
        for l in range(L):
            if l:
                streams[l].wait_event(f_ev)
            api.render_frame_device(batch, vp, cam.position, cl, VD, ctxs[l])
            e = torch.cuda.Event(enable_timing=True); e.record(streams[l]); ends.append(e)
        for l in range(1, L):
            streams[0].wait_event(ends[l])  # the next flush starts when every lane is done
        evs.append((f_ev, ends))
    torch.cuda.synchronize()
    for f_ev, ends in evs:
        tot += max(f_ev.elapsed_time(e) for e in ends)
    cnt = (C.c_uint32 * 32)()
    ctxs[0].check(ctxs[0].lib.vx_frame_counters(ctxs[0].handle, cnt))
    print(f"group mode lanes {L}: {tot / (G * L) * 1e3:.1f} us per frame ({G * L / tot * 1e3:.0f} frames/s), group {tot / G * 1e3:.1f} us | tasks {cnt[14]} items {cnt[9]} entries {cnt[3]}")

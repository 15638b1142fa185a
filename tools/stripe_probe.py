"""Per-stripe device throughput with frames in flight (development probe): how long one GPU needs per frame for a stripe
[y0, y0 + rows) of the bench frame when it keeps L frames in flight -- the quantity the multi-GPU split has to balance.
   python tools/stripe_probe.py [L]"""
import os, sys, time
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vx_scenes
from differential_projection_voxel_renderer_b200 import api
W, H, VD = 1280, 720, 12
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
cam = vx_scenes.main_camera(W, H)
lanes = api.FrameLanes(0, L)
batch = api.BinaryGreedyMesher.mesh_batch(world.voxels, pos, world.neighbor_table(), world.uniform_flags, lanes[0], validate=False)
vp = cam.view_projection()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
streams = [torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", 0)) for c in lanes.ctxs]

def measure(y0, rows, n=180):
    cfg = api.default_frame_config(W, H)
    cfg.stripe_y0, cfg.stripe_rows = y0, rows
    for c in lanes.ctxs:
        api.render_frame_device(batch, vp, cam.position, cfg, VD, c)
    cfg.async_submit = 1
    cfg.frames_in_flight = L
    for c in lanes.ctxs:
        for _ in range(3):
            api.render_frame_device(batch, vp, cam.position, cfg, VD, c)
    torch.cuda.synchronize()
    evs = []
    for g in range(n // L):
        with torch.cuda.stream(streams[0]):
            flush.fill_(g & 1)
        f_ev = torch.cuda.Event(enable_timing=True); f_ev.record(streams[0])
        ends = []
        for l in range(L):
            if l:
                streams[l].wait_event(f_ev)
            api.render_frame_device(batch, vp, cam.position, cfg, VD, lanes[l])
            e = torch.cuda.Event(enable_timing=True); e.record(streams[l]); ends.append(e)
        for l in range(1, L):
            streams[0].wait_event(ends[l])
        evs.append((f_ev, ends))
    torch.cuda.synchronize()
    tot = sum(max(f.elapsed_time(e) for e in ends) for f, ends in evs)
    cnt = (C.c_uint32 * 32)()
    lanes[0].check(lanes[0].lib.vx_frame_counters(lanes[0].handle, cnt))
    return tot / (len(evs) * L) * 1e3, cnt[7], cnt[3], cnt[14]

print(f"lanes {L}")
for y0, rows in [(0, 720), (0, 392), (392, 328), (0, 360), (360, 360), (0, 424), (424, 296), (0, 456), (456, 264), (0, 180), (180, 180), (360, 180), (540, 180), (0, 96), (336, 96), (624, 96)]:
    us, units, entries, tasks = measure(y0, rows)
    print(f"stripe y0 {y0:4d} rows {rows:4d}: {us:6.1f} us per frame | setup units {units} bin entries {entries} tasks {tasks}")

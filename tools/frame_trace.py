"""Raster work-item timeline of one 1280x720 vd12 frame (diagnostics; not a benchmark):
   python tools/frame_trace.py [W H VD]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vx_scenes  # noqa: E402
from differential_projection_voxel_renderer_b200 import api  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
H = int(sys.argv[2]) if len(sys.argv) > 2 else 720
VD = int(sys.argv[3]) if len(sys.argv) > 3 else 12
pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
cam = vx_scenes.main_camera(W, H)
ctx = api.Context(0)
batch = api.BinaryGreedyMesher.mesh_batch(v, p, nb, None, ctx)
cfg = api.default_frame_config(W, H)
for _ in range(3):
    api.render_frame_device(batch, cam.view_projection(), cam.position, cfg, VD, ctx)
cfg.profile_kernels = 2
api.render_frame_device(batch, cam.view_projection(), cam.position, cfg, VD, ctx)
print("kernel ms", api.frame_kernel_times(ctx))
cap = 1 << 17
out = np.zeros((cap, 14), dtype=np.uint64)
n = C.c_int32()
ctx.check(ctx.lib.vx_frame_trace(ctx.handle, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
t = out[:n.value].astype(np.int64)
tile, part, parts = t[:, 0], t[:, 1] & 0xffff, t[:, 1] >> 16
t0, t1, sm, nsrc = t[:, 2], t[:, 3], t[:, 4], t[:, 5]
base = t0.min()
dur = (t1 - t0) / 1000.0
print(f"items {n.value}  span {(t1.max() - base) / 1000.0:.1f} us  first start {(t0.min() - base) / 1000:.1f}  last start {(t0.max() - base) / 1000:.1f} us")
print("item duration us: mean %.2f  median %.2f  p90 %.2f  max %.2f  sum %.0f" % (dur.mean(), np.median(dur), np.percentile(dur, 90), dur.max(), dur.sum()))
for lo, hi in ((0, 0), (1, 8), (9, 64), (65, 255), (256, 100000)):
    m = (nsrc >= lo) & (nsrc <= hi)
    if m.any():
        print(f"  n_src {lo:4d}..{hi:<6d}: {int(m.sum()):5d} items  mean {dur[m].mean():6.2f}  max {dur[m].max():6.2f} us")
ph = t[:, 6:9] - t[:, 2:3]
m = nsrc > 0
print("phase ends after item start (us, items with work): keys ready %.2f  warp 0 out of work %.2f  all warps done %.2f  end %.2f" % tuple(
    [float(np.mean(ph[m, k])) / 1000 for k in range(3)] + [float(dur[m].mean())]))
pp = t[:, 10:13]
rounds = t[:, 13]
mm2 = m & (pp[:, 0] > 0) & (t[:, 9] > 0)
if mm2.any():
    print("first round (single-part items): phase P starts %.2f  warp 0 done with P %.2f  second round starts %.2f us after item start; rounds mean %.2f max %d" % (
        float(np.mean(pp[mm2, 0] - t0[mm2])) / 1000, float(np.mean(pp[mm2, 1] - t0[mm2])) / 1000,
        float(np.mean(np.where(pp[mm2, 2] > 0, pp[mm2, 2] - t0[mm2], 0))) / 1000, float(rounds[mm2].mean()), int(rounds[mm2].max())))
wo = t[:, 9]
mm = m & (wo > 0)
if mm.any():
    print("single-part items: write-out starts %.2f us after item start, takes %.2f us (mean over %d items)" % (
        float(np.mean(wo[mm] - t0[mm])) / 1000, float(np.mean(t1[mm] - wo[mm])) / 1000, int(mm.sum())))
order = np.argsort(-dur)[:12]
for i in order[:6]:
    rel = lambda v: (v - t0[i]) / 1000.0 if v > 0 else float("nan")
    print(f"  slow item {i} phases (us after start): keys {rel(t[i, 6]):.2f} | R0 warp0 done {rel(t[i, 7]):.2f} | P0 starts {rel(pp[i, 0]):.2f} | P0 warp0 done {rel(pp[i, 1]):.2f} | round 2 starts {rel(pp[i, 2]):.2f} | "
          f"all rounds done {rel(t[i, 8]):.2f} | write-out starts {rel(t[i, 9]):.2f} | end {dur[i]:.2f} | rounds {rounds[i]}")
for i in order:
    print(f"  slow: item {i} tile ({tile[i] % ((W + 127) // 128)},{tile[i] // ((W + 127) // 128)}) part {part[i]}/{parts[i]} n_src {nsrc[i]} dur {dur[i]:.1f} us start {(t0[i] - base) / 1000:.1f} sm {sm[i]}")
late = np.argsort(-t1)[:8]
for i in late:
    print(f"  late: item {i} tile ({tile[i] % ((W + 127) // 128)},{tile[i] // ((W + 127) // 128)}) part {part[i]}/{parts[i]} n_src {nsrc[i]} dur {dur[i]:.1f} us end {(t1[i] - base) / 1000:.1f}")
so = np.zeros((148 * 12, 12), dtype=np.uint64)
ns = C.c_int32()
ctx.check(ctx.lib.vx_frame_setup_trace(ctx.handle, so.ctypes.data_as(C.c_void_p), so.shape[0], C.byref(ns)))
so = so[:ns.value].astype(np.int64)
so = so[so[:, 0] > 0]
b0 = so[:, 0].min()
rel = lambda c: (so[so[:, c] > 0, c] - b0) / 1000.0
print(f"setup: {so.shape[0]} working CTAs, units/CTA max {so[:, 7].max()}; start mean {rel(0).mean():.1f} max {rel(0).max():.1f} | ranked mean {rel(1).mean():.1f} max {rel(1).max():.1f} | "
      f"projected mean {rel(2).mean():.1f} max {rel(2).max():.1f} | counted mean {rel(8).mean():.1f} max {rel(8).max():.1f} | reserved mean {rel(9).mean():.1f} max {rel(9).max():.1f} | binned mean {rel(3).mean():.1f} max {rel(3).max():.1f} | done mean {rel(4).mean():.1f} max {rel(4).max():.1f}")
slow = np.argsort(-(so[:, 4] - so[:, 0]))[:8]
for i in slow:
    r = [(so[i, c] - b0) / 1000.0 if so[i, c] > 0 else float("nan") for c in (0, 1, 2, 8, 9, 3, 4)]
    print("  slow setup CTA: start %.1f ranked %.1f projected %.1f counted %.1f reserved %.1f binned %.1f done %.1f us | box tiles %d, triangles %d" % (*r, so[i, 5], so[i, 6]))
last = so[so[:, 10] > 0]
if last.shape[0]:
    print(f"setup plan: start {(last[0, 5] - b0) / 1000:.1f} loaded {(last[0, 10] - b0) / 1000:.1f} written {(last[0, 11] - b0) / 1000:.1f} end {(last[0, 6] - b0) / 1000:.1f} us;  raster first item starts {(base - b0) / 1000:.1f} us after the first setup CTA")
cnt = np.zeros(32, dtype=np.uint32)
ctx.check(ctx.lib.vx_frame_counters(ctx.handle, cnt.ctypes.data_as(C.c_void_p)))
print("control block: big triangles", int(cnt[6]), "setup units", int(cnt[7]), "items per class (K=0 class last)", cnt[16:25].tolist(), "second clip pieces", int(cnt[13]))
st = api.frame_stats(ctx)
print("survivors", st.n_survivors, "tris", st.n_triangles, "entries", st.n_bin_entries, "max_bin", st.reserved[0], "items", st.reserved[1])
batch.release()
ctx.close()

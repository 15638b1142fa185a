"""Workload statistics of a frame from the CPU oracle built with -DVXO_STATS (triangles, rows, spans, fragments,
depth passes, span-length histogram).  Test/analysis tooling only.  usage: python tools/frame_stats_cpu.py [W H VD]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
so = os.path.join(ROOT, "oracle", "build", "libvx_oracle_stats.so")
subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-DVXO_STATS", "-shared",
                       "-o", so, os.path.join(ROOT, "oracle", "vx_oracle.c"), "-lm", "-lpthread"])
from oracle import binding as ob  # noqa: E402

ob._LIB_PATH = so
ob.build = lambda force=False: so
import vx_scenes  # noqa: E402

W, H, VD = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1280, 720, 12)
pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
cam = vx_scenes.main_camera(W, H)
ref = ob.mesh_chunks(v, nb, None, p)
vp = cam.view_projection()
vis = ob.cull_chunks(p, vp, cam.position, VD)
ids = np.flatnonzero((vis != 0) & (ref.has_mesh != 0)).astype(np.int32)
cfg = ob.default_frame_config(W, H, n_threads=1)
stats = (C.c_uint64 * 64).in_dll(ob.lib(), "vxo_stats")
for i in range(64):
    stats[i] = 0
color, depth, surv = ob.render_frame(ref, ids, vp, cam.position, cfg, ob.default_atlas())
s = list(stats)
print(f"{W}x{H} vd{VD}: candidates {ids.size} survivors {surv.size} quads {sum(int(ref.quad_count[i]) for i in surv) if hasattr(ref,'quad_count') else '?'}")
print(f"triangles {s[0]} rows {s[1]} spans {s[2]} fragments {s[3]} depth passes {s[4]} covered px {(color != cfg.clear_color).sum()}")
for lg in range(16):
    if s[8 + lg]:
        print(f"  span len [{1 << lg:5d},{(2 << lg) - 1:5d}]: {s[8 + lg]:8d} spans {s[24 + lg]:9d} px")

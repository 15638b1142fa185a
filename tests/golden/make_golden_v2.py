"""Regenerates tests/golden/golden_v2.json from the CPU oracle: regression fixtures for the restatements added after
golden_v1 (macrotile renderer, occlusion pass, barycentric mesh path, span walker, Hyper-Pipeline composition).
       python tests/golden/make_golden_v2.py
Like golden_v1 these are oracle outputs that guard the oracle against drift, not outputs of the Rust reference."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vx_scenes  # noqa: E402
from oracle import binding as ob  # noqa: E402

VD, W, H = 3, 320, 180
CAMERAS = (0, 1, 3, 5)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def span_walker_case():
    rng = np.random.default_rng(20261019)
    n = 200
    x0 = rng.uniform(-1.2, 1.0, n).astype(np.float32); x1 = x0 + rng.uniform(0.0, 0.5, n).astype(np.float32)
    y0 = rng.uniform(-1.2, 1.0, n).astype(np.float32); y1 = y0 + rng.uniform(0.0, 0.5, n).astype(np.float32)
    z = rng.choice(np.linspace(0.1, 0.9, 9).astype(np.float32), n)
    bt = rng.integers(0, 5, n).astype(np.uint8)
    vis = (rng.random(n) > 0.15).astype(np.uint8)
    return x0, y0, x1, y1, z, bt, vis


def build():
    out = {"version": 2, "generator": "tests/golden/make_golden_v2.py (CPU oracle)", "vd": VD, "width": W, "height": H}
    pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
    ref = ob.mesh_chunks(v, nb, None, p)
    atlas = ob.default_atlas()
    out["frames"] = {}
    for ci in CAMERAS:
        cam = vx_scenes.path_camera(ci, W, H)
        vp = cam.view_projection()
        visv = ob.cull_chunks(p, vp, cam.position, VD)
        ids = np.flatnonzero((visv != 0) & (ref.has_mesh != 0)).astype(np.int32)
        e = {}
        cfg = ob.default_frame_config(W, H, n_threads=2)
        c, d, proj, kind = ob.render_frame_macrotile(ref, ids, vp, cfg, atlas, want_kinds=True)
        e["macrotile"] = {"color_sha": sha(c), "tile_depth_sha": sha(d), "projected": proj.astype(int).tolist(), "kinds": kind.astype(int).tolist()}
        cfg.occlusion_culling = 1
        c, d, s = ob.render_frame(ref, ids, vp, cam.position, cfg, atlas)
        e["occlusion"] = {"color_sha": sha(c), "depth_sha": sha(d), "order": s.astype(int).tolist()}
        c = np.full((H, W), 0xFF87CEEB, dtype=np.uint32); d = np.full((H, W), np.inf, dtype=np.float32)
        plain = ob.default_frame_config(W, H)
        for m in ids.tolist():
            ob.render_mesh_tiny_quads(ref, m, vp, plain, atlas, (0, 0, W, H), False, c, d)
        e["barycentric"] = {"color_sha": sha(c), "depth_sha": sha(d), "covered": int((c != 0xFF87CEEB).sum())}
        c = np.zeros((H, W), dtype=np.uint32); d = np.full((H, W), np.inf, dtype=np.float32)
        packets, quads = ob.hyper_pipeline_render(ref, ids, vp, c, d)
        e["hyper_pipeline"] = {"color_sha": sha(c), "depth_sha": sha(d), "packets": packets, "quads": quads}
        out["frames"][str(ci)] = e
    x0, y0, x1, y1, z, bt, vis = span_walker_case()
    c = np.zeros((120, 160), dtype=np.uint32); d = np.full((120, 160), np.inf, dtype=np.float32)
    ob.span_walk_quads(c, d, x0, y0, x1, y1, z, bt, vis)
    out["span_walker"] = {"color_sha": sha(c), "depth_sha": sha(d), "covered": int((d != np.inf).sum())}
    return out


if __name__ == "__main__":
    g = build()
    path = os.path.join(HERE, "golden_v2.json")
    json.dump(g, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")

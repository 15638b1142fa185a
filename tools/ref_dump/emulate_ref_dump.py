"""Writes a file in the format of ref_dump.rs from the CPU ORACLE -- only to exercise the loader in
tests/test_reference_golden.py (VX_REF_VECTORS=/tmp/emulated.json python -m pytest tests/test_reference_golden.py).
Its output is NOT a reference vector and must never be committed as tests/golden/ref_vectors.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import binding as ob  # noqa: E402
from differential_projection_voxel_renderer_b200 import camera, worldgen  # noqa: E402
from test_reference_golden import fnv1a, mesh_rows  # noqa: E402

positions = [(0, 0, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1), (0, 1, 0), (3, 0, -5), (-12, 0, 11), (2, -2, 2),
             (0, 0, -2), (1, 0, -2), (-1, 0, -2), (0, -1, -2)]
pos = np.asarray(positions, dtype=np.int32)
order = np.lexsort((pos[:, 2], pos[:, 1], pos[:, 0]))
w = worldgen.generate_world(pos[order])
back = np.empty_like(order)
back[order] = np.arange(order.size)
ref = ob.mesh_chunks(w.voxels, w.neighbor_table(), w.uniform_flags, w.positions)


def dump(rows):
    out = []
    for f in range(6):
        for s in range(32):
            q = [[int(x) for x in r[2:]] for r in rows if r[0] == f and r[1] == s]
            if q:
                out.append([f, s, q])
    return out


V = {"format": 1, "crate": "EMULATED from the oracle -- not a reference vector", "heights": [], "chunks": [], "frames": []}
for cx, cz in ((0, 0), (3, -5), (-12, 11)):
    V["heights"].append({"chunk_xz": [cx, cz], "top": ob.terrain_heights(cx * 32, cz * 32, 32, 32).tolist()})
for i, p in enumerate(positions):
    j = int(back[i])
    uni = int(w.uniform_flags[j])
    rows = mesh_rows(ref.slice_offsets[j], ref.chunk_quads(j), ob.unpack_quads) if ref.has_mesh[j] else np.zeros((0, 7), np.int64)
    alone = ob.mesh_chunks(w.voxels[j:j + 1]) if uni == 0 else None
    rows_a = mesh_rows(alone.slice_offsets[0], alone.chunk_quads(0), ob.unpack_quads) if alone is not None and alone.has_mesh[0] else None
    V["chunks"].append({"pos": list(p), "uniform": uni, "voxels_fnv": fnv1a(w.voxels[j].tobytes()), "quads_in_world": dump(rows) if ref.has_mesh[j] else None,
                        "quad_count_in_world": int(ref.quad_count[j]), "quads_alone": dump(rows_a) if rows_a is not None else None})
cam = camera.Camera((0.0, 10.0, 20.0), 1280 / 720)
V["camera"] = {"position": [0.0, 10.0, 20.0], "aspect": "1280/720", "vp_bits": cam.view_projection().reshape(16).view(np.uint32).tolist()}
for wd, ht, lst in ((1280, 720, [0]), (640, 360, [0, 2, 3, 4, 5, 10, 11, 12]), (320, 180, [12, 11, 10, 5, 0])):
    c2 = camera.Camera((0.0, 10.0, 20.0), wd / ht)
    vp = c2.view_projection().reshape(16)
    cfg = ob.default_frame_config(wd, ht)
    color = np.full((ht, wd), cfg.clear_color, dtype=np.uint32)
    depth = np.full((ht, wd), np.inf, dtype=np.float32)
    drawn = [i for i in lst if ref.has_mesh[int(back[i])]]
    for i in drawn:
        ob.render_mesh(ref, int(back[i]), vp, cfg, ob.default_atlas(), (0, 0, wd, ht), color, depth)
    y0 = ht * 2 // 3
    V["frames"].append({"width": wd, "height": ht, "meshes": drawn, "vp_bits": vp.view(np.uint32).tolist(), "covered": int((color != cfg.clear_color).sum()),
                        "color_fnv": fnv1a(color.tobytes()), "depth_fnv": fnv1a(depth.tobytes()), "row0": y0,
                        "color_rows": color[y0:y0 + 4].tolist(), "depth_bits_rows": depth[y0:y0 + 4].view(np.uint32).tolist()})
json.dump(V, open(sys.argv[1], "w"))
print("wrote", sys.argv[1])

"""Regenerates tests/golden/golden_v1.json from the CPU oracle (oracle/vx_oracle.c):
       python tests/golden/make_golden.py
The reference crate ships no golden vectors and cannot be built in this image, so these fixtures are
oracle outputs: they pin the oracle AND the CUDA path against drift (tests/test_golden.py); they are not outputs of the
Rust reference."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vx_kat as kat  # noqa: E402
import vx_scenes  # noqa: E402
from oracle import binding as ob  # noqa: E402

VD, W, H = 3, 320, 180
CAMERAS = (0, 1, 5)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def slice_masks():
    rng = np.random.default_rng(20261018)
    masks = {
        "full": np.full(32, 0xFFFFFFFF, dtype=np.uint32),
        "checker_rows": np.array([0xFFFFFFFF if r % 2 == 0 else 0 for r in range(32)], dtype=np.uint32),
        "sparse": np.full(32, 0x80000001, dtype=np.uint32),
        "staircase": np.array([(1 << (r + 1)) - 1 if r < 31 else 0xFFFFFFFF for r in range(32)], dtype=np.uint32),
        "random": rng.integers(0, 1 << 32, size=32, dtype=np.uint64).astype(np.uint32),
    }
    return masks


def build():
    out = {"version": 1, "generator": "tests/golden/make_golden.py (CPU oracle)", "vd": VD, "width": W, "height": H}
    out["slices"] = {}
    for name, m in slice_masks().items():
        q = ob.greedy_mesh_slice(m)
        out["slices"][name] = {"mask": [int(x) for x in m], "quads": q.reshape(-1, 4).astype(int).tolist()}
    pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
    ref = ob.mesh_chunks(v, nb, None, p)
    out["world"] = {"chunks": int(p.shape[0]), "total_quads": int(ref.quad_count.sum()),
                    "quad_count_sha": sha(ref.quad_count.astype(np.uint32)),
                    "slice_offsets_sha": sha(ref.slice_offsets.astype(np.uint32)),
                    "face_aabb_sha": sha(ref.face_aabb.astype(np.int32)),
                    "chunk_quads_sha": [sha(ref.chunk_quads(i)) for i in range(p.shape[0])]}
    shapes = {"slab": kat.chunk_slab(), "checker3d": kat.chunk_checker3d()}
    out["shapes"] = {}
    for name, vox in shapes.items():
        m = ob.mesh_chunks(vox.reshape(1, -1))
        out["shapes"][name] = {"quads": int(m.quad_count[0]), "sha": sha(m.chunk_quads(0))}
    out["frames"] = {}
    for ci in CAMERAS:
        cam = vx_scenes.path_camera(ci, W, H)
        vp = cam.view_projection()
        vis = ob.cull_chunks(pos, vp, cam.position, VD)
        visv = ob.cull_chunks(p, vp, cam.position, VD)
        ids = np.flatnonzero((visv != 0) & (ref.has_mesh != 0)).astype(np.int32)
        c, d, s = ob.render_frame(ref, ids, vp, cam.position, ob.default_frame_config(W, H, n_threads=2), ob.default_atlas())
        out["frames"][str(ci)] = {"vp_bits": [int(x) for x in np.asarray(vp, dtype=np.float32).reshape(16).view(np.uint32)],
                                  "cam": [float(x) for x in cam.position],
                                  "vp_sha": sha(vp.astype(np.float32)), "visible_sha": sha(vis.astype(np.uint8)),
                                  "visible": int(vis.sum()), "order": s.astype(int).tolist(), "color_sha": sha(c), "depth_sha": sha(d),
                                  "covered": int((c != 0xFF87CEEB).sum())}
    return out


if __name__ == "__main__":
    g = build()
    path = os.path.join(HERE, "golden_v1.json")
    json.dump(g, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")

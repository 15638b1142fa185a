"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- see vx_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "libvx_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "vx_oracle.c"), os.path.join(_HERE, "vx_oracle.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


class Atlas(C.Structure):
    _fields_ = [("palette", (C.c_uint32 * 16) * 4), ("indices", (C.c_uint8 * 32) * 4)]


class FrameConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("clear_color", C.c_uint32),
                ("backface_culling", C.c_int32), ("enable_shading", C.c_int32),
                ("light_dir", C.c_float * 3), ("ambient", C.c_float), ("diffuse", C.c_float),
                ("n_threads", C.c_int32), ("occlusion_culling", C.c_int32), ("occlusion_grid_w", C.c_int32),
                ("occlusion_grid_h", C.c_int32)]


class MeshBatchView(C.Structure):
    _fields_ = [("quads", C.c_void_p), ("quad_base", C.c_void_p), ("slice_offsets", C.c_void_p),
                ("face_aabb", C.c_void_p), ("positions", C.c_void_p), ("has_mesh", C.c_void_p),
                ("n_chunks", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.vxo_mesh_chunks.restype = C.c_int64
        _lib.vxo_shade_color_u32.restype = C.c_uint32
        _lib.vxo_shade_color_u32.argtypes = [C.c_uint32, C.c_float]
        _lib.vxo_face_light.restype = C.c_float
        _lib.vxo_texture_sample.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def greedy_mesh_slice(mask) -> np.ndarray:
    mask = np.ascontiguousarray(mask, dtype=np.uint32)
    assert mask.shape == (32,)
    out = np.zeros((512, 4), dtype=np.uint8)
    n = lib().vxo_greedy_mesh_slice(_p(mask), _p(out))
    return out[:n].copy()


def tinyquad_pack(u, v, w, h, bt) -> np.ndarray:
    out = np.zeros(3, dtype=np.uint8)
    lib().vxo_tinyquad_pack(C.c_uint8(u), C.c_uint8(v), C.c_uint8(w), C.c_uint8(h), C.c_uint8(bt), _p(out))
    return out


def unpack_quads(q3: np.ndarray) -> np.ndarray:
    """(n,3) u8 TinyQuads -> (n,5) [u, v, w, h, block_type]  (mesh.rs:309-341)."""
    q3 = np.asarray(q3, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    u = q3[:, 0] & 0x1F
    v = ((q3[:, 0] >> 5) & 7) | ((q3[:, 1] & 3) << 3)
    w = ((q3[:, 1] >> 2) & 0x3F) + 1
    h = (q3[:, 2] & 0x3F) + 1
    bt = (q3[:, 2] >> 6) & 3
    return np.stack([u, v, w, h, bt], axis=1)


class MeshBatch:
    """Host-side mesh batch in the same layout the product ABI returns."""

    def __init__(self, quads, quad_base, quad_count, slice_offsets, face_aabb, has_mesh, positions):
        self.quads = quads
        self.quad_base = quad_base
        self.quad_count = quad_count
        self.slice_offsets = slice_offsets
        self.face_aabb = face_aabb
        self.has_mesh = has_mesh
        self.positions = np.ascontiguousarray(positions, dtype=np.int32)

    @property
    def n_chunks(self):
        return int(self.quad_base.shape[0])

    def chunk_quads(self, i) -> np.ndarray:
        b, n = int(self.quad_base[i]), int(self.quad_count[i])
        return self.quads[3 * b:3 * (b + n)].reshape(-1, 3)

    def view(self) -> MeshBatchView:
        return MeshBatchView(_p(self.quads), _p(self.quad_base), _p(self.slice_offsets), _p(self.face_aabb),
                             _p(self.positions), _p(self.has_mesh), self.n_chunks)


def mesh_chunks(voxels, neighbors=None, uniform_flags=None, positions=None, cap_quads=None) -> MeshBatch:
    voxels = np.ascontiguousarray(voxels, dtype=np.uint8).reshape(-1, 32768)
    n = voxels.shape[0]
    if neighbors is not None:
        neighbors = np.ascontiguousarray(neighbors, dtype=np.int32).reshape(n, 6)
    if uniform_flags is not None:
        uniform_flags = np.ascontiguousarray(uniform_flags, dtype=np.uint8).reshape(n)
    if positions is None:
        positions = np.zeros((n, 3), dtype=np.int32)
    cap = int(cap_quads) if cap_quads is not None else max(4096, n * 2048)
    while True:
        quads = np.zeros(cap * 3, dtype=np.uint8)
        quad_base = np.zeros(n, dtype=np.uint32)
        quad_count = np.zeros(n, dtype=np.uint32)
        so = np.zeros((n, 6, 33), dtype=np.uint32)
        ab = np.zeros((n, 6, 6), dtype=np.int32)
        hm = np.zeros(n, dtype=np.uint8)
        tot = lib().vxo_mesh_chunks(_p(voxels), _p(neighbors), _p(uniform_flags), C.c_int32(n), _p(quads),
                                    C.c_int64(cap), _p(quad_base), _p(quad_count), _p(so), _p(ab), _p(hm))
        if tot >= 0:
            break
        cap *= 4
    return MeshBatch(quads[:3 * tot].copy(), quad_base, quad_count, so, ab, hm, positions)


def noise_permutation_table(seed: int = 12345) -> np.ndarray:
    """noise 0.9.0 PermutationTable::new(seed).values (256,) u8."""
    out = np.zeros(256, dtype=np.uint8)
    lib().vxo_noise_permutation_table(C.c_uint32(seed), _p(out))
    return out


def perlin2(x: float, y: float, seed: int = 12345) -> float:
    L = lib()
    L.vxo_perlin2.restype = C.c_double
    L.vxo_perlin2.argtypes = [C.c_void_p, C.c_double, C.c_double]
    return float(L.vxo_perlin2(_p(noise_permutation_table(seed)), float(x), float(y)))


def terrain_heights(x0: int, z0: int, nx: int, nz: int, seed: int = 12345) -> np.ndarray:
    """heights[z, x] = sample_terrain_height (chunk.rs:173-177) for world columns x0.., z0.."""
    L = lib()
    L.vxo_terrain_height.restype = C.c_int32
    L.vxo_terrain_height.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    perm = noise_permutation_table(seed)
    pp = _p(perm)
    out = np.zeros((nz, nx), dtype=np.int32)
    for z in range(nz):
        for x in range(nx):
            out[z, x] = L.vxo_terrain_height(pp, x0 + x, z0 + z)
    return out


def generate_terrain(position, seed: int = 12345):
    """Chunk::generate_terrain (chunk.rs:114-170) -> (uniform flag 0 / 1 / 4, voxels (32768,) u8 -- zeros when Uniform)."""
    pos = np.ascontiguousarray(position, dtype=np.int32).reshape(3)
    vox = np.zeros(32768, dtype=np.uint8)
    flag = int(lib().vxo_generate_terrain(_p(pos), C.c_uint32(seed), _p(vox)))
    return flag, vox


def frustum_from_vp(vp) -> np.ndarray:
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    planes = np.zeros((6, 4), dtype=np.float32)
    lib().vxo_frustum_from_vp(_p(vp), _p(planes))
    return planes


def frustum_intersects_aabb(planes, mn, mx) -> bool:
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    mn = np.ascontiguousarray(mn, dtype=np.float32)
    mx = np.ascontiguousarray(mx, dtype=np.float32)
    return bool(lib().vxo_frustum_intersects_aabb(_p(planes), _p(mn), _p(mx)))


def cull_chunks(positions, vp, cam_pos, view_distance, frustum_culling=True) -> np.ndarray:
    positions = np.ascontiguousarray(positions, dtype=np.int32).reshape(-1, 3)
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    cam = np.ascontiguousarray(cam_pos, dtype=np.float32).reshape(3)
    out = np.zeros(positions.shape[0], dtype=np.uint8)
    lib().vxo_cull_chunks(_p(positions), C.c_int32(positions.shape[0]), _p(vp), _p(cam), C.c_int32(view_distance),
                          C.c_int32(1 if frustum_culling else 0), _p(out))
    return out


def horizon_cull(cam_pos, centers, order=None, bins=128, base_margin=0.1, margin_dist_factor=0.05,
                 min_dist_chunks=2.0) -> np.ndarray:
    centers = np.ascontiguousarray(centers, dtype=np.float32).reshape(-1, 3)
    n = centers.shape[0]
    order = np.arange(n, dtype=np.int32) if order is None else np.ascontiguousarray(order, dtype=np.int32).copy()
    cam = np.ascontiguousarray(cam_pos, dtype=np.float32).reshape(3)
    k = lib().vxo_horizon_cull(_p(cam), _p(centers), C.c_int32(order.shape[0]), _p(order), C.c_int32(bins),
                               C.c_float(base_margin), C.c_float(margin_dist_factor), C.c_float(min_dist_chunks))
    return order[:k].copy()


def default_atlas() -> Atlas:
    a = Atlas()
    lib().vxo_default_atlas(C.byref(a))
    return a


def default_frame_config(w, h, n_threads=1) -> FrameConfig:
    cfg = FrameConfig()
    lib().vxo_default_frame_config(C.byref(cfg), C.c_int(w), C.c_int(h))
    cfg.n_threads = n_threads
    return cfg


def render_mesh(mb: MeshBatch, mesh_id, vp, cfg: FrameConfig, atlas: Atlas, rect, color, depth):
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    rect = np.ascontiguousarray(rect, dtype=np.int32).reshape(4)
    v = mb.view()
    lib().vxo_render_mesh(C.byref(v), C.c_int32(mesh_id), _p(vp), C.byref(cfg), C.byref(atlas), _p(rect),
                          _p(color), _p(depth))


def render_mesh_tiny_quads(mb: MeshBatch, mesh_id, vp, cfg: FrameConfig, atlas: Atlas, rect, use_span_renderer, color, depth):
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    rect = np.ascontiguousarray(rect, dtype=np.int32).reshape(4)
    v = mb.view()
    lib().vxo_render_mesh_tiny_quads(C.byref(v), C.c_int32(mesh_id), _p(vp), C.byref(cfg), C.byref(atlas), _p(rect),
                                     C.c_int32(1 if use_span_renderer else 0), _p(color), _p(depth))


def render_mesh_with_up(mb: MeshBatch, mesh_id, vp, cfg: FrameConfig, atlas: Atlas, camera_up, color, depth):
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    up = np.ascontiguousarray(camera_up, dtype=np.float32).reshape(3)
    v = mb.view()
    lib().vxo_render_mesh_with_up(C.byref(v), C.c_int32(mesh_id), _p(vp), C.byref(cfg), C.byref(atlas), _p(up), _p(color), _p(depth))


def render_frame(mb: MeshBatch, mesh_ids, vp, cam_pos, cfg: FrameConfig, atlas: Atlas):
    """Returns (color (H,W) u32, depth (H,W) f32, survivors (k,) i32 in draw order)."""
    mesh_ids = np.ascontiguousarray(mesh_ids, dtype=np.int32)
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    cam = np.ascontiguousarray(cam_pos, dtype=np.float32).reshape(3)
    color = np.zeros((cfg.height, cfg.width), dtype=np.uint32)
    depth = np.zeros((cfg.height, cfg.width), dtype=np.float32)
    surv = np.zeros(max(1, mesh_ids.shape[0]), dtype=np.int32)
    v = mb.view()
    k = lib().vxo_render_frame(C.byref(v), _p(mesh_ids), C.c_int32(mesh_ids.shape[0]), _p(vp), _p(cam),
                               C.byref(cfg), C.byref(atlas), _p(color), _p(depth), _p(surv))
    return color, depth, surv[:k].copy()


def face_packets(mb: "MeshBatch", mesh_id: int):
    """ChunkFacePackets::from_chunk_mesh + FacePacketBuilder (face_packets.rs:72-174), restated with plain loops:
    6 lists of packets, each a dict of len + the six u8[32] arrays (unused lanes zero, FacePacket32::new :28-38)."""
    so = mb.slice_offsets[mesh_id]
    quads = unpack_quads(mb.chunk_quads(mesh_id))
    faces = []
    for f in range(6):
        packets, cur = [], None
        for s in range(32):
            axis_pos = s + 1 if f % 2 == 0 else s  # :146-151
            for q in range(int(so[f, s]), int(so[f, s + 1])):
                if cur is None or cur["len"] >= 32:  # FacePacketBuilder::push :93-99
                    if cur is not None:
                        packets.append(cur)
                    cur = {"len": 0, **{k: np.zeros(32, dtype=np.uint8) for k in ("u_min", "v_min", "u_len", "v_len", "axis_pos", "block_type")}}
                u, v, w, h, bt = (int(x) for x in quads[q])
                i = cur["len"]
                cur["u_min"][i], cur["v_min"][i], cur["u_len"][i], cur["v_len"][i] = u, v, w, h
                cur["axis_pos"][i], cur["block_type"][i] = axis_pos, bt
                cur["len"] = i + 1
        if cur is not None and cur["len"] > 0:  # finish :102-107
            packets.append(cur)
        faces.append(packets)
    return faces


def face_basis(face, chunk_pos, slice_idx, vp) -> np.ndarray:
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    cp = np.ascontiguousarray(chunk_pos, dtype=np.int32).reshape(3)
    out = np.zeros((4, 4), dtype=np.float32)
    lib().vxo_face_basis(C.c_int(face), _p(cp), C.c_uint8(slice_idx), _p(vp), _p(out))
    return out


def basis_project_point(basis, u, v) -> np.ndarray:
    basis = np.ascontiguousarray(basis, dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    lib().vxo_basis_project_point(_p(basis), C.c_float(u), C.c_float(v), _p(out))
    return out


def project_packet(basis, u_min, v_min, u_len, v_len):
    basis = np.ascontiguousarray(basis, dtype=np.float32)
    arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in (u_min, v_min, u_len, v_len)]
    n = arrs[0].shape[0]
    outs = [np.zeros(n, dtype=np.float32) for _ in range(5)]
    lib().vxo_project_packet(_p(basis), *[_p(a) for a in arrs], C.c_int(n), *[_p(o) for o in outs])
    return outs


def transform_vertices(verts8, offset, vp) -> np.ndarray:
    verts8 = np.ascontiguousarray(verts8, dtype=np.uint8).reshape(-1, 8)
    off = np.ascontiguousarray(offset, dtype=np.float32).reshape(3)
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    out = np.zeros((verts8.shape[0], 4), dtype=np.float32)
    lib().vxo_transform_vertices(_p(verts8), C.c_int32(verts8.shape[0]), _p(off), _p(vp), _p(out))
    return out


def quad_clip_vertices(face, slice_pos, u, v, w, h, chunk_pos, vp) -> np.ndarray:
    cp = np.ascontiguousarray(chunk_pos, dtype=np.int32).reshape(3)
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    out = np.zeros((4, 4), dtype=np.float32)
    lib().vxo_quad_clip_vertices(C.c_int(face), C.c_uint8(slice_pos), C.c_uint8(u), C.c_uint8(v), C.c_uint8(w),
                                 C.c_uint8(h), _p(cp), _p(vp), _p(out))
    return out


# ---- adjacent rasterizers -----------------------------------------------------------------

def span_walker_block_color(block_type: int) -> int:
    f = lib().vxo_span_walker_block_color
    f.restype = C.c_uint32
    return int(f(C.c_uint8(block_type)))


def fill_span(color, depth, y, x_start, x_end, d, c):
    """FrameSlice::fill_span on (H, W) colour / depth arrays, in place."""
    lib().vxo_fill_span(C.c_int32(color.shape[1]), C.c_int32(y), C.c_int32(x_start), C.c_int32(x_end), C.c_float(d),
                        C.c_uint32(c), _p(color), _p(depth))


def span_walk_quads(color, depth, x_min, y_min, x_max, y_max, depth_near, block_type, visible=None):
    """SpanWalkerRasterizer::rasterize_projected_packet over n projected quads submitted as consecutive
    ProjectedPackets of up to 32 (the layout PacketPipeline produces), in place on (H, W) colour / depth arrays."""
    arrs = [np.ascontiguousarray(a, dtype=np.float32).ravel() for a in (x_min, y_min, x_max, y_max, depth_near)]
    bt = np.ascontiguousarray(block_type, dtype=np.uint8).ravel()
    n = bt.size
    vis = np.ones(n, dtype=np.uint8) if visible is None else np.ascontiguousarray(visible, dtype=np.uint8).ravel()
    h, w = color.shape
    for p0 in range(0, n, 32):
        cnt = min(32, n - p0)
        mask = 0
        for i in range(cnt):
            if vis[p0 + i]:
                mask |= 1 << i
        sl = [np.ascontiguousarray(a[p0:p0 + cnt]) for a in arrs]
        b = np.ascontiguousarray(bt[p0:p0 + cnt])
        lib().vxo_span_walk_packet(_p(sl[0]), _p(sl[1]), _p(sl[2]), _p(sl[3]), _p(sl[4]), _p(b), C.c_uint32(mask),
                                   C.c_int32(cnt), C.c_int32(w), C.c_int32(h), _p(color), _p(depth))


def macrotile_bin(min_x, min_y, max_x, max_y, fb_w, fb_h):
    """MacroTileBins::add_mesh: (kind, (tx0, ty0, tx1, ty1)); kind 1 binned, 2 large primitive, 0 off-screen."""
    t = np.zeros(4, dtype=np.int32)
    k = lib().vxo_macrotile_bin(C.c_int32(min_x), C.c_int32(min_y), C.c_int32(max_x), C.c_int32(max_y), C.c_int32(fb_w),
                                C.c_int32(fb_h), _p(t))
    return int(k), tuple(int(x) for x in t)


def render_frame_macrotile(mb: MeshBatch, mesh_ids, vp, cfg: FrameConfig, atlas: Atlas, want_kinds: bool = False):
    """render_frame_macrotile: returns (color (H,W) u32, tile depth (H,W) f32 [diagnostic], projected mesh ids in list
    order[, kind per projected mesh: 1 binned, 2 large primitive])."""
    mesh_ids = np.ascontiguousarray(mesh_ids, dtype=np.int32)
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    color = np.zeros((cfg.height, cfg.width), dtype=np.uint32)
    depth = np.zeros((cfg.height, cfg.width), dtype=np.float32)
    proj = np.zeros(max(1, mesh_ids.shape[0]), dtype=np.int32)
    kind = np.zeros(max(1, mesh_ids.shape[0]), dtype=np.int32)
    v = mb.view()
    k = lib().vxo_render_frame_macrotile(C.byref(v), _p(mesh_ids), C.c_int32(mesh_ids.shape[0]), _p(vp), C.byref(cfg),
                                         C.byref(atlas), _p(color), _p(depth), _p(proj), _p(kind))
    if want_kinds:
        return color, depth, proj[:k].copy(), kind[:k].copy()
    return color, depth, proj[:k].copy()


def hyper_pipeline_render(mb: MeshBatch, mesh_ids, vp, color, depth):
    """The reference's Hyper-Pipeline for a list of meshes, composed from the restated pieces exactly as
    tests/span_walker_fuzz_tests.rs:158-173 / benches/differential_projection.rs do: ChunkFacePackets::from_chunk_mesh
    (face_packets.rs:122-174) -> PacketPipeline::process_chunk_packets (packet_pipeline.rs:69-142: one FaceBasis per
    packet taken from the packet's FIRST quad, packet-level backface test normal.z < 0, scalar projection,
    frustum mask :279-293) -> SpanWalkerRasterizer::rasterize_projected_packet per surviving packet.
    Returns (packets drawn, quads visible)."""
    vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
    h, w = color.shape
    n_packets = n_quads = 0
    for mid in np.asarray(mesh_ids, dtype=np.int64).tolist():
        if not mb.has_mesh[mid]:
            continue
        faces = face_packets(mb, mid)
        for f in range(6):
            for pk in faces[f]:
                n = int(pk["len"])
                if n == 0:
                    continue
                basis = face_basis(f, mb.positions[mid], int(pk["axis_pos"][0]), vp)
                if not (basis[3, 2] < np.float32(0.0)):  # is_front_facing differential_projection.rs:78-82
                    continue
                x0, y0, x1, y1, z = project_packet(basis, pk["u_min"][:n], pk["v_min"][:n], pk["u_len"][:n], pk["v_len"][:n])
                one, zero = np.float32(1.0), np.float32(0.0)
                vis = (x1 >= -one) & (x0 <= one) & (y1 >= -one) & (y0 <= one) & (z >= zero) & (z <= one)  # test_aabb_inside
                if not vis.any():
                    continue
                mask = 0
                for i in range(n):
                    if vis[i]:
                        mask |= 1 << i
                bt = np.ascontiguousarray(pk["block_type"][:n])
                lib().vxo_span_walk_packet(_p(x0), _p(y0), _p(x1), _p(y1), _p(z), _p(bt), C.c_uint32(mask), C.c_int32(n),
                                           C.c_int32(w), C.c_int32(h), _p(color), _p(depth))
                n_packets += 1
                n_quads += int(vis.sum())
    return n_packets, n_quads

"""Literal restatement of the reference's frame-loop bookkeeping with HOST chunks and the oracle mesher (test
infrastructure): World::update (world.rs:57-100), visible chunks (world.rs:118-146) and the mesh cache
(main.rs:224-280).  The streaming tests compare the device-resident World / MeshCache against it frame by frame."""
import numpy as np

from differential_projection_voxel_renderer_b200 import worldgen

OFFSETS = ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1))


class RefLoop:
    def __init__(self, ob, view_distance, max_chunks_per_frame):
        self.ob = ob
        self.vd = view_distance
        self.cap = max_chunks_per_frame
        self.chunks = {}       # pos -> (uniform flag, voxels (32768,) u8)
        self.mesh_cache = {}   # pos -> (quads (n,3) u8, slice_offsets (6,33), face_aabb (6,6)) or None
        self.generated = []
        self.meshed = []

    @staticmethod
    def camera_chunk(cam):
        c = np.floor(np.asarray(cam, dtype=np.float32) / np.float32(32))
        return int(c[0]), int(c[1]), int(c[2])

    def update(self, cam):
        cc = self.camera_chunk(cam)
        vd = self.vd
        self.generated = []
        for cx in range(cc[0] - vd, cc[0] + vd + 1):
            for cy in range(cc[1] - vd, cc[1] + vd + 1):
                for cz in range(cc[2] - vd, cc[2] + vd + 1):
                    d = (cx - cc[0]) ** 2 + (cy - cc[1]) ** 2 + (cz - cc[2]) ** 2
                    if np.float32(d) > np.float32(vd * vd):
                        continue
                    if (cx, cy, cz) not in self.chunks:
                        w = worldgen.generate_world(np.array([[cx, cy, cz]], dtype=np.int32), store_uniform_voxels=True)
                        self.chunks[(cx, cy, cz)] = (int(w.uniform_flags[0]), w.voxels[0].copy())
                        self.generated.append((cx, cy, cz))
                        if len(self.generated) >= self.cap:
                            return True
        ud = vd + 2
        self.chunks = {p: c for p, c in self.chunks.items()
                       if np.float32((p[0] - cc[0]) ** 2 + (p[1] - cc[1]) ** 2 + (p[2] - cc[2]) ** 2) <= np.float32(ud * ud)}
        return len(self.generated) > 0

    def visible(self, cam, vp):
        allp = sorted(self.chunks)
        if not allp:
            return []
        vis = self.ob.cull_chunks(np.array(allp, dtype=np.int32), vp, cam, self.vd, True)
        return [p for p, v in zip(allp, vis.tolist()) if v]

    def mesh_world(self):
        """Oracle meshes of every chunk of the current world (index = sorted position order)."""
        allp = sorted(self.chunks)
        index = {p: i for i, p in enumerate(allp)}
        vox = np.stack([self.chunks[p][1] for p in allp])
        flags = np.array([self.chunks[p][0] for p in allp], dtype=np.uint8)
        nb = np.full((len(allp), 6), -1, dtype=np.int32)
        for i, p in enumerate(allp):
            for f, o in enumerate(OFFSETS):
                nb[i, f] = index.get((p[0] + o[0], p[1] + o[1], p[2] + o[2]), -1)
        mb = self.ob.mesh_chunks(vox, nb, flags, np.array(allp, dtype=np.int32))
        return index, mb

    def update_cache(self, visible):
        to_mesh = []
        for pos in visible:
            if pos not in self.mesh_cache:
                to_mesh.append(pos)
                for o in OFFSETS:
                    npos = (pos[0] + o[0], pos[1] + o[1], pos[2] + o[2])
                    if npos in self.chunks and npos in self.mesh_cache:
                        to_mesh.append(npos)
        to_mesh = sorted(set(to_mesh))
        if to_mesh:
            index, mb = self.mesh_world()
            for pos in to_mesh:
                i = index.get(pos)
                if i is None:
                    continue
                self.mesh_cache[pos] = (mb.chunk_quads(i).copy(), mb.slice_offsets[i].copy(), mb.face_aabb[i].copy()) if mb.has_mesh[i] else None
        self.mesh_cache = {p: m for p, m in self.mesh_cache.items() if p in self.chunks}
        self.meshed = to_mesh
        return to_mesh

    def frame(self, cam, vp):
        self.update(cam)
        vis = self.visible(cam, vp)
        self.update_cache(vis)
        return vis

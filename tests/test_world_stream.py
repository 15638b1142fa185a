"""Host logic of the streaming world (differential_projection_voxel_renderer_b200/world.py: World::update world.rs:57-100,
mesh cache main.rs:224-280) against a literal restatement of the reference loop, with the device calls replaced by a
recorder that keeps the same state a world batch would (slots, neighbour rows, per-slot meshes from the oracle).  CPU."""
import numpy as np
import pytest

import vx_refloop
from differential_projection_voxel_renderer_b200 import camera, world as vxw, worldgen


class FakeDevice:
    """State of a world batch: per slot position / flag / voxels / neighbour row / mesh, meshes built by the oracle."""

    def __init__(self, ob, capacity):
        self.ob = ob
        self.capacity = capacity
        self.pos = {}
        self.flag = {}
        self.vox = {}
        self.nbr = {s: [-1] * 6 for s in range(capacity)}
        self.mesh = {}   # slot -> quads or None (has_mesh = 0)
        self.calls = []
        self.ctx = None
        self.batch = None

    def generate(self, slots, positions):
        w = worldgen.generate_world(positions, store_uniform_voxels=True)
        for i, s in enumerate(slots.tolist()):
            assert s not in self.pos, "slot reused while occupied"
            self.pos[s] = tuple(int(v) for v in positions[i])
            self.flag[s] = int(w.uniform_flags[i])
            self.vox[s] = w.voxels[i].copy()
        self.calls.append(("generate", slots.tolist()))
        return w.uniform_flags.copy()

    def assign(self, slots, neighbors):
        for s, row in zip(slots.tolist(), neighbors.tolist()):
            assert s in self.pos
            self.nbr[s] = list(row)
        self.calls.append(("assign", slots.tolist()))

    def grow(self, capacity):
        assert capacity > self.capacity
        for s in range(self.capacity, capacity):
            self.nbr[s] = [-1] * 6
        self.capacity = capacity
        self.calls.append(("grow", capacity))

    def unload(self, slots):
        for s in slots.tolist():
            del self.pos[s], self.flag[s], self.vox[s]
            self.nbr[s] = [-1] * 6
            self.mesh.pop(s, None)
        self.calls.append(("unload", slots.tolist()))

    def remesh(self, slots):
        occupied = sorted(self.pos)
        idx = {s: i for i, s in enumerate(occupied)}
        vox = np.stack([self.vox[s] for s in occupied])
        flags = np.array([self.flag[s] for s in occupied], dtype=np.uint8)
        nb = np.array([[idx.get(n, -1) if n >= 0 else -1 for n in self.nbr[s]] for s in occupied], dtype=np.int32)
        for s in occupied:  # a neighbour row may only reference occupied slots
            assert all(n < 0 or n in self.pos for n in self.nbr[s]), "stale neighbour reference"
        mb = self.ob.mesh_chunks(vox, nb, flags)
        for s in slots.tolist():
            i = idx[s]
            self.mesh[s] = mb.chunk_quads(i).copy() if mb.has_mesh[i] else None
        self.calls.append(("remesh", slots.tolist()))


CAMERA_WALK = [(0.0, 10.0, 20.0), (0.0, 10.0, 20.0), (10.0, 12.0, 5.0), (40.0, 14.0, -20.0), (75.0, 20.0, -40.0), (75.0, 20.0, -40.0),
               (140.0, 30.0, -40.0), (140.0, 30.0, -40.0), (140.0, 30.0, -40.0), (20.0, 5.0, 0.0), (20.0, 5.0, 0.0)]


@pytest.mark.parametrize("vd,cap", [(2, 4), (2, 1000), (3, 16)])
def test_streaming_world_follows_the_reference_loop(ob, vd, cap):
    ref = vx_refloop.RefLoop(ob, vd, cap)
    dev = FakeDevice(ob, vxw.sphere_capacity(vd))
    w = vxw.World(vxw.WorldConfig(view_distance=vd, max_chunks_per_frame=cap), device=dev)
    cache = vxw.MeshCache(w)
    meshed_total = 0
    for step, pos in enumerate(CAMERA_WALK * 2):
        cam = camera.Camera(pos, 16 / 9, yaw=0.3 * step)
        vp = cam.view_projection()
        r_changed = ref.update(cam.position)
        w_changed = w.update(cam.position)
        assert r_changed == w_changed
        assert w.generated_last_update == ref.generated, f"step {step}: generation order differs"
        assert sorted(w.chunks) == sorted(ref.chunks)
        assert {p: w.uniform_flags[p] for p in w.chunks} == {p: c[0] for p, c in ref.chunks.items()}
        # slots are unique, inside the capacity, and neighbour rows describe the present world
        assert len(set(w.chunks.values())) == len(w.chunks) and all(0 <= s < dev.capacity for s in w.chunks.values())
        for p, s in w.chunks.items():
            want = [w.chunks.get((p[0] + o[0], p[1] + o[1], p[2] + o[2]), -1) for o in vxw.FACE_OFFSETS]
            assert dev.nbr[s] == want and dev.pos[s] == p
        vis = ref.visible(cam.position, vp)
        assert w.get_visible_chunks(cam.position) == [p for p in sorted(ref.chunks)
                                                      if (p[0] - ref.camera_chunk(cam.position)[0]) ** 2 + (p[1] - ref.camera_chunk(cam.position)[1]) ** 2
                                                      + (p[2] - ref.camera_chunk(cam.position)[2]) ** 2 <= vd * vd]
        r_list = ref.update_cache(vis)
        w_list = cache.update(vis)
        assert w_list == r_list, f"step {step}: chunks to (re)mesh differ"
        assert cache.cached == set(ref.mesh_cache)
        meshed_total += len(w_list)
        for p, m in ref.mesh_cache.items():  # every cached mesh, stale ones included, equals the reference's
            got = dev.mesh[w.chunks[p]]
            if m is None:
                assert got is None
            else:
                assert got is not None and np.array_equal(got, m[0]), f"step {step}: mesh of {p} differs"
        ids = cache.visible_mesh_slots(vis)
        assert ids.tolist() == [w.chunks[p] for p in sorted(vis) if p in ref.mesh_cache]
    assert meshed_total > 0 and any(c[0] == "unload" for c in dev.calls)


def test_world_update_cap_returns_before_unloading(ob):
    """world.rs:84-87: hitting max_chunks_per_frame returns before the unload step."""
    dev = FakeDevice(ob, vxw.sphere_capacity(1))
    w = vxw.World(vxw.WorldConfig(view_distance=1, max_chunks_per_frame=1000), device=dev)
    assert w.update((0.0, 0.0, 0.0)) and w.chunk_count() == 7
    w.config.max_chunks_per_frame = 2
    assert w.update((32.0 * 10, 0.0, 0.0))          # far away: two new chunks, the old seven are NOT unloaded yet
    assert w.chunk_count() == 9 and w.unloaded_last_update == []
    for _ in range(2):
        assert w.update((32.0 * 10, 0.0, 0.0))
    assert w.chunk_count() == 13
    assert w.update((32.0 * 10, 0.0, 0.0))          # the seventh chunk: below the cap, so the unload step runs
    assert w.chunk_count() == 7 and len(w.unloaded_last_update) == 7
    assert not w.update((32.0 * 10, 0.0, 0.0))      # nothing left to do
    assert vxw.world_to_chunk_pos((-0.5, 31.9, 32.0)) == (-1, 0, 1)  # world.rs:201-207
    # a batch that is too small grows (the reference's HashMap has no capacity): nothing is lost, slots stay unique
    dev3 = FakeDevice(ob, 3)
    small = vxw.World(vxw.WorldConfig(view_distance=1, max_chunks_per_frame=1000), device=dev3, capacity=3)
    small.update((0.0, 0.0, 0.0))
    assert small.chunk_count() == 7 and small.capacity >= 7 and any(c[0] == "grow" for c in dev3.calls)
    assert sorted(small.chunks.values()) == sorted(set(small.chunks.values())) and max(small.chunks.values()) < small.capacity


def test_moving_camera_never_unloads_and_the_world_grows_like_the_reference(ob):
    """A camera that keeps moving hits max_chunks_per_frame on every update, so the reference returns before its unload
    step (world.rs:84-87) and its chunk map just grows -- far beyond the view sphere.  The port follows it chunk for chunk
    (same loaded set every frame) by growing the device batch; set_view_distance (world.rs:181-184, main.rs:168-176)
    takes effect at the next update."""
    vd, cap = 2, 4
    ref = vx_refloop.RefLoop(ob, vd, cap)
    dev = FakeDevice(ob, vxw.sphere_capacity(vd))
    w = vxw.World(vxw.WorldConfig(view_distance=vd, max_chunks_per_frame=cap), device=dev)
    first_capacity = w.capacity
    for step in range(160):
        pos = (16.0 * step, 10.0, 20.0)  # half a chunk per frame
        assert ref.update(pos) == w.update(pos)
        assert sorted(w.chunks) == sorted(ref.chunks), f"step {step}"
        assert w.unloaded_last_update == []
    assert w.chunk_count() > first_capacity and w.capacity > first_capacity and any(c[0] == "grow" for c in dev.calls)
    assert len(set(w.chunks.values())) == len(w.chunks) and max(w.chunks.values()) < w.capacity
    for p, s in w.chunks.items():  # neighbour rows survived the growth
        assert dev.nbr[s] == [w.chunks.get((p[0] + o[0], p[1] + o[1], p[2] + o[2]), -1) for o in vxw.FACE_OFFSETS]
    # standing still lets the unload step run; widening the view distance loads the bigger sphere
    for _ in range(40):
        ref.update(pos)
        w.update(pos)
    assert sorted(w.chunks) == sorted(ref.chunks) and w.chunk_count() <= vxw.sphere_capacity(vd)
    w.set_view_distance(3)
    ref.vd = 3
    for _ in range(60):
        ref.update(pos)
        w.update(pos)
    assert w.view_distance() == 3 and sorted(w.chunks) == sorted(ref.chunks) and w.chunk_count() > vxw.sphere_capacity(2) // 2
    w.set_view_distance(0)
    assert w.view_distance() == 1  # .max(1)

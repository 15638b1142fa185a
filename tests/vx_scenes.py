"""Seeded scenes shared by the parity tests, the golden-fixture generator and bench.py."""
import functools

import numpy as np

from differential_projection_voxel_renderer_b200 import camera, worldgen


@functools.lru_cache(maxsize=8)
def terrain_scene(view_distance: int):
    """All lattice chunks within view_distance of chunk (0,0,0) (world.rs:57-100), Varied ones compacted.
    Returns (positions_all, world, positions_v, voxels_v, neighbors_v)."""
    pos = worldgen.lattice_sphere((0, 0, 0), view_distance)
    world = worldgen.generate_world(pos)
    p, v, nb = world.compact_varied()
    return pos, world, p, v, nb


def main_camera(width, height):
    """main.rs:51: Camera::new((0,10,20), aspect), looking down -Z."""
    return camera.Camera((0.0, 10.0, 20.0), width / height)


# horizon_culling_pipeline_movement_tests.rs:218-224 style camera path (positions / yaw / pitch)
CAMERA_PATH = [
    ((0.0, 10.0, 20.0), 0.0, 0.0),
    ((40.0, 25.0, -30.0), 0.7, -0.25),
    ((-60.0, 18.0, 10.0), 2.4, -0.1),
    ((5.0, 60.0, 5.0), 1.1, -1.2),
    ((100.0, 12.0, 100.0), -2.0, 0.05),
    ((3.3, 2.2, 7.7), 0.3, 0.4),      # low, near/inside terrain: exercises near-plane clipping
]


def path_camera(i, width, height):
    p, yaw, pitch = CAMERA_PATH[i]
    return camera.Camera(p, width / height, yaw=yaw, pitch=pitch)

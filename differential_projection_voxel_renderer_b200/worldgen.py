"""Synthetic world generator (INPUT DATA for tests and bench.py).

Restates the data shape of the reference's `Chunk::generate_terrain`
(/root/reference/src/voxel/chunk.rs:114-177): a 2-D gradient-noise heightfield
`h = trunc(noise(x*0.01, z*0.01) * 20)`, Grass at y == h, Dirt for h-3 < y < h,
Stone below, Air above, with the uniform shortcuts of chunk.rs:127-134
(chunk entirely above max height -> Uniform(Air); chunk top < min height - 10
-> Uniform(Stone)).

The reference samples `noise 0.9.0` `Perlin::new(12345)`, a crates.io
dependency that is not vendored under /root/reference.  Its published
algorithm (XorShift-shuffled permutation table, perlin_2d with four diagonal
gradients, quintic fade, bilinear blend, 2/sqrt(2) scale, clamp) is restated
in the CPU restatement under oracle/; this module and the CUDA generator reproduce that
restatement bit for bit (tests pin all three against each other).  No
reference test pins a height value, so until tools/ref_dump vectors exist the
heights are still parity-unpinned against the crate itself.
"""
from __future__ import annotations

import numpy as np

CHUNK_SIZE = 32
CHUNK_VOLUME = CHUNK_SIZE ** 3
AIR, GRASS, DIRT, STONE = 0, 1, 2, 3

NBR_NONE = -1
NBR_UNIFORM_AIR = -2
NBR_UNIFORM_SOLID = -3
# +X, -X, +Y, -Y, +Z, -Z  (FaceDir order, /root/reference/src/meshing/mesh.rs:136-143)
FACE_OFFSETS = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.int32)


def _perm_table(seed: int) -> np.ndarray:
    """`noise 0.9.0` PermutationTable::new(seed): XorShiftRng seeded with the bytes [1,0,0,0, seed, seed, seed] (LE),
    then (0..256).shuffle(rng) as rand 0.8.5 does it (Fisher-Yates from the back, UniformInt<u32>::sample_single with the
    widening-multiply / zone rejection).  Returned doubled (512 entries) for the device table."""
    m32 = 0xFFFFFFFF
    x, y, z, w = 1, seed & m32, seed & m32, seed & m32
    if (x | y | z | w) == 0:
        x = y = z = w = 0x0BAD5EED
    vals = list(range(256))
    for i in range(255, 0, -1):
        rng_range = i + 1
        lz = 32 - rng_range.bit_length()
        zone = ((rng_range << lz) & m32) - 1
        while True:
            t = (x ^ (x << 11)) & m32
            x, y, z = y, z, w
            w = (w ^ (w >> 19) ^ (t ^ (t >> 8))) & m32
            m = w * rng_range
            if (m & m32) <= zone:
                pick = m >> 32
                break
        vals[i], vals[pick] = vals[pick], vals[i]
    p = np.asarray(vals, dtype=np.int64)
    return np.concatenate([p, p])


# gradients selected by hash & 3 (core/perlin.rs): +x+y, -x+y, +x-y, -x-y; the table is 8 long for the device struct
_GRAD2 = np.array([[1, 1], [-1, 1], [1, -1], [-1, -1]] * 2, dtype=np.float64)
_SCALE_2D = 2.0 / 1.4142135623730951  # 2 / core::f64::consts::SQRT_2


def perlin2(x: np.ndarray, y: np.ndarray, seed: int = 12345) -> np.ndarray:
    """`noise 0.9.0` perlin_2d over PermutationTable::new(seed), f64, operation by operation as restated in
    the CPU restatement under oracle/; output clamped to [-1, 1]."""
    perm = _perm_table(seed)[:256]
    cx = np.floor(x).astype(np.int64)
    cy = np.floor(y).astype(np.int64)
    dx = x - cx
    dy = y - cy

    def quintic(t):
        t = np.clip(t, 0.0, 1.0)
        return t * t * t * (t * (t * 6.0 - 15.0) + 10.0)

    def grad(ox, oy):
        qx = dx - float(ox)
        qy = dy - float(oy)
        h = perm[perm[(cx + ox) & 255] ^ ((cy + oy) & 255)] & 3
        return np.where(h == 0, qx + qy, np.where(h == 1, -qx + qy, np.where(h == 2, qx - qy, -qx - qy)))

    g00, g10, g01, g11 = grad(0, 0), grad(1, 0), grad(0, 1), grad(1, 1)
    u = quintic(dx)
    v = quintic(dy)
    k0, k1, k2, k3 = g00, g10 - g00, g01 - g00, g00 + g11 - g10 - g01
    unscaled = k0 + k1 * u + k2 * v + k3 * u * v
    return np.clip(unscaled * _SCALE_2D, -1.0, 1.0)


def terrain_heights(x0: int, z0: int, nx: int, nz: int, seed: int = 12345) -> np.ndarray:
    """heights[z, x] for world columns x0..x0+nx, z0..z0+nz (chunk.rs:173-177)."""
    xs = (np.arange(x0, x0 + nx, dtype=np.float64) * 0.01)[None, :]
    zs = (np.arange(z0, z0 + nz, dtype=np.float64) * 0.01)[:, None]
    n = perlin2(np.broadcast_to(xs, (nz, nx)), np.broadcast_to(zs, (nz, nx)), seed)
    return np.trunc(n * 20.0).astype(np.int32)


def chunk_voxels_from_heights(h: np.ndarray, world_y0: int) -> np.ndarray:
    """h[z, x] (32x32) -> voxels[z, y, x] u8 (index = z*1024 + y*32 + x, chunk.rs:52,139-165)."""
    wy = (world_y0 + np.arange(CHUNK_SIZE, dtype=np.int32))[None, :, None]
    hh = h[:, None, :]
    v = np.full((CHUNK_SIZE, CHUNK_SIZE, CHUNK_SIZE), STONE, dtype=np.uint8)
    v[wy > hh - 3] = DIRT
    v[wy == hh] = GRASS
    v[wy > hh] = AIR
    return v


class World:
    """A set of chunks: positions (N,3) i32 sorted by (x,y,z), uniform_flags (N,) u8
    (0 = Varied, else 1 + block type), and voxels for the Varied ones.

    `voxels` is (N, 32768) u8; rows of Uniform chunks are filled with their block
    type (they are never read by the mesher for uniform_flags != 0)."""

    def __init__(self, positions, uniform_flags, voxels):
        self.positions = np.ascontiguousarray(positions, dtype=np.int32)
        self.uniform_flags = np.ascontiguousarray(uniform_flags, dtype=np.uint8)
        self.voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
        self.index = {tuple(p): i for i, p in enumerate(self.positions.tolist())}

    @property
    def n_chunks(self) -> int:
        return int(self.positions.shape[0])

    def neighbor_table(self) -> np.ndarray:
        """(N,6) i32: index of the neighbour chunk in this world or NBR_NONE
        (build_neighbors_indexed, /root/reference/src/meshing/binary_greedy.rs:195-209)."""
        nb = np.full((self.n_chunks, 6), NBR_NONE, dtype=np.int32)
        for i, p in enumerate(self.positions.tolist()):
            for f in range(6):
                q = (p[0] + int(FACE_OFFSETS[f, 0]), p[1] + int(FACE_OFFSETS[f, 1]), p[2] + int(FACE_OFFSETS[f, 2]))
                j = self.index.get(q)
                if j is not None:
                    nb[i, f] = j
        return nb

    def compact_varied(self):
        """Drop the voxel rows of Uniform chunks: returns (positions_v, voxels_v, neighbors_v)
        where neighbours index the compacted array or carry NBR_UNIFORM_* codes."""
        varied = np.flatnonzero(self.uniform_flags == 0)
        remap = np.full(self.n_chunks, -1, dtype=np.int64)
        remap[varied] = np.arange(varied.size)
        nb_full = self.neighbor_table()[varied]
        nb = np.full_like(nb_full, NBR_NONE)
        has = nb_full >= 0
        tgt = np.where(has, nb_full, 0)
        flags = self.uniform_flags[tgt]
        nb[has & (flags == 0)] = remap[tgt[has & (flags == 0)]].astype(np.int32)
        nb[has & (flags == 1)] = NBR_UNIFORM_AIR
        nb[has & (flags > 1)] = NBR_UNIFORM_SOLID
        return self.positions[varied].copy(), self.voxels[varied].copy(), nb


def lattice_sphere(center_chunk, view_distance: int) -> np.ndarray:
    """All chunk coordinates with |c - center|^2 <= vd^2, sorted by (x,y,z)
    (streaming predicate of /root/reference/src/world.rs:57-100,130-133)."""
    r = view_distance
    g = np.arange(-r, r + 1, dtype=np.int32)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    keep = (x * x + y * y + z * z) <= r * r
    pts = np.stack([x[keep], y[keep], z[keep]], axis=1) + np.asarray(center_chunk, dtype=np.int32)[None, :]
    order = np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))
    return np.ascontiguousarray(pts[order])


def generate_world(positions: np.ndarray, seed: int = 12345, store_uniform_voxels: bool = False) -> World:
    """Chunk::generate_terrain for every position (chunk.rs:114-170)."""
    positions = np.ascontiguousarray(positions, dtype=np.int32)
    n = positions.shape[0]
    flags = np.zeros(n, dtype=np.uint8)
    vox = np.zeros((n, CHUNK_VOLUME), dtype=np.uint8)
    cache = {}
    for i in range(n):
        cx, cy, cz = (int(v) for v in positions[i])
        key = (cx, cz)
        h = cache.get(key)
        if h is None:
            h = terrain_heights(cx * CHUNK_SIZE, cz * CHUNK_SIZE, CHUNK_SIZE, CHUNK_SIZE, seed)
            cache[key] = h
        mn, mx = int(h.min()), int(h.max())
        y0 = cy * CHUNK_SIZE
        if y0 > mx:  # all air above terrain (chunk.rs:127-129)
            flags[i] = 1 + AIR
        elif y0 + CHUNK_SIZE < mn - 10:  # all solid below (chunk.rs:132-134)
            flags[i] = 1 + STONE
            if store_uniform_voxels:
                vox[i, :] = STONE
        else:
            vox[i] = chunk_voxels_from_heights(h, y0).reshape(-1)
    return World(positions, flags, vox)


def terrain_chunk(cx: int, cy: int, cz: int, seed: int = 12345) -> np.ndarray:
    """Voxels (32768,) u8 of one terrain chunk, always Varied-style full data."""
    h = terrain_heights(cx * CHUNK_SIZE, cz * CHUNK_SIZE, CHUNK_SIZE, CHUNK_SIZE, seed)
    return chunk_voxels_from_heights(h, cy * CHUNK_SIZE).reshape(-1)


def random_chunk(rng: np.random.Generator, density: float = 0.5, types: int = 3) -> np.ndarray:
    """Adversarial noise chunk: each voxel solid with probability `density`."""
    solid = rng.random(CHUNK_VOLUME) < density
    t = rng.integers(1, types + 1, size=CHUNK_VOLUME, dtype=np.uint8)
    return np.where(solid, t, 0).astype(np.uint8)

/*
 * vx_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see vx_oracle.h).
 *
 * Plain-C restatement of the reference crate's CPU path.  Every function
 * cites the reference file:line it follows.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math   (no FMA contraction, IEEE ops)
 * so that f32 arithmetic is evaluated exactly as rustc evaluates the Rust
 * source (one rounding per operation, source order).
 *
 * glam 0.25.0 (not vendored) semantics restated here:
 *   Mat4 * Vec4 = ((c0*x + c1*y) + c2*z) + c3*w     (SSE2 backend, unfused)
 *   Vec4 / f32  = per-component IEEE division
 */
#include "vx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>

#define CS VXO_CHUNK_SIZE

/* ---------- small helpers --------------------------------------------- */

/* Rust `f32 as i32`: saturating, NaN -> 0. */
static inline int32_t f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
/* Rust f32::min / f32::max ignore a NaN operand, like C fminf/fmaxf. */
static inline float rmin(float a, float b) { return fminf(a, b); }
static inline float rmax(float a, float b) { return fmaxf(a, b); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static inline int is_solid(uint8_t t) { return t != 0; } /* block_type.rs:16-21 */

/* glam Mat4::mul_vec4 (column-major m[col*4+row]) */
static inline void mat4_mul_vec4(const float m[16], float x, float y, float z, float w, float out[4]) {
    for (int r = 0; r < 4; ++r) {
        float a = m[0 + r] * x;
        a = a + m[4 + r] * y;
        a = a + m[8 + r] * z;
        a = a + m[12 + r] * w;
        out[r] = a;
    }
}

/* ---------- meshing ---------------------------------------------------- */

static inline int ctz32(uint32_t v) { return v ? __builtin_ctz(v) : 32; }

/* binary_greedy.rs:683-807.  The reference unrolls the row-extension loop by
 * four (:721-778); that block is semantically the simple loop below (a failed
 * group falls through to the scalar tail :781-791 which re-tests row by row). */
int vxo_greedy_mesh_slice(const uint32_t mask[32], uint8_t *q) {
    uint32_t data[CS];
    memcpy(data, mask, sizeof(data));
    int n = 0;
    for (int row = 0; row < CS; ++row) {
        if (data[row] == 0) continue;
        uint32_t col = 0;
        while (col < CS) {
            col += (uint32_t)ctz32(data[row] >> col);
            if (col >= CS) break;
            uint32_t height = (uint32_t)ctz32(~(data[row] >> col)); /* trailing_ones */
            uint32_t hmask = height >= 32 ? ~0u : ((1u << height) - 1u);
            uint32_t m = hmask << col;
            uint32_t width = 1;
            uint32_t max_width = (uint32_t)(CS - row);
            while (width < max_width) {
                uint32_t next = (data[row + width] >> col) & hmask;
                if (next != hmask) break;
                data[row + width] &= ~m;
                width += 1;
            }
            q[4 * n + 0] = (uint8_t)row;
            q[4 * n + 1] = (uint8_t)col;
            q[4 * n + 2] = (uint8_t)width;
            q[4 * n + 3] = (uint8_t)height;
            n++;
            data[row] &= ~m;
            col += height;
        }
    }
    return n;
}

/* mesh.rs:283-307 */
void vxo_tinyquad_pack(uint8_t u, uint8_t v, uint8_t w, uint8_t h, uint8_t bt, uint8_t out[3]) {
    uint8_t wp = (uint8_t)(w - 1), hp = (uint8_t)(h - 1);
    out[0] = (uint8_t)((u & 0x1F) | ((v & 0x07) << 5));
    out[1] = (uint8_t)(((v >> 3) & 0x03) | ((wp & 0x3F) << 2));
    out[2] = (uint8_t)((hp & 0x3F) | ((bt & 0x03) << 6));
}
/* mesh.rs:309-341 */
void vxo_tinyquad_unpack(const uint8_t in[3], uint8_t *u, uint8_t *v, uint8_t *w, uint8_t *h, uint8_t *bt) {
    *u = in[0] & 0x1F;
    *v = (uint8_t)(((in[0] >> 5) & 0x07) | ((in[1] & 0x03) << 3));
    *w = (uint8_t)(((in[1] >> 2) & 0x3F) + 1);
    *h = (uint8_t)((in[2] & 0x3F) + 1);
    *bt = (in[2] >> 6) & 0x03;
}

/* binary_greedy.rs:446-458 */
static inline void slice_to_chunk_coords(int axis, int slice, int row, int col, int *x, int *y, int *z) {
    if (axis == 0) { *x = slice; *y = row; *z = col; }
    else if (axis == 1) { *x = row; *y = slice; *z = col; }
    else { *x = row; *y = col; *z = slice; }
}
static inline int vidx(int x, int y, int z) { return z * CS * CS + y * CS + x; }

/* binary_greedy.rs:463-513 */
static int has_solid_neighbor_pos(const uint8_t *c, const uint8_t *nb, int nb_uniform_solid, int x, int y, int z, int axis) {
    switch (axis) {
    case 0:
        if (x + 1 < CS) return is_solid(c[vidx(x + 1, y, z)]);
        if (nb) return is_solid(nb[vidx(0, y, z)]);
        return nb_uniform_solid;
    case 1:
        if (y + 1 < CS) return is_solid(c[vidx(x, y + 1, z)]);
        if (nb) return is_solid(nb[vidx(x, 0, z)]);
        return nb_uniform_solid;
    default:
        if (z + 1 < CS) return is_solid(c[vidx(x, y, z + 1)]);
        if (nb) return is_solid(nb[vidx(x, y, 0)]);
        return nb_uniform_solid;
    }
}
/* binary_greedy.rs:518-570 */
static int has_solid_neighbor_neg(const uint8_t *c, const uint8_t *nb, int nb_uniform_solid, int x, int y, int z, int axis) {
    switch (axis) {
    case 0:
        if (x > 0) return is_solid(c[vidx(x - 1, y, z)]);
        if (nb) return is_solid(nb[vidx(CS - 1, y, z)]);
        return nb_uniform_solid;
    case 1:
        if (y > 0) return is_solid(c[vidx(x, y - 1, z)]);
        if (nb) return is_solid(nb[vidx(x, CS - 1, z)]);
        return nb_uniform_solid;
    default:
        if (z > 0) return is_solid(c[vidx(x, y, z - 1)]);
        if (nb) return is_solid(nb[vidx(x, y, CS - 1)]);
        return nb_uniform_solid;
    }
}

/* binary_greedy.rs:286-440 (Varied fast path; the axis-2 branch :364-409 uses
 * the same (row=x, col=y) mapping, only the loop nest differs). */
static void generate_binary_masks(const uint8_t *c, const uint8_t *nb, int nb_uniform_solid, int face, int slice,
                                  uint32_t masks[4][CS], int used[4]) {
    memset(masks, 0, sizeof(uint32_t) * 4 * CS);
    used[0] = used[1] = used[2] = used[3] = 0;
    int axis = face >> 1, positive = (face & 1) == 0;
    for (int row = 0; row < CS; ++row)
        for (int col = 0; col < CS; ++col) {
            int x, y, z;
            slice_to_chunk_coords(axis, slice, row, col, &x, &y, &z);
            uint8_t cur = c[vidx(x, y, z)];
            if (!is_solid(cur)) continue;
            int hn = positive ? has_solid_neighbor_pos(c, nb, nb_uniform_solid, x, y, z, axis)
                              : has_solid_neighbor_neg(c, nb, nb_uniform_solid, x, y, z, axis);
            if (!hn) {
                masks[cur & 3][row] |= 1u << col;
                used[cur & 3] = 1;
            }
        }
}

/* mesh_chunk_in_world binary_greedy.rs:83-121 + mesh_face :213-264 +
 * ChunkMesh::add_quad mesh.rs:489-523 + FaceList::add_quad mesh.rs:369-397. */
int vxo_mesh_chunk(const uint8_t *voxels, const uint8_t *const nbr_voxels[6], const int32_t nbr_code[6],
                   uint8_t *quads_out, int cap, uint32_t slice_offsets[6 * 33], int32_t face_aabb[6 * 6]) {
    int n = 0;
    uint32_t masks[4][CS];
    int used[4];
    uint8_t q[512 * 4];
    for (int face = 0; face < 6; ++face) { /* +X,-X,+Y,-Y,+Z,-Z  :105-112 */
        int axis = face >> 1, positive = (face & 1) == 0;
        int32_t *mn = face_aabb + face * 6, *mx = mn + 3;
        mn[0] = mn[1] = mn[2] = 32; /* FaceList::new mesh.rs:359-365 */
        mx[0] = mx[1] = mx[2] = 0;
        const uint8_t *nb = nbr_voxels ? nbr_voxels[face] : NULL;
        int nb_solid = (!nb && nbr_code && nbr_code[face] == VXO_NBR_UNIFORM_SOLID) ? 1 : 0;
        for (int slice = 0; slice < CS; ++slice) {
            slice_offsets[face * 33 + slice] = (uint32_t)n;
            generate_binary_masks(voxels, nb, nb_solid, face, slice, masks, used);
            for (int t = 0; t < 4; ++t) { /* BLOCK_TYPES order :239 */
                if (!used[t]) continue;
                int nq = vxo_greedy_mesh_slice(masks[t], q);
                int axis_pos = positive ? slice + 1 : slice; /* :251-255 */
                for (int i = 0; i < nq; ++i) {
                    if (n >= cap) return -1;
                    uint8_t u = q[4 * i], v = q[4 * i + 1], w = q[4 * i + 2], h = q[4 * i + 3];
                    vxo_tinyquad_pack(u, v, w, h, (uint8_t)t, quads_out + 3 * n);
                    n++;
                    int lo[3], hi[3];
                    if (axis == 0) { lo[0] = axis_pos; lo[1] = u; lo[2] = v; hi[0] = axis_pos; hi[1] = u + w; hi[2] = v + h; }
                    else if (axis == 1) { lo[0] = u; lo[1] = axis_pos; lo[2] = v; hi[0] = u + w; hi[1] = axis_pos; hi[2] = v + h; }
                    else { lo[0] = u; lo[1] = v; lo[2] = axis_pos; hi[0] = u + w; hi[1] = v + h; hi[2] = axis_pos; }
                    for (int k = 0; k < 3; ++k) { mn[k] = imin(mn[k], lo[k]); mx[k] = imax(mx[k], hi[k]); }
                }
            }
        }
        slice_offsets[face * 33 + 32] = (uint32_t)n;
    }
    return n;
}

/* mesh_world binary_greedy.rs:62-78 with explicit neighbour table. */
int64_t vxo_mesh_chunks(const uint8_t *voxels, const int32_t *neighbors, const uint8_t *uniform_flags,
                        int32_t n_chunks, uint8_t *quads_out, int64_t cap, uint32_t *quad_base,
                        uint32_t *quad_count, uint32_t *slice_offsets, int32_t *face_aabb, uint8_t *has_mesh) {
    int64_t total = 0;
    for (int32_t i = 0; i < n_chunks; ++i) {
        quad_base[i] = (uint32_t)total;
        quad_count[i] = 0;
        has_mesh[i] = 0;
        uint32_t *so = slice_offsets + (size_t)i * 6 * 33;
        int32_t *ab = face_aabb + (size_t)i * 36;
        memset(so, 0, sizeof(uint32_t) * 6 * 33);
        for (int f = 0; f < 6; ++f) { ab[f * 6 + 0] = ab[f * 6 + 1] = ab[f * 6 + 2] = 32; ab[f * 6 + 3] = ab[f * 6 + 4] = ab[f * 6 + 5] = 0; }
        if (uniform_flags && uniform_flags[i]) continue; /* is_uniform fast path :87 */
        const uint8_t *nbv[6];
        int32_t code[6];
        for (int f = 0; f < 6; ++f) {
            int32_t nb = neighbors ? neighbors[(size_t)i * 6 + f] : VXO_NBR_NONE;
            nbv[f] = NULL;
            code[f] = VXO_NBR_NONE;
            if (nb >= 0) {
                if (uniform_flags && uniform_flags[nb]) code[f] = (uniform_flags[nb] - 1) != 0 ? VXO_NBR_UNIFORM_SOLID : VXO_NBR_UNIFORM_AIR;
                else nbv[f] = voxels + (size_t)nb * VXO_CHUNK_VOLUME;
            } else code[f] = nb;
        }
        int64_t room = cap - total;
        int n = vxo_mesh_chunk(voxels + (size_t)i * VXO_CHUNK_VOLUME, nbv, code, quads_out + 3 * total,
                               room > 0x7fffffff ? 0x7fffffff : (int)room, so, ab);
        if (n < 0) return -1;
        quad_count[i] = (uint32_t)n;
        has_mesh[i] = n > 0; /* mesh.is_empty() -> None :116-120 */
        total += n;
    }
    return total;
}

/* ---------- culling ---------------------------------------------------- */

/* ----------------------------------------------------------------------------------------------
 * Terrain generation: Chunk::generate_terrain (voxel/chunk.rs:114-207) over `noise 0.9.0` Perlin::new(12345).
 *
 * The `noise` crate (Cargo.lock: noise 0.9.0, rand 0.8.5, rand_xorshift 0.3.0) is a crates.io dependency that is NOT
 * vendored under /root/reference and cannot be fetched here, so the functions below restate its published algorithm from
 * the crate's source as documented:
 *   PermutationTable::new(seed)  permutation_table.rs: 16 seed bytes = [1,0,0,0, seed_le, seed_le, seed_le] into
 *                                XorShiftRng (rand_xorshift: t = x ^ x << 11; x,y,z = y,z,w; w = w ^ w >> 19 ^ t ^ t >> 8),
 *                                then (0..256).shuffle(rng) with rand 0.8.5's SliceRandom::shuffle (i from len-1 down to
 *                                1: swap(i, gen_range(0..i+1))) and UniformInt<u32>::sample_single (widening multiply,
 *                                zone = (range << leading_zeros(range)) - 1, accept when the low word <= zone)
 *   NoiseHasher::hash            values[values[x & 255] ^ (y & 255)]
 *   perlin_2d                    core/perlin.rs: corner = floor(point), four gradients from hash & 3 over
 *                                (+x+y, -x+y, +x-y, -x-y), quintic fade t^3 (t (6 t - 15) + 10) of the clamped distance,
 *                                bilinear_interpolation k0 + k1 u + k2 v + k3 u v, scaled by 2 / sqrt(2), clamped to [-1, 1]
 * No reference test pins a height, so until tools/ref_dump vectors exist these heights are PARITY UNPINNED (header of
 * this file).  What IS pinned: the CUDA generator and the host generator reproduce this restatement bit for
 * bit (tests/test_mesher_gpu.py, tests/test_oracle_kat.py).
 * ---------------------------------------------------------------------------------------------- */
void vxo_noise_permutation_table(uint32_t seed, uint8_t values[256]) {
    uint32_t x = 1u, y = seed, z = seed, w = seed; /* le::read_u32_into of the 16 seed bytes */
    if (x == 0 && y == 0 && z == 0 && w == 0) { x = 0x0BAD5EEDu; y = 0x0BAD5EEDu; z = 0x0BAD5EEDu; w = 0x0BAD5EEDu; }
    for (int i = 0; i < 256; ++i) values[i] = (uint8_t)i;
    for (uint32_t i = 255; i >= 1; --i) {
        const uint32_t range = i + 1u; /* gen_range(0..i+1) */
        const uint32_t zone = (range << __builtin_clz(range)) - 1u;
        uint32_t pick;
        for (;;) {
            const uint32_t t = x ^ (x << 11);
            x = y; y = z; z = w;
            w = w ^ (w >> 19) ^ (t ^ (t >> 8));
            const uint64_t m = (uint64_t)w * (uint64_t)range;
            if ((uint32_t)m <= zone) { pick = (uint32_t)(m >> 32); break; }
        }
        const uint8_t tmp = values[i]; values[i] = values[pick]; values[pick] = tmp;
    }
}

static inline double noise_quintic(double t) {
    const double x = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    return x * x * x * (x * (x * 6.0 - 15.0) + 10.0);
}

double vxo_perlin2(const uint8_t values[256], double px, double py) {
    const double fx = floor(px), fy = floor(py);
    const int64_t cx = (int64_t)fx, cy = (int64_t)fy;
    const double dx = px - (double)cx, dy = py - (double)cy;
    double g[2][2];
    for (int ox = 0; ox < 2; ++ox)
        for (int oy = 0; oy < 2; ++oy) {
            const double qx = dx - (double)ox, qy = dy - (double)oy;
            const unsigned h = values[values[(unsigned)((cx + ox) & 0xff)] ^ (unsigned)((cy + oy) & 0xff)] & 3u;
            g[ox][oy] = h == 0 ? qx + qy : h == 1 ? -qx + qy : h == 2 ? qx - qy : -qx - qy;
        }
    const double u = noise_quintic(dx), v = noise_quintic(dy);
    const double g00 = g[0][0], g10 = g[1][0], g01 = g[0][1], g11 = g[1][1];
    const double k0 = g00, k1 = g10 - g00, k2 = g01 - g00, k3 = g00 + g11 - g10 - g01;
    const double unscaled = k0 + k1 * u + k2 * v + k3 * u * v;
    const double scaled = unscaled * (2.0 / 1.4142135623730951); /* 2 / core::f64::consts::SQRT_2 */
    return scaled < -1.0 ? -1.0 : (scaled > 1.0 ? 1.0 : scaled);
}

/* sample_terrain_height chunk.rs:173-177 */
int32_t vxo_terrain_height(const uint8_t values[256], int32_t x, int32_t z) {
    const double n = vxo_perlin2(values, (double)x * 0.01, (double)z * 0.01);
    return (int32_t)(n * 20.0); /* `as i32`: truncation, |n * 20| <= 20 */
}

/* Chunk::generate_terrain chunk.rs:114-170.  voxels_out: 32768 bytes (index z*1024 + y*32 + x, chunk.rs:52), written only
 * for Varied chunks.  Returns the uniform flag: 0 Varied, 1 Uniform(Air), 4 Uniform(Stone) (1 + BlockType). */
int vxo_generate_terrain(const int32_t pos[3], uint32_t seed, uint8_t *voxels_out) {
    uint8_t values[256];
    vxo_noise_permutation_table(seed, values);
    int32_t h[CS][CS], mn = INT32_MAX, mx = INT32_MIN; /* [z][x]; get_height_range chunk.rs:191-207 */
    for (int z = 0; z < CS; ++z)
        for (int x = 0; x < CS; ++x) {
            h[z][x] = vxo_terrain_height(values, pos[0] * CS + x, pos[2] * CS + z);
            if (h[z][x] < mn) mn = h[z][x];
            if (h[z][x] > mx) mx = h[z][x];
        }
    const int32_t y0 = pos[1] * CS;
    if (y0 > mx) return 1;            /* all air above terrain :127-129 */
    if (y0 + CS < mn - 10) return 4;  /* all solid below :132-134 */
    for (int z = 0; z < CS; ++z)
        for (int y = 0; y < CS; ++y)
            for (int x = 0; x < CS; ++x) {
                const int32_t wy = y0 + y, hh = h[z][x];
                uint8_t t = 3;                 /* Stone */
                if (wy > hh) t = 0;            /* Air   :145-146 */
                else if (wy == hh) t = 1;      /* Grass */
                else if (wy > hh - 3) t = 2;   /* Dirt  */
                voxels_out[z * CS * CS + y * CS + x] = t;
            }
    return 0;
}

/* camera/mod.rs:123-160.  row(i) = (c0[i], c1[i], c2[i], c3[i]). */
void vxo_frustum_from_vp(const float vp[16], float planes[24]) {
    float row[4][4];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) row[r][c] = vp[c * 4 + r];
    for (int p = 0; p < 6; ++p) {
        int r = p >> 1;
        float pl[4];
        for (int k = 0; k < 4; ++k) pl[k] = (p & 1) ? row[3][k] - row[r][k] : row[3][k] + row[r][k];
        /* normalize_plane :153-160 ; Vec3::length = sqrt(x*x + y*y + z*z) */
        float len = sqrtf(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]);
        if (len > 0.0001f)
            for (int k = 0; k < 4; ++k) pl[k] = pl[k] / len;
        memcpy(planes + p * 4, pl, sizeof(pl));
    }
}

/* camera/mod.rs:164-183 */
int vxo_frustum_intersects_aabb(const float planes[24], const float mn[3], const float mx[3]) {
    for (int p = 0; p < 6; ++p) {
        const float *pl = planes + p * 4;
        float px = pl[0] > 0.0f ? mx[0] : mn[0];
        float py = pl[1] > 0.0f ? mx[1] : mn[1];
        float pz = pl[2] > 0.0f ? mx[2] : mn[2];
        if (pl[0] * px + pl[1] * py + pl[2] * pz + pl[3] < 0.0f) return 0;
    }
    return 1;
}

/* world.rs:118-146 with world_to_chunk_pos :201-207 and chunk_bounds :211-215 */
void vxo_cull_chunks(const int32_t *positions, int32_t n, const float vp[16], const float cam_pos[3],
                     int32_t view_distance, int32_t frustum_culling, uint8_t *visible_out) {
    float planes[24];
    vxo_frustum_from_vp(vp, planes);
    int32_t cc[3];
    for (int k = 0; k < 3; ++k) cc[k] = f2i(floorf(cam_pos[k] / (float)CS));
    float vd_sq = (float)(view_distance * view_distance);
    for (int32_t i = 0; i < n; ++i) {
        const int32_t *p = positions + 3 * (size_t)i;
        int32_t dx = p[0] - cc[0], dy = p[1] - cc[1], dz = p[2] - cc[2];
        float dist_sq = (float)(dx * dx + dy * dy + dz * dz);
        if (dist_sq > vd_sq) { visible_out[i] = 0; continue; }
        if (frustum_culling) {
            float mn[3], mx[3];
            for (int k = 0; k < 3; ++k) { mn[k] = (float)(p[k] * CS); mx[k] = mn[k] + (float)CS; }
            visible_out[i] = (uint8_t)vxo_frustum_intersects_aabb(planes, mn, mx);
        } else visible_out[i] = 1;
    }
}

/* stable merge sort of an index array by float key; comparator =
 * partial_cmp().unwrap_or(Equal) (main.rs:372-376, :494-498, culling.rs:46-50) */
static void stable_sort_by_key(int32_t *idx, int32_t n, const float *key) {
    if (n < 2) return;
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int32_t w = 1; w < n; w *= 2) {
        for (int32_t lo = 0; lo < n; lo += 2 * w) {
            int32_t mid = imin(lo + w, n), hi = imin(lo + 2 * w, n);
            int32_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                if (key[idx[b]] < key[idx[a]]) tmp[o++] = idx[b++]; /* strictly less moves ahead */
                else tmp[o++] = idx[a++];
            }
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, sizeof(int32_t) * (size_t)n);
    }
    free(tmp);
}

/* culling.rs:40-119 */
int vxo_horizon_cull(const float cam_pos[3], const float *centers, int32_t n, int32_t *order, int32_t bins,
                     float base_margin, float margin_dist_factor, float min_dist_chunks) {
    if (n <= 0) return 0;
    float *dsq = (float *)malloc(sizeof(float) * (size_t)n);
    int32_t maxid = 0;
    for (int32_t i = 0; i < n; ++i) maxid = imax(maxid, order[i]);
    float *key = (float *)malloc(sizeof(float) * (size_t)(maxid + 1));
    for (int32_t i = 0; i < n; ++i) {
        const float *c = centers + 3 * (size_t)order[i];
        float dx = c[0] - cam_pos[0], dy = c[1] - cam_pos[1], dz = c[2] - cam_pos[2];
        key[order[i]] = dx * dx + dy * dy + dz * dz;
    }
    stable_sort_by_key(order, n, key);
    float *horizon = (float *)malloc(sizeof(float) * (size_t)bins);
    for (int32_t b = 0; b < bins; ++b) horizon[b] = -INFINITY;
    int32_t write_idx = 0;
    const float chunk_size = (float)CS, half_chunk = chunk_size * 0.5f;
    const float PI = 3.14159265358979323846f;
    for (int32_t i = 0; i < n; ++i) {
        int32_t id = order[i];
        const float *c = centers + 3 * (size_t)id;
        float tx = c[0] - cam_pos[0], tz = c[2] - cam_pos[2];
        float dist_xz = sqrtf(tx * tx + tz * tz);
        if (dist_xz < 1e-3f) { order[write_idx++] = id; continue; }
        float dist_chunks = dist_xz / chunk_size;
        if (dist_chunks < min_dist_chunks) { order[write_idx++] = id; continue; }
        float angle = atan2f(tz, tx);
        float bin_f = (angle + PI) / (2.0f * PI) * (float)bins;
        long bin = (long)f2i(floorf(bin_f));
        if (bin < 0) bin += bins;
        bin = bin % bins;
        float height = c[1] - cam_pos[1];
        float slope = height / dist_xz;
        float margin = base_margin * (1.0f + dist_chunks * margin_dist_factor);
        float cur = horizon[bin];
        int should_cull = slope >= 0.0f && (slope + margin) < cur;
        if (!should_cull) {
            order[write_idx++] = id;
            float top_slope = (c[1] + half_chunk - cam_pos[1]) / dist_xz;
            if (top_slope > cur) horizon[bin] = top_slope;
        }
    }
    free(horizon); free(key); free(dsq);
    return write_idx;
}

/* ---------- textures / shading ---------------------------------------- */

/* texture.rs:42-54 */
static uint32_t rgb565_to_argb32(uint16_t c) {
    uint32_t r = (c >> 11) & 0x1F, g = (c >> 5) & 0x3F, b = c & 0x1F;
    uint32_t r8 = (r << 3) | (r >> 2), g8 = (g << 2) | (g >> 4), b8 = (b << 3) | (b >> 2);
    return 0xFF000000u | (r8 << 16) | (g8 << 8) | b8;
}
/* texture.rs:81-101 */
static void create_checkerboard(uint16_t c1, uint16_t c2, uint32_t pal[16], uint8_t ind[32]) {
    memset(pal, 0, 64); memset(ind, 0, 32);
    pal[0] = rgb565_to_argb32(c1); pal[1] = rgb565_to_argb32(c2);
    for (int i = 0; i < 64; ++i) {
        int x = i % 8, y = i / 8;
        uint8_t ci = (uint8_t)((x + y) % 2);
        if (i % 2 == 0) ind[i / 2] |= (uint8_t)(ci << 4); else ind[i / 2] |= ci;
    }
}
/* texture.rs:103-123 */
static void create_noise(uint16_t base, uint16_t dark, uint32_t pal[16], uint8_t ind[32]) {
    for (int i = 0; i < 16; ++i) pal[i] = rgb565_to_argb32((i % 2 == 0) ? base : dark);
    uint32_t seed = 12345;
    for (int i = 0; i < 32; ++i) {
        seed = seed * 1103515245u + 12345u;
        ind[i] = (uint8_t)(seed >> 16);
    }
}
/* texture.rs:60-79 */
void vxo_default_atlas(vxo_atlas *a) {
    create_checkerboard(0xF81F, 0x0000, a->palette[0], a->indices[0]);
    create_noise(0x03E0, 0x02E0, a->palette[1], a->indices[1]);
    create_noise(0x8A22, 0x71C2, a->palette[2], a->indices[2]);
    create_noise(0x8410, 0x73AE, a->palette[3], a->indices[3]);
}
/* texture.rs:19-38 */
uint32_t vxo_texture_sample(const vxo_atlas *a, int tex, uint8_t u, uint8_t v) {
    uint8_t x = u & 7, y = v & 7;
    uint8_t pixel_idx = (uint8_t)((y << 3) | x);
    uint8_t byte = a->indices[tex][pixel_idx >> 1];
    uint8_t pi = ((pixel_idx & 1) == 0) ? ((byte >> 4) & 0xF) : (byte & 0xF);
    return a->palette[tex][pi];
}

void vxo_default_frame_config(vxo_frame_config *cfg, int w, int h) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->width = w; cfg->height = h;
    cfg->clear_color = 0xFF87CEEBu; /* main.rs:393 */
    cfg->backface_culling = 1;      /* rasterizer.rs:366 */
    cfg->enable_shading = 1;        /* rasterizer.rs:368 */
    /* shading.rs:21-31; normalized constants as rasterizer.rs:1206-1208 */
    cfg->light_dir[0] = 0.35634832f; cfg->light_dir[1] = 0.8908708f; cfg->light_dir[2] = 0.2672612f;
    cfg->ambient = 0.35f; cfg->diffuse = 0.65f;
    cfg->n_threads = 1;
    cfg->occlusion_culling = 0;     /* main.rs:112 */
    cfg->occlusion_grid_w = 128; cfg->occlusion_grid_h = 72; /* main.rs:46-47 */
}

/* shading.rs:90-110 */
uint32_t vxo_shade_color_u32(uint32_t base, float light) {
    uint32_t r = (base >> 16) & 0xFF, g = (base >> 8) & 0xFF, b = base & 0xFF;
    float lf = light * 256.0f;
    uint32_t light_fp = lf != lf ? 0u : (lf <= 0.0f ? 0u : (lf >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)lf));
    uint32_t rl = (r * light_fp) >> 8, gl = (g * light_fp) >> 8, bl = (b * light_fp) >> 8;
    if (rl > 255) rl = 255;
    if (gl > 255) gl = 255;
    if (bl > 255) bl = 255;
    return 0xFF000000u | (rl << 16) | (gl << 8) | bl;
}

/* rasterizer.rs:1204-1216 (constants come from cfg; defaults equal the reference's) */
float vxo_face_light(const vxo_frame_config *cfg, int face) {
    float nx = 0, ny = 0, nz = 0;
    switch (face) {
    case 0: nx = 1; break; case 1: nx = -1; break;
    case 2: ny = 1; break; case 3: ny = -1; break;
    case 4: nz = 1; break; default: nz = -1; break;
    }
    float lambert = rmax(nx * cfg->light_dir[0] + ny * cfg->light_dir[1] + nz * cfg->light_dir[2], 0.0f);
    float light = cfg->ambient + cfg->diffuse * lambert;
    /* f32::clamp(0,1) */
    if (light < 0.0f) light = 0.0f;
    if (light > 1.0f) light = 1.0f;
    return light;
}

/* ---------- span rasterizer ------------------------------------------- */

typedef struct { float pos[4]; float uv[2]; } clip_vtx;         /* ClipTexturedVertex */
typedef struct { float x, y, z, u_over_w, v_over_w, inv_w; } span_vtx; /* SpanVertex :1308-1315 */

typedef struct {
    int W, H;           /* target.width(), target.full_height() */
    int rx0, ry0, rw, rh; /* target.rect() */
    uint32_t *color; float *depth;
    const vxo_frame_config *cfg;
    const vxo_atlas *atlas;
    int barycentric;    /* render_mesh_tiny_quads(.., use_span_renderer = false) */
} target_t;

/* rasterizer.rs:2628-2641 */
static clip_vtx intersect_near_textured(const clip_vtx *a, const clip_vtx *b, float threshold) {
    float wa = a->pos[3], wb = b->pos[3];
    float t = (threshold - wa) / (wb - wa);
    clip_vtx r;
    for (int k = 0; k < 4; ++k) r.pos[k] = a->pos[k] + (b->pos[k] - a->pos[k]) * t;
    for (int k = 0; k < 2; ++k) r.uv[k] = a->uv[k] + (b->uv[k] - a->uv[k]) * t;
    return r;
}

/* rasterizer.rs:2645-2697 */
static int clip_triangle_near_textured(const clip_vtx tri[3], float threshold, clip_vtx out_tris[2][3]) {
    clip_vtx output[4];
    int out_len = 0;
    clip_vtx prev = tri[2];
    int prev_inside = prev.pos[3] >= threshold;
    for (int i = 0; i < 3; ++i) {
        clip_vtx curr = tri[i];
        int curr_inside = curr.pos[3] >= threshold;
        if (prev_inside && curr_inside) output[out_len++] = curr;
        else if (prev_inside && !curr_inside) output[out_len++] = intersect_near_textured(&prev, &curr, threshold);
        else if (!prev_inside && curr_inside) {
            output[out_len++] = intersect_near_textured(&prev, &curr, threshold);
            output[out_len++] = curr;
        }
        prev = curr;
        prev_inside = curr_inside;
    }
    if (out_len == 3) {
        out_tris[0][0] = output[0]; out_tris[0][1] = output[1]; out_tris[0][2] = output[2];
        return 1;
    }
    if (out_len == 4) {
        out_tris[0][0] = output[0]; out_tris[0][1] = output[1]; out_tris[0][2] = output[2];
        out_tris[1][0] = output[0]; out_tris[1][1] = output[2]; out_tris[1][2] = output[3];
        return 2;
    }
    return 0;
}

#ifdef VXO_STATS
/* workload statistics (tools/frame_stats_cpu.py builds a separate library with -DVXO_STATS; single-threaded use) */
uint64_t vxo_stats[64]; /* 0 triangles setup, 1 rows visited, 2 spans, 3 fragments, 4 depth passes, 8.. span length histogram (log2) */
#define VXO_STAT(i, n) (vxo_stats[i] += (uint64_t)(n))
#else
#define VXO_STAT(i, n) ((void)0)
#endif

/* rasterizer.rs:1219-1467 */
static void render_triangle_span_from_clip(const clip_vtx tri_in[3], int block_type, float light, target_t *tg) {
    const float NEAR_W_EPS = 0.001f; /* rasterizer.rs:18 */
    clip_vtx tris[2][3];
    int tri_count = clip_triangle_near_textured(tri_in, NEAR_W_EPS, tris);
    if (tri_count == 0) return;

    float fb_width = (float)tg->W, fb_height = (float)tg->H;
    int rect_x0 = tg->rx0, rect_y0 = tg->ry0;
    float rect_x_limit = (float)(tg->rx0 + tg->rw);
    float rect_y_limit = (float)(tg->ry0 + tg->rh);
    int tex_id = block_type; /* block_type.rs:58-65 */

    for (int ti = 0; ti < tri_count; ++ti) {
        const clip_vtx *tri = tris[ti];
        float ndc[3][4];
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 4; ++k) ndc[i][k] = tri[i].pos[k] / tri[i].pos[3];

        if (tg->cfg->backface_culling) {
            float v01x = ndc[1][0] - ndc[0][0], v01y = ndc[1][1] - ndc[0][1];
            float v02x = ndc[2][0] - ndc[0][0], v02y = ndc[2][1] - ndc[0][1];
            float cross_z = v01x * v02y - v01y * v02x;
            if (cross_z <= 0.0f) continue;
        }

        float sx[3], sy[3];
        for (int i = 0; i < 3; ++i) { /* ndc_to_screen :2546-2551 */
            sx[i] = (ndc[i][0] + 1.0f) * 0.5f * fb_width;
            sy[i] = (1.0f - ndc[i][1]) * 0.5f * fb_height;
        }
        float min_y = rmin(rmin(sy[0], sy[1]), sy[2]);
        float max_y = rmax(rmax(sy[0], sy[1]), sy[2]);
        min_y = rmax(min_y, (float)rect_y0);
        max_y = rmin(max_y, rect_y_limit);
        if (min_y > max_y) continue;

        span_vtx vs[3];
        for (int i = 0; i < 3; ++i) {
            vs[i].x = sx[i]; vs[i].y = sy[i]; vs[i].z = ndc[i][2];
            vs[i].u_over_w = tri[i].uv[0] / tri[i].pos[3];
            vs[i].v_over_w = tri[i].uv[1] / tri[i].pos[3];
            vs[i].inv_w = 1.0f / tri[i].pos[3];
        }

        int y_start = f2i(floorf(min_y));
        int y_end = f2i(ceilf(max_y));
        VXO_STAT(0, 1);
        for (int y = y_start; y <= y_end; ++y) {
            if (y < rect_y0 || y >= f2i(rect_y_limit)) continue;
            VXO_STAT(1, 1);
            float y_center = (float)y + 0.5f;
            span_vtx pts[2] = {vs[0], vs[0]};
            int count = 0;
            for (int i = 0; i < 3; ++i) {
                span_vtx v0 = vs[i], v1 = vs[(i + 1) % 3];
                float y0 = v0.y, y1 = v1.y;
                if ((y0 <= y_center && y_center < y1) || (y1 <= y_center && y_center < y0)) {
                    float dy = y1 - y0;
                    if (fabsf(dy) < 1e-6f) continue;
                    float t = (y_center - y0) / dy;
                    span_vtx p;
                    p.x = v0.x + (v1.x - v0.x) * t;
                    p.y = y_center;
                    p.z = v0.z + (v1.z - v0.z) * t;
                    p.u_over_w = v0.u_over_w + (v1.u_over_w - v0.u_over_w) * t;
                    p.v_over_w = v0.v_over_w + (v1.v_over_w - v0.v_over_w) * t;
                    p.inv_w = v0.inv_w + (v1.inv_w - v0.inv_w) * t;
                    pts[count++] = p;
                    if (count == 2) break;
                }
            }
            if (count < 2) continue;
            if (pts[0].x > pts[1].x) { span_vtx tmp = pts[0]; pts[0] = pts[1]; pts[1] = tmp; }

            float x_start_f = rmax(pts[0].x, (float)rect_x0);
            float x_end_f = rmin(pts[1].x, rect_x_limit);
            int x_start = f2i(ceilf(x_start_f - 0.5f));
            int x_end = f2i(floorf(x_end_f - 0.5f));
            if (x_start > x_end) continue;

            float span_width = pts[1].x - pts[0].x;
            if (fabsf(span_width) < 1e-6f) continue;
            float inv_span = 1.0f / span_width;

            float offset = ((float)x_start + 0.5f) - pts[0].x;
            float z_val = pts[0].z + (pts[1].z - pts[0].z) * inv_span * offset;
            float u_over_w = pts[0].u_over_w + (pts[1].u_over_w - pts[0].u_over_w) * inv_span * offset;
            float v_over_w = pts[0].v_over_w + (pts[1].v_over_w - pts[0].v_over_w) * inv_span * offset;
            float inv_w = pts[0].inv_w + (pts[1].inv_w - pts[0].inv_w) * inv_span * offset;

            float step_z = (pts[1].z - pts[0].z) * inv_span;
            float step_u = (pts[1].u_over_w - pts[0].u_over_w) * inv_span;
            float step_v = (pts[1].v_over_w - pts[0].v_over_w) * inv_span;
            float step_w = (pts[1].inv_w - pts[0].inv_w) * inv_span;
#ifdef VXO_STATS
            {
                int len = x_end - x_start + 1, lg = 0;
                while ((1 << (lg + 1)) <= len) lg++;
                VXO_STAT(2, 1);
                VXO_STAT(3, len);
                VXO_STAT(8 + lg, 1);
                VXO_STAT(24 + lg, len);
            }
#endif

            for (int x = x_start; x <= x_end; ++x) {
                /* FrameSlice::test_depth_and_get_index framebuffer.rs:30-51.  The reference
                 * indexes `y_local*width + x` unchecked in x; the clamp of x_end_f to the rect
                 * limit keeps x < W, except for x == W when pts[1].x >= W + 0.5 can not happen
                 * (x_end = floor(min(.., W) - 0.5) <= W - 1). */
                size_t idx = (size_t)y * (size_t)tg->W + (size_t)x;
                if (z_val < tg->depth[idx]) {
                    VXO_STAT(4, 1);
                    tg->depth[idx] = z_val;
                    float u = u_over_w / inv_w;
                    float v = v_over_w / inv_w;
                    uint8_t tex_u = (uint8_t)(f2i(u * 8.0f) & 7);
                    uint8_t tex_v = (uint8_t)(f2i(v * 8.0f) & 7);
                    uint32_t c = vxo_texture_sample(tg->atlas, tex_id, tex_u, tex_v);
                    if (tg->cfg->enable_shading) c = vxo_shade_color_u32(c, light);
                    tg->color[idx] = c;
                }
                z_val += step_z;
                u_over_w += step_u;
                v_over_w += step_v;
                inv_w += step_w;
            }
        }
    }
}

/* Rasterizer::edge_function rasterizer.rs:2556-2558 */
static inline float edge_function(float ax, float ay, float bx, float by, float cx, float cy) {
    return (cx - ax) * (by - ay) - (cy - ay) * (bx - ax);
}

/* render_triangle_from_clip_textured rasterizer.rs:1881-2107: the barycentric rasterizer the mesh path falls back to
 * when the camera is rolled (render_mesh_with_up with |up.y| < 0.995, :377-411) or when render_mesh_tiny_quads is
 * called with use_span_renderer = false.  Edge functions advance by one rounded add per pixel / per row from the top-left
 * pixel of the triangle's box clipped to the framebuffer and the target rect. */
static void render_triangle_from_clip_textured(const clip_vtx tri_in[3], int block_type, float light, target_t *tg) {
    const float NEAR_W_EPS = 0.001f;
    clip_vtx tris[2][3];
    int tri_count = clip_triangle_near_textured(tri_in, NEAR_W_EPS, tris);
    if (tri_count == 0) return;
    const float fb_width = (float)tg->W, fb_height = (float)tg->H;
    const int tex_id = block_type;
    for (int ti = 0; ti < tri_count; ++ti) {
        const clip_vtx *tri = tris[ti];
        const float w0c = tri[0].pos[3], w1c = tri[1].pos[3], w2c = tri[2].pos[3];
        float ndc[3][3];
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k) ndc[i][k] = tri[i].pos[k] / tri[i].pos[3];
        if (tg->cfg->backface_culling) {
            float v01x = ndc[1][0] - ndc[0][0], v01y = ndc[1][1] - ndc[0][1];
            float v02x = ndc[2][0] - ndc[0][0], v02y = ndc[2][1] - ndc[0][1];
            float cross_z = v01x * v02y - v01y * v02x;
            if (cross_z <= 0.0f) continue;
        }
        float px[3], py[3];
        for (int i = 0; i < 3; ++i) {
            px[i] = (ndc[i][0] + 1.0f) * 0.5f * fb_width;
            py[i] = (1.0f - ndc[i][1]) * 0.5f * fb_height;
        }
        const float z0 = ndc[0][2], z1 = ndc[1][2], z2 = ndc[2][2];
        int min_x = f2i(floorf(rmin(rmin(px[0], px[1]), px[2])));
        int max_x = f2i(ceilf(rmax(rmax(px[0], px[1]), px[2])));
        int min_y = f2i(floorf(rmin(rmin(py[0], py[1]), py[2])));
        int max_y = f2i(ceilf(rmax(rmax(py[0], py[1]), py[2])));
        min_x = imax(min_x, 0); max_x = imin(max_x, tg->W - 1);
        min_y = imax(min_y, 0); max_y = imin(max_y, tg->H - 1);
        const int tx1 = tg->rx0 + tg->rw - 1, ty1 = tg->ry0 + tg->rh - 1;
        min_x = imax(min_x, tg->rx0); max_x = imin(max_x, tx1);
        min_y = imax(min_y, tg->ry0); max_y = imin(max_y, ty1);
        if (min_x > max_x || min_y > max_y) continue;
        const float area = edge_function(px[0], py[0], px[1], py[1], px[2], py[2]);
        if (area <= 0.0f) continue;
        if (area < 0.1f) continue; /* MIN_TRIANGLE_AREA :1998 */
        const float inv_area = 1.0f / area;
        const float e0dx = py[2] - py[1], e0dy = px[1] - px[2];
        const float e1dx = py[0] - py[2], e1dy = px[2] - px[0];
        const float e2dx = py[1] - py[0], e2dy = px[0] - px[1];
        const float inv_w0 = 1.0f / w0c, inv_w1 = 1.0f / w1c, inv_w2 = 1.0f / w2c;
        const float u0w = tri[0].uv[0] * inv_w0, u1w = tri[1].uv[0] * inv_w1, u2w = tri[2].uv[0] * inv_w2;
        const float v0w = tri[0].uv[1] * inv_w0, v1w = tri[1].uv[1] * inv_w1, v2w = tri[2].uv[1] * inv_w2;
        const float sx = (float)min_x + 0.5f, sy = (float)min_y + 0.5f;
        float w0_row = edge_function(px[1], py[1], px[2], py[2], sx, sy);
        float w1_row = edge_function(px[2], py[2], px[0], py[0], sx, sy);
        float w2_row = edge_function(px[0], py[0], px[1], py[1], sx, sy);
        for (int y = min_y; y <= max_y; ++y) {
            float w0 = w0_row, w1 = w1_row, w2 = w2_row;
            for (int x = min_x; x <= max_x; ++x) {
                if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {
                    const float bw0 = w0 * inv_area, bw1 = w1 * inv_area, bw2 = w2 * inv_area;
                    const float depth = bw0 * z0 + bw1 * z1 + bw2 * z2;
                    size_t idx = (size_t)y * (size_t)tg->W + (size_t)x;
                    if (depth < tg->depth[idx]) {
                        tg->depth[idx] = depth;
                        const float inv_w_interp = bw0 * inv_w0 + bw1 * inv_w1 + bw2 * inv_w2;
                        const float u = (bw0 * u0w + bw1 * u1w + bw2 * u2w) / inv_w_interp;
                        const float v = (bw0 * v0w + bw1 * v1w + bw2 * v2w) / inv_w_interp;
                        uint8_t tex_u = (uint8_t)(f2i(u * 8.0f) & 7);
                        uint8_t tex_v = (uint8_t)(f2i(v * 8.0f) & 7);
                        uint32_t c = vxo_texture_sample(tg->atlas, tex_id, tex_u, tex_v);
                        if (tg->cfg->enable_shading) c = vxo_shade_color_u32(c, light);
                        tg->color[idx] = c;
                    }
                }
                w0 += e0dx; w1 += e1dx; w2 += e2dx;
            }
            w0_row += e0dy; w1_row += e1dy; w2_row += e2dy;
        }
    }
}

/* vertex table rasterizer.rs:1092-1129 (== mesh.rs:624-661); `u + w` is u8 arithmetic. */
static void quad_local_positions(int face, uint8_t s, uint8_t u, uint8_t v, uint8_t w, uint8_t h, float lp[4][3], float uv[4][2]) {
    float fs = (float)s, u0 = (float)u, v0 = (float)v, u1 = (float)(uint8_t)(u + w), v1 = (float)(uint8_t)(v + h);
    /* corner order per face: sequence of (du,dv) flags */
    static const int order[6][4][2] = {
        {{0, 0}, {1, 0}, {1, 1}, {0, 1}}, /* +X */
        {{0, 0}, {0, 1}, {1, 1}, {1, 0}}, /* -X */
        {{0, 0}, {0, 1}, {1, 1}, {1, 0}}, /* +Y */
        {{0, 0}, {1, 0}, {1, 1}, {0, 1}}, /* -Y */
        {{0, 0}, {1, 0}, {1, 1}, {0, 1}}, /* +Z */
        {{0, 0}, {0, 1}, {1, 1}, {1, 0}}, /* -Z */
    };
    int axis = face >> 1;
    for (int i = 0; i < 4; ++i) {
        float cu = order[face][i][0] ? u1 : u0;
        float cv = order[face][i][1] ? v1 : v0;
        if (axis == 0) { lp[i][0] = fs; lp[i][1] = cu; lp[i][2] = cv; }
        else if (axis == 1) { lp[i][0] = cu; lp[i][1] = fs; lp[i][2] = cv; }
        else { lp[i][0] = cu; lp[i][1] = cv; lp[i][2] = fs; }
        uv[i][0] = cu; uv[i][1] = cv; /* rasterizer.rs:1136-1173 */
    }
}

void vxo_quad_clip_vertices(int face, uint8_t slice_pos, uint8_t u, uint8_t v, uint8_t w, uint8_t h,
                            const int32_t chunk_pos[3], const float vp[16], float clip[16]) {
    float lp[4][3], uv[4][2];
    quad_local_positions(face, slice_pos, u, v, w, h, lp, uv);
    float off[3];
    for (int k = 0; k < 3; ++k) off[k] = (float)(chunk_pos[k] * CS); /* mesh.rs:483-485 */
    for (int i = 0; i < 4; ++i)
        mat4_mul_vec4(vp, off[0] + lp[i][0], off[1] + lp[i][1], off[2] + lp[i][2], 1.0f, clip + 4 * i);
}

/* rasterizer.rs:1074-1201 */
static void render_tiny_quad_span(const uint8_t q3[3], int face, uint8_t slice_pos, const float off[3], float light,
                                  const float vp[16], target_t *tg) {
    uint8_t u, v, w, h, bt;
    vxo_tinyquad_unpack(q3, &u, &v, &w, &h, &bt);
    float lp[4][3], uv[4][2];
    quad_local_positions(face, slice_pos, u, v, w, h, lp, uv);
    clip_vtx cv[4];
    for (int i = 0; i < 4; ++i) {
        mat4_mul_vec4(vp, off[0] + lp[i][0], off[1] + lp[i][1], off[2] + lp[i][2], 1.0f, cv[i].pos);
        cv[i].uv[0] = uv[i][0]; cv[i].uv[1] = uv[i][1];
    }
    static const int tri_idx[2][3] = {{0, 1, 2}, {0, 2, 3}};
    for (int t = 0; t < 2; ++t) {
        clip_vtx tri[3] = {cv[tri_idx[t][0]], cv[tri_idx[t][1]], cv[tri_idx[t][2]]};
        if (tg->barycentric) render_triangle_from_clip_textured(tri, bt, light, tg); /* render_tiny_quad :932-1071 (same tables) */
        else render_triangle_span_from_clip(tri, bt, light, tg);
    }
}

/* rasterizer.rs:782-929 */
static void render_mesh_tiny_quads(const vxo_mesh_batch *mb, int32_t id, const float vp[16], target_t *tg) {
    if (!mb->has_mesh[id]) return;
    const uint32_t *so = mb->slice_offsets + (size_t)id * 6 * 33;
    const uint8_t *quads = mb->quads + 3 * (size_t)mb->quad_base[id];
    float off[3];
    for (int k = 0; k < 3; ++k) off[k] = (float)(mb->positions[3 * (size_t)id + k] * CS);

    for (int face = 0; face < 6; ++face) {
        if (so[face * 33 + 32] == so[face * 33]) continue; /* face_list.is_empty() */
        const int32_t *mn = mb->face_aabb + (size_t)id * 36 + face * 6, *mx = mn + 3;
        if (mn[0] > mx[0] || mn[1] > mx[1] || mn[2] > mx[2]) continue;
        float wmin[3], wmax[3];
        for (int k = 0; k < 3; ++k) { wmin[k] = off[k] + (float)mn[k]; wmax[k] = off[k] + (float)mx[k]; }
        int rect_min_x = INT32_MAX, rect_min_y = INT32_MAX, rect_max_x = INT32_MIN, rect_max_y = INT32_MIN;
        int any_behind = 0;
        for (int c = 0; c < 8; ++c) { /* corner order :832-841 */
            float cx = (c & 1) ? wmax[0] : wmin[0], cy = (c & 2) ? wmax[1] : wmin[1], cz = (c & 4) ? wmax[2] : wmin[2];
            float clip[4];
            mat4_mul_vec4(vp, cx, cy, cz, 1.0f, clip);
            if (clip[3] < 0.001f) any_behind = 1;
            if (fabsf(clip[3]) > 1e-4f) {
                float nx = clip[0] / clip[3], ny = clip[1] / clip[3];
                float sx = (nx + 1.0f) * 0.5f * (float)tg->W;
                float sy = (1.0f - ny) * 0.5f * (float)tg->H;
                rect_min_x = imin(rect_min_x, f2i(floorf(sx)));
                rect_max_x = imax(rect_max_x, f2i(ceilf(sx)));
                rect_min_y = imin(rect_min_y, f2i(floorf(sy)));
                rect_max_y = imax(rect_max_y, f2i(ceilf(sy)));
            }
        }
        if (!any_behind) {
            int tx0 = tg->rx0, ty0 = tg->ry0, tx1 = tg->rx0 + tg->rw - 1, ty1 = tg->ry0 + tg->rh - 1;
            if (rect_max_x < tx0 || rect_min_x > tx1 || rect_max_y < ty0 || rect_min_y > ty1) continue;
        }
        float light = vxo_face_light(tg->cfg, face);
        int positive = (face & 1) == 0;
        for (int slice = 0; slice < CS; ++slice) {
            uint32_t a = so[face * 33 + slice], b = so[face * 33 + slice + 1];
            uint8_t slice_pos = (uint8_t)(positive ? slice + 1 : slice); /* :896-900 */
            for (uint32_t qi = a; qi < b; ++qi) render_tiny_quad_span(quads + 3 * (size_t)qi, face, slice_pos, off, light, vp, tg);
        }
    }
}

void vxo_render_mesh(const vxo_mesh_batch *mb, int32_t mesh_id, const float vp[16], const vxo_frame_config *cfg,
                     const vxo_atlas *atlas, const int32_t rect[4], uint32_t *color, float *depth) {
    target_t tg = {cfg->width, cfg->height, rect[0], rect[1], rect[2], rect[3], color, depth, cfg, atlas, 0};
    render_mesh_tiny_quads(mb, mesh_id, vp, &tg);
}

/* Rasterizer::render_mesh_tiny_quads(mesh, vp, target, use_span_renderer) rasterizer.rs:782-929 */
void vxo_render_mesh_tiny_quads(const vxo_mesh_batch *mb, int32_t mesh_id, const float vp[16], const vxo_frame_config *cfg,
                                const vxo_atlas *atlas, const int32_t rect[4], int32_t use_span_renderer, uint32_t *color,
                                float *depth) {
    target_t tg = {cfg->width, cfg->height, rect[0], rect[1], rect[2], rect[3], color, depth, cfg, atlas, use_span_renderer ? 0 : 1};
    render_mesh_tiny_quads(mb, mesh_id, vp, &tg);
}

/* Rasterizer::render_mesh_with_up rasterizer.rs:399-411 + is_camera_level :377-382: whole framebuffer; the span
 * renderer only when the camera is level (|up.y| >= 0.995). */
void vxo_render_mesh_with_up(const vxo_mesh_batch *mb, int32_t mesh_id, const float vp[16], const vxo_frame_config *cfg,
                             const vxo_atlas *atlas, const float camera_up[3], uint32_t *color, float *depth) {
    const int level = fabsf(camera_up[1]) >= 0.995f;
    target_t tg = {cfg->width, cfg->height, 0, 0, cfg->width, cfg->height, color, depth, cfg, atlas, level ? 0 : 1};
    render_mesh_tiny_quads(mb, mesh_id, vp, &tg);
}

typedef struct {
    const vxo_mesh_batch *mb; const int32_t *mesh_ids; const float *vp; const vxo_frame_config *cfg;
    const vxo_atlas *atlas; uint32_t *color; float *depth;
    const int32_t *proj, *ord2, *rect_y; int32_t n_proj; int stripe_count, stripe_h;
    atomic_int next;
} stripe_job;

static void *stripe_worker(void *arg) {
    stripe_job *j = (stripe_job *)arg;
    const int W = j->cfg->width, H = j->cfg->height;
    for (;;) {
        int s = atomic_fetch_add(&j->next, 1);
        if (s >= j->stripe_count) break;
        int y0 = s * j->stripe_h; /* split_into_stripes framebuffer.rs:392-431 */
        if (y0 >= H) continue;
        int rows = imin(H - y0, j->stripe_h);
        target_t tg = {W, H, 0, y0, W, rows, j->color, j->depth, j->cfg, j->atlas, 0};
        for (int32_t si = 0; si < j->n_proj; ++si) {
            int32_t pj = j->ord2[si];
            int start = imin(j->rect_y[2 * pj] / j->stripe_h, j->stripe_count - 1);
            int end = imin(j->rect_y[2 * pj + 1] / j->stripe_h, j->stripe_count - 1);
            if (s < start || s > end) continue;
            render_mesh_tiny_quads(j->mb, j->mesh_ids[j->proj[pj]], j->vp, &tg);
        }
    }
    return NULL;
}

/* OcclusionBuffer occlusion.rs:6-154 (the parts render_frame uses) */
typedef struct { int sw, sh, gw, gh; float *cells; } occlusion_buffer;

static int occ_clamp_rect(const occlusion_buffer *o, int *min_x, int *min_y, int *max_x, int *max_y) {
    if (o->sw == 0 || o->sh == 0) return 0;
    if (*max_x < 0 || *max_y < 0 || *min_x >= o->sw || *min_y >= o->sh) return 0;
    *min_x = imax(*min_x, 0); *min_y = imax(*min_y, 0);
    *max_x = imin(*max_x, o->sw - 1); *max_y = imin(*max_y, o->sh - 1);
    return !(*min_x > *max_x || *min_y > *max_y);
}

/* occlusion.rs:60-99 */
static void occ_mark_rect(occlusion_buffer *o, int min_x, int min_y, int max_x, int max_y, float depth) {
    if (!occ_clamp_rect(o, &min_x, &min_y, &max_x, &max_y)) return;
    int cx0 = (int)(((int64_t)min_x * o->gw) / o->sw), cx1 = (int)(((int64_t)max_x * o->gw) / o->sw);
    int cy0 = (int)(((int64_t)min_y * o->gh) / o->sh), cy1 = (int)(((int64_t)max_y * o->gh) / o->sh);
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            float *cell = &o->cells[cy * o->gw + cx];
            if (depth < *cell) *cell = depth;
        }
}

/* occlusion.rs:105-153 */
static int occ_is_occluded(const occlusion_buffer *o, int min_x, int min_y, int max_x, int max_y, float near_depth) {
    if (!occ_clamp_rect(o, &min_x, &min_y, &max_x, &max_y)) return 0;
    int cx0 = (int)(((int64_t)min_x * o->gw) / o->sw), cx1 = (int)(((int64_t)max_x * o->gw) / o->sw);
    int cy0 = (int)(((int64_t)min_y * o->gh) / o->sh), cy1 = (int)(((int64_t)max_y * o->gh) / o->sh);
    const float epsilon = 0.005f;
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx)
            if (!(o->cells[cy * o->gw + cx] < near_depth - epsilon)) return 0;
    return 1;
}

/* main.rs:283-297, :368-377, :379-608 (occlusion pass :501-526 when cfg->occlusion_culling, off by default as main.rs:112) */
int vxo_render_frame(const vxo_mesh_batch *mb, const int32_t *mesh_ids, int32_t n_meshes, const float vp[16],
                     const float cam_pos[3], const vxo_frame_config *cfg, const vxo_atlas *atlas,
                     uint32_t *color, float *depth, int32_t *survivors_out) {
    const int W = cfg->width, H = cfg->height;
    const size_t npx = (size_t)W * (size_t)H;
    for (size_t i = 0; i < npx; ++i) { color[i] = cfg->clear_color; depth[i] = INFINITY; } /* framebuffer.rs:219 */
    if (n_meshes <= 0) return 0;

    /* VisibleMesh{center, distance_sq} main.rs:283-297 */
    float *center = (float *)malloc(sizeof(float) * 3 * (size_t)n_meshes);
    float *dist_sq = (float *)malloc(sizeof(float) * (size_t)n_meshes);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_meshes);
    for (int32_t i = 0; i < n_meshes; ++i) {
        const int32_t *p = mb->positions + 3 * (size_t)mesh_ids[i];
        float d[3];
        for (int k = 0; k < 3; ++k) {
            float mn = (float)(p[k] * CS);
            float mx = mn + (float)CS;
            center[3 * i + k] = (mn + mx) * 0.5f;
            d[k] = center[3 * i + k] - cam_pos[k];
        }
        dist_sq[i] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        order[i] = i;
    }
    stable_sort_by_key(order, n_meshes, dist_sq); /* main.rs:368-377 */

    /* 1. projection pass main.rs:405-490 (filter B), order preserved */
    const float half_size = (float)CS * 0.5f;
    const float width = (float)W, height = (float)H;
    float *near_depth = (float *)malloc(sizeof(float) * (size_t)n_meshes);
    int32_t *proj = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_meshes);
    int32_t *rect_y = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n_meshes);
    int32_t *rect_x = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n_meshes);
    uint8_t *use_occ = (uint8_t *)malloc((size_t)n_meshes);
    int32_t n_proj = 0;
    for (int32_t oi = 0; oi < n_meshes; ++oi) {
        int32_t i = order[oi];
        float mn[3], mx[3];
        for (int k = 0; k < 3; ++k) { mn[k] = center[3 * i + k] - half_size; mx[k] = center[3 * i + k] + half_size; }
        int rect_min_x = INT32_MAX, rect_min_y = INT32_MAX, rect_max_x = INT32_MIN, rect_max_y = INT32_MIN;
        float nd = INFINITY;
        int any_behind = 0;
        for (int c = 0; c < 8; ++c) {
            float cx = (c & 1) ? mx[0] : mn[0], cy = (c & 2) ? mx[1] : mn[1], cz = (c & 4) ? mx[2] : mn[2];
            float clip[4];
            mat4_mul_vec4(vp, cx, cy, cz, 1.0f, clip);
            if (clip[3] <= 0.001f) any_behind = 1;
            if (clip[3] > 0.001f) {
                float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
                nd = rmin(nd, nz);
                float sx = (nx + 1.0f) * 0.5f * width;
                float sy = (1.0f - ny) * 0.5f * height;
                rect_min_x = imin(rect_min_x, f2i(floorf(sx)));
                rect_max_x = imax(rect_max_x, f2i(ceilf(sx)));
                rect_min_y = imin(rect_min_y, f2i(floorf(sy)));
                rect_max_y = imax(rect_max_y, f2i(ceilf(sy)));
            }
        }
        if (any_behind) {
            rect_min_x = 0; rect_min_y = 0; rect_max_x = f2i(width) - 1; rect_max_y = f2i(height) - 1;
            nd = 0.0f;
        } else {
            if (isinf(nd) || nd > 1.0f) continue;
            rect_min_x = imax(rect_min_x, 0); rect_min_y = imax(rect_min_y, 0);
            rect_max_x = imin(rect_max_x, f2i(width) - 1); rect_max_y = imin(rect_max_y, f2i(height) - 1);
            if (rect_min_x > rect_max_x || rect_min_y > rect_max_y) continue;
        }
        near_depth[n_proj] = nd;
        rect_y[2 * n_proj] = rect_min_y; rect_y[2 * n_proj + 1] = rect_max_y;
        rect_x[2 * n_proj] = rect_min_x; rect_x[2 * n_proj + 1] = rect_max_x;
        { /* main.rs:473-478: OCCLUSION_MIN_DISTANCE_CHUNKS = 2 */
            const float t = (float)CS * 2.0f;
            use_occ[n_proj] = (uint8_t)(cfg->occlusion_culling && dist_sq[i] >= t * t);
        }
        proj[n_proj] = i;
        n_proj++;
    }
    /* sort front-to-back by near_depth main.rs:494-498 (stable) */
    int32_t *ord2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_proj > 0 ? n_proj : 1));
    for (int32_t i = 0; i < n_proj; ++i) ord2[i] = i;
    stable_sort_by_key(ord2, n_proj, near_depth);
    if (cfg->occlusion_culling) { /* 2. occlusion pass main.rs:501-526, serial front to back */
        occlusion_buffer occ = {W, H, cfg->occlusion_grid_w, cfg->occlusion_grid_h, NULL};
        const size_t nc = (size_t)imax(occ.gw, 0) * (size_t)imax(occ.gh, 0);
        occ.cells = (float *)malloc(sizeof(float) * (nc ? nc : 1));
        for (size_t i = 0; i < nc; ++i) occ.cells[i] = INFINITY; /* occlusion.clear() main.rs:394 */
        int32_t kept = 0;
        for (int32_t i = 0; i < n_proj; ++i) {
            const int32_t pj = ord2[i];
            if (use_occ[pj] && occ_is_occluded(&occ, rect_x[2 * pj], rect_y[2 * pj], rect_x[2 * pj + 1], rect_y[2 * pj + 1], near_depth[pj])) continue;
            occ_mark_rect(&occ, rect_x[2 * pj], rect_y[2 * pj], rect_x[2 * pj + 1], rect_y[2 * pj + 1], near_depth[pj]);
            ord2[kept++] = pj;
        }
        n_proj = kept;
        free(occ.cells);
    }
    for (int32_t i = 0; i < n_proj; ++i) survivors_out[i] = mesh_ids[proj[ord2[i]]];

    /* 3. stripe binning main.rs:528-557, 4. stripe rendering :559-597.  Rayon's
     * work-stealing for_each is restated as a pthread pool pulling stripes from an
     * atomic counter (stripes are disjoint row ranges, so scheduling can not change
     * the result). */
    int thread_count = cfg->n_threads > 0 ? cfg->n_threads : 1;
    stripe_job job;
    job.mb = mb; job.mesh_ids = mesh_ids; job.vp = vp; job.cfg = cfg; job.atlas = atlas;
    job.color = color; job.depth = depth; job.proj = proj; job.ord2 = ord2; job.rect_y = rect_y;
    job.n_proj = n_proj;
    job.stripe_count = thread_count * 4;
    job.stripe_h = (H + job.stripe_count - 1) / job.stripe_count;
    atomic_init(&job.next, 0);
    if (thread_count == 1) {
        stripe_worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)thread_count);
        for (int t = 0; t < thread_count; ++t) pthread_create(&th[t], NULL, stripe_worker, &job);
        for (int t = 0; t < thread_count; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    free(center); free(dist_sq); free(order); free(near_depth); free(proj); free(rect_y); free(rect_x); free(use_occ); free(ord2);
    return n_proj;
}

/* ---------- hyper-pipeline pieces ------------------------------------- */

/* differential_projection.rs:231-290 + :37-62 */
void vxo_face_basis(int face, const int32_t chunk_pos[3], uint8_t slice_idx, const float vp[16], float basis[16]) {
    float cw[3];
    for (int k = 0; k < 3; ++k) cw[k] = (float)chunk_pos[k] * (float)CS;
    float s = (float)slice_idx;
    float o[3] = {cw[0], cw[1], cw[2]}, t[3] = {0, 0, 0}, b[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    switch (face) {
    case 0: o[0] = cw[0] + s; o[1] = cw[1] + 0.0f; o[2] = cw[2] + 0.0f; t[1] = 1; b[2] = 1; n[0] = 1; break;
    case 1: o[0] = cw[0] + s; o[1] = cw[1] + 0.0f; o[2] = cw[2] + 0.0f; t[1] = 1; b[2] = -1; n[0] = -1; break;
    case 2: o[0] = cw[0] + 0.0f; o[1] = cw[1] + s; o[2] = cw[2] + 0.0f; t[0] = 1; b[2] = 1; n[1] = 1; break;
    case 3: o[0] = cw[0] + 0.0f; o[1] = cw[1] + s; o[2] = cw[2] + 0.0f; t[0] = 1; b[2] = -1; n[1] = -1; break;
    case 4: o[0] = cw[0] + 0.0f; o[1] = cw[1] + 0.0f; o[2] = cw[2] + s; t[0] = 1; b[1] = 1; n[2] = 1; break;
    default: o[0] = cw[0] + 0.0f; o[1] = cw[1] + 0.0f; o[2] = cw[2] + s; t[0] = -1; b[1] = 1; n[2] = -1; break;
    }
    mat4_mul_vec4(vp, o[0], o[1], o[2], 1.0f, basis + 0);
    mat4_mul_vec4(vp, t[0], t[1], t[2], 0.0f, basis + 4);
    mat4_mul_vec4(vp, b[0], b[1], b[2], 0.0f, basis + 8);
    mat4_mul_vec4(vp, n[0], n[1], n[2], 0.0f, basis + 12);
}

/* differential_projection.rs:69-71: origin + u*tangent + v*bitangent (left to right) */
void vxo_basis_project_point(const float basis[16], float u, float v, float out[4]) {
    for (int k = 0; k < 4; ++k) out[k] = (basis[k] + u * basis[4 + k]) + v * basis[8 + k];
}

/* differential_projection.rs:167-196 + perspective_divide :412-414 */
void vxo_project_packet(const float basis[16], const uint8_t *u_min, const uint8_t *v_min, const uint8_t *u_len,
                        const uint8_t *v_len, int n, float *x_min, float *y_min, float *x_max, float *y_max,
                        float *depth_near) {
    for (int i = 0; i < n; ++i) {
        float u0 = (float)u_min[i], v0 = (float)v_min[i];
        float u1 = u0 + (float)u_len[i], v1 = v0 + (float)v_len[i];
        float c[4][4], nd[4][3];
        vxo_basis_project_point(basis, u0, v0, c[0]);
        vxo_basis_project_point(basis, u1, v0, c[1]);
        vxo_basis_project_point(basis, u0, v1, c[2]);
        vxo_basis_project_point(basis, u1, v1, c[3]);
        for (int k = 0; k < 4; ++k)
            for (int j = 0; j < 3; ++j) nd[k][j] = c[k][j] / c[k][3];
        x_min[i] = rmin(rmin(rmin(nd[0][0], nd[1][0]), nd[2][0]), nd[3][0]);
        y_min[i] = rmin(rmin(rmin(nd[0][1], nd[1][1]), nd[2][1]), nd[3][1]);
        x_max[i] = rmax(rmax(rmax(nd[0][0], nd[1][0]), nd[2][0]), nd[3][0]);
        y_max[i] = rmax(rmax(rmax(nd[0][1], nd[1][1]), nd[2][1]), nd[3][1]);
        depth_near[i] = rmin(rmin(rmin(nd[0][2], nd[1][2]), nd[2][2]), nd[3][2]);
    }
}

/* simd_vertex.rs:48-58 with Vertex::world_position mesh.rs:106-112 */
void vxo_transform_vertices(const uint8_t *verts, int32_t n, const float offset[3], const float vp[16], float *out4) {
    for (int32_t i = 0; i < n; ++i) {
        const uint8_t *v = verts + 8 * (size_t)i;
        mat4_mul_vec4(vp, offset[0] + (float)v[0], offset[1] + (float)v[1], offset[2] + (float)v[2], 1.0f, out4 + 4 * (size_t)i);
    }
}

/* ---------- adjacent rasterizers -------------------- */

/* SpanWalkerRasterizer::get_block_color span_walker.rs:386-396 (BlockType::from_u8 block_type.rs:70-78) */
uint32_t vxo_span_walker_block_color(uint8_t block_type) {
    switch (block_type) {
    case 1: return 0x00FF00FFu;
    case 2: return 0x8B4513FFu;
    case 3: return 0x808080FFu;
    default: return 0x00000000u;
    }
}

/* FrameSlice::fill_span span_walker.rs:412-441 on a full-frame slice of `width` columns. */
void vxo_fill_span(int32_t width, int32_t y, int32_t x_start, int32_t x_end, float depth, uint32_t color,
                   uint32_t *cbuf, float *dbuf) {
    x_start = imin(imax(x_start, 0), width - 1);
    x_end = imin(imax(x_end, 0), width);
    if (x_start >= x_end) return;
    size_t row = (size_t)y * (size_t)width;
    for (int32_t x = x_start; x < x_end; ++x) {
        size_t idx = row + (size_t)x;
        if (depth < dbuf[idx]) {
            dbuf[idx] = depth;
            cbuf[idx] = color;
        }
    }
}

typedef struct { /* TrapezoidBatch span_walker.rs:20-52 */
    int count;
    float left_x[8], right_x[8], left_slope[8], right_slope[8], start_y[8], end_y[8], depth[8];
    uint32_t color[8];
    unsigned active_mask;
} trapezoid_batch;

/* rasterize_batch_scalar span_walker.rs:213-283 */
static void span_walker_batch(trapezoid_batch *b, int32_t width, int32_t height, uint32_t *cbuf, float *dbuf) {
    if (b->count == 0) return;
    float min_y = INFINITY, max_y = -INFINITY;
    for (int i = 0; i < b->count; ++i) {
        min_y = rmin(min_y, b->start_y[i]);
        max_y = rmax(max_y, b->end_y[i]);
    }
    int32_t current_y = f2i(floorf(min_y));
    const int32_t end_y = f2i(ceilf(max_y));
    while (current_y < end_y) {
        if (current_y >= height) break;
        if (current_y >= 0) {
            const float yc = (float)current_y + 0.5f; /* update_active_mask :76-84 */
            unsigned mask = 0;
            for (int i = 0; i < b->count; ++i)
                if (yc >= b->start_y[i] && yc < b->end_y[i]) mask |= 1u << i;
            b->active_mask = mask;
            if (mask) {
                for (int i = 0; i < b->count; ++i) {
                    if (!(mask & (1u << i))) continue;
                    int32_t xs = f2i(roundf(b->left_x[i])), xe = f2i(roundf(b->right_x[i]));
                    vxo_fill_span(width, current_y, xs, xe, b->depth[i], b->color[i], cbuf, dbuf);
                }
            }
        }
        for (int i = 0; i < b->count; ++i) {
            b->left_x[i] += b->left_slope[i];
            b->right_x[i] += b->right_slope[i];
        }
        current_y += 1;
    }
}

/* SpanWalkerRasterizer::rasterize_projected_packet span_walker.rs:116-194: one ProjectedPacket of `count` quads
 * (NDC boxes, differential_projection.rs:295-304) into a width x height framebuffer. */
void vxo_span_walk_packet(const float *x_min, const float *y_min, const float *x_max, const float *y_max,
                          const float *depth_near, const uint8_t *block_type, uint32_t visibility_mask, int32_t count,
                          int32_t width, int32_t height, uint32_t *cbuf, float *dbuf) {
    const float vp_w = (float)width, vp_h = (float)height;
    const float EPSILON = 0.001f;
    trapezoid_batch cur;
    memset(&cur, 0, sizeof(cur));
    for (int32_t i = 0; i < count && i < 32; ++i) {
        if (!(visibility_mask & (1u << i))) continue;
        float sx_min = rmax((x_min[i] + 1.0f) * 0.5f * vp_w, 0.0f);
        float sy_min = rmax((1.0f - y_max[i]) * 0.5f * vp_h, 0.0f);
        float sx_max = rmin((x_max[i] + 1.0f) * 0.5f * vp_w + EPSILON, vp_w);
        float sy_max = rmin((1.0f - y_min[i]) * 0.5f * vp_h + EPSILON, vp_h);
        if (sx_min >= vp_w || sy_min >= vp_h || sx_max <= 0.0f || sy_max <= 0.0f) continue;
        int k = cur.count;
        cur.left_x[k] = sx_min; cur.right_x[k] = sx_max;
        cur.left_slope[k] = 0.0f; cur.right_slope[k] = 0.0f;
        cur.start_y[k] = sy_min; cur.end_y[k] = sy_max;
        cur.depth[k] = depth_near[i];
        cur.color[k] = vxo_span_walker_block_color(block_type[i]);
        cur.active_mask |= 1u << k;
        cur.count++;
        if (cur.count == 8) {
            span_walker_batch(&cur, width, height, cbuf, dbuf);
            memset(&cur, 0, sizeof(cur));
        }
    }
    if (cur.count > 0) span_walker_batch(&cur, width, height, cbuf, dbuf);
}

/* MacroTileBins::add_mesh macrotile.rs:179-224 for one screen box: returns 1 when the mesh is binned into the tile
 * range tiles[4] = (tx0, ty0, tx1, ty1) (inclusive), 2 for a large primitive (> 25 % of the screen, not binned),
 * 0 off-screen. */
int vxo_macrotile_bin(int32_t min_x_in, int32_t min_y_in, int32_t max_x_in, int32_t max_y_in, int32_t fb_w, int32_t fb_h,
                      int32_t tiles[4]) {
    const int64_t min_x = imax(min_x_in, 0), min_y = imax(min_y_in, 0);
    const int64_t max_x = imin(max_x_in, fb_w - 1), max_y = imin(max_y_in, fb_h - 1);
    if (min_x > max_x || min_y > max_y) return 0;
    const int64_t coverage = (max_x - min_x + 1) * (max_y - min_y + 1);
    const int64_t total = (int64_t)fb_w * (int64_t)fb_h;
    const float fraction = (float)coverage / (float)total;
    if (fraction > 0.25f) return 2; /* LARGE_PRIMITIVE_SCREEN_FRACTION :26 */
    const int tiles_x = (fb_w + 127) / 128, tiles_y = (fb_h + 127) / 128;
    tiles[0] = (int32_t)(min_x / 128);
    tiles[1] = (int32_t)(min_y / 128);
    tiles[2] = imin((int32_t)(max_x / 128), tiles_x - 1);
    tiles[3] = imin((int32_t)(max_y / 128), tiles_y - 1);
    return 1;
}

/* project_mesh_aabb macrotile_renderer.rs:175-250 (the same arithmetic as main.rs:405-470) */
static int project_mesh_aabb(const int32_t pos[3], const float vp[16], float width, float height, int32_t rect[4]) {
    float mn[3], mx[3];
    for (int k = 0; k < 3; ++k) {
        float chunk_pos = (float)pos[k] * (float)CS;
        float half = (float)CS * 0.5f;
        float center = chunk_pos + half;
        mn[k] = center - half;
        mx[k] = center + half;
    }
    int rect_min_x = INT32_MAX, rect_min_y = INT32_MAX, rect_max_x = INT32_MIN, rect_max_y = INT32_MIN;
    float nd = INFINITY;
    int any_behind = 0;
    for (int c = 0; c < 8; ++c) {
        float cx = (c & 1) ? mx[0] : mn[0], cy = (c & 2) ? mx[1] : mn[1], cz = (c & 4) ? mx[2] : mn[2];
        float clip[4];
        mat4_mul_vec4(vp, cx, cy, cz, 1.0f, clip);
        if (clip[3] <= 0.001f) any_behind = 1;
        if (clip[3] > 0.001f) {
            float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
            nd = rmin(nd, nz);
            float sx = (nx + 1.0f) * 0.5f * width;
            float sy = (1.0f - ny) * 0.5f * height;
            rect_min_x = imin(rect_min_x, f2i(floorf(sx)));
            rect_max_x = imax(rect_max_x, f2i(ceilf(sx)));
            rect_min_y = imin(rect_min_y, f2i(floorf(sy)));
            rect_max_y = imax(rect_max_y, f2i(ceilf(sy)));
        }
    }
    if (any_behind) {
        rect_min_x = 0; rect_min_y = 0; rect_max_x = f2i(width) - 1; rect_max_y = f2i(height) - 1;
    } else {
        if (isinf(nd) || nd > 1.0f) return 0;
        rect_min_x = imax(rect_min_x, 0); rect_min_y = imax(rect_min_y, 0);
        rect_max_x = imin(rect_max_x, f2i(width) - 1); rect_max_y = imin(rect_max_y, f2i(height) - 1);
        if (rect_min_x > rect_max_x || rect_min_y > rect_max_y) return 0;
    }
    rect[0] = rect_min_x; rect[1] = rect_min_y; rect[2] = rect_max_x; rect[3] = rect_max_y;
    return 1;
}

/* render_frame_macrotile macrotile_renderer.rs:51-170: clear, project every mesh's chunk box, bin into 128x128
 * macrotiles (meshes covering > 25 % of the screen go to the large-primitive list), per tile render the binned meshes
 * in list order and then the large primitives through render_mesh_tiny_quads with the tile as PixelTarget, flush the
 * tile colours.  The Hi-Z buffer argument of the reference is only cleared there, never consulted.  Rayon's tile
 * parallelism cannot change the result (tiles are disjoint).
 *   color       W*H, the framebuffer colour the reference flushes
 *   tile_depth  W*H or NULL: the tiles' depth buffers (the reference drops them; the framebuffer's own depth buffer
 *               stays at +inf) -- a diagnostic for the tests
 *   projected_out  mesh ids that survived project_mesh_aabb, list order;  returns their number (the reference's
 *               return value)
 *   kind_out    per projected mesh: 1 binned, 2 large primitive (drawn after the binned ones in every tile) */
int vxo_render_frame_macrotile(const vxo_mesh_batch *mb, const int32_t *mesh_ids, int32_t n_meshes, const float vp[16],
                               const vxo_frame_config *cfg, const vxo_atlas *atlas, uint32_t *color, float *tile_depth,
                               int32_t *projected_out, int32_t *kind_out) {
    const int W = cfg->width, H = cfg->height;
    const size_t npx = (size_t)W * (size_t)H;
    float *depth = tile_depth ? tile_depth : (float *)malloc(sizeof(float) * npx);
    for (size_t i = 0; i < npx; ++i) { color[i] = cfg->clear_color; depth[i] = INFINITY; }
    int32_t n_proj = 0;
    int32_t *proj = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_meshes > 0 ? n_meshes : 1));
    int32_t *kind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_meshes > 0 ? n_meshes : 1));
    int32_t *tiles = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(n_meshes > 0 ? n_meshes : 1));
    int n_large = 0;
    for (int32_t i = 0; i < n_meshes; ++i) {
        int32_t rect[4];
        if (!project_mesh_aabb(mb->positions + 3 * (size_t)mesh_ids[i], vp, (float)W, (float)H, rect)) continue;
        proj[n_proj] = mesh_ids[i];
        kind[n_proj] = vxo_macrotile_bin(rect[0], rect[1], rect[2], rect[3], W, H, tiles + 4 * (size_t)n_proj);
        if (kind[n_proj] == 2) n_large++;
        if (projected_out) projected_out[n_proj] = mesh_ids[i];
        if (kind_out) kind_out[n_proj] = kind[n_proj];
        n_proj++;
    }
    if (n_proj > 0) {
        const int tiles_x = (W + 127) / 128, tiles_y = (H + 127) / 128;
        for (int ty = 0; ty < tiles_y; ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                const int x0 = tx * 128, y0 = ty * 128;
                target_t tg = {W, H, x0, y0, imin(x0 + 128, W) - x0, imin(y0 + 128, H) - y0, color, depth, cfg, atlas, 0};
                for (int32_t j = 0; j < n_proj; ++j) /* bins.get_bin(tx, ty), push order = list order */
                    if (kind[j] == 1 && tx >= tiles[4 * j] && tx <= tiles[4 * j + 2] && ty >= tiles[4 * j + 1] && ty <= tiles[4 * j + 3])
                        render_mesh_tiny_quads(mb, proj[j], vp, &tg);
                if (n_large)
                    for (int32_t j = 0; j < n_proj; ++j)
                        if (kind[j] == 2) render_mesh_tiny_quads(mb, proj[j], vp, &tg);
            }
    }
    free(proj); free(kind); free(tiles);
    if (!tile_depth) free(depth);
    return n_proj;
}

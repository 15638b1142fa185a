/*
 * vx_oracle.h -- CPU oracle for the per-frame voxel pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * crate's CPU algorithm (gatewaytofredom/differential_projection_voxel_renderer,
 * `voxel_engine` 0.1.0) for the hot path: binary greedy meshing, frustum /
 * screen-rect culling, draw ordering, and the span (scanline) rasterizer.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library (libvx_b200.so)
 * never links, loads or calls anything in this directory.
 *
 * PARITY PINNING: the reference is a Rust crate and neither cargo nor rustc
 * exist in the build container, and its crates.io dependencies (glam 0.25.0,
 * noise 0.9.0, rayon 1.11.0) are not vendored, so the reference itself can not
 * be compiled or run here ("unbuildable": oracle/_ref is intentionally empty).
 * The oracle is pinned against every known-answer test the reference's own
 * test-suite holds for this path (see tests/test_oracle_kat.py; sources cited
 * there).  The reference ships no golden images / vectors, so full-frame
 * colour+depth values are pinned only by this restatement: for those the
 * header says it plainly -- *frame-level parity is unpinned beyond the KATs*.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef VX_ORACLE_H
#define VX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXO_CHUNK_SIZE 32
#define VXO_CHUNK_VOLUME 32768

/* neighbour codes in the `neighbors` table (N x 6, order +X,-X,+Y,-Y,+Z,-Z) */
#define VXO_NBR_NONE (-1)          /* no chunk there: faces exposed          */
#define VXO_NBR_UNIFORM_AIR (-2)   /* uniform non-solid neighbour             */
#define VXO_NBR_UNIFORM_SOLID (-3) /* uniform solid neighbour                 */

/* ---- meshing ---------------------------------------------------------- */

/* binary_greedy.rs:683-807.  quads_xywh receives (x=row, y=col, width, height)
 * 4 bytes per quad.  Returns the number of quads (<= 512). */
int vxo_greedy_mesh_slice(const uint32_t mask[32], uint8_t *quads_xywh);

/* mesh.rs:283-342 TinyQuad pack / unpack. */
void vxo_tinyquad_pack(uint8_t u, uint8_t v, uint8_t w, uint8_t h, uint8_t block_type, uint8_t out[3]);
void vxo_tinyquad_unpack(const uint8_t in[3], uint8_t *u, uint8_t *v, uint8_t *w, uint8_t *h, uint8_t *block_type);

/* binary_greedy.rs:83-121 (mesh_chunk_in_world) for one Varied chunk.
 *   voxels        32768 bytes, index = z*1024 + y*32 + x (chunk.rs:52)
 *   nbr_voxels[f] voxel array of the Varied neighbour in face direction f or NULL
 *   nbr_code[f]   used when nbr_voxels[f]==NULL: VXO_NBR_NONE / _UNIFORM_AIR / _UNIFORM_SOLID
 *   quads_out     3 bytes per quad, face -> slice -> type -> greedy order
 *   slice_offsets [6][33] quad index where list (face, slice) starts; [f][32] = end of face f
 *   face_aabb     [6][6] = min.xyz, max.xyz in 0..32 (FaceList::min/max, mesh.rs:347-397)
 * Returns the quad count, or -1 when `cap` quads is too small. */
int vxo_mesh_chunk(const uint8_t *voxels, const uint8_t *const nbr_voxels[6], const int32_t nbr_code[6],
                   uint8_t *quads_out, int cap, uint32_t slice_offsets[6 * 33], int32_t face_aabb[6 * 6]);

/* Batch form mirroring the product ABI (mesh_world, binary_greedy.rs:62-78).
 *   uniform_flags[i] == 0 : Varied chunk;  else 1 + block_type of a Uniform chunk
 *   neighbors[i*6+f] >= 0 : index into this batch;  else one of VXO_NBR_*
 *   quad_base[i]  first quad of chunk i inside quads_out; quad_count[i]; has_mesh[i]
 * Returns total quads, or -1 on overflow of `cap`. */
int64_t vxo_mesh_chunks(const uint8_t *voxels, const int32_t *neighbors, const uint8_t *uniform_flags,
                        int32_t n_chunks, uint8_t *quads_out, int64_t cap, uint32_t *quad_base,
                        uint32_t *quad_count, uint32_t *slice_offsets, int32_t *face_aabb, uint8_t *has_mesh);

/* ---- culling ---------------------------------------------------------- */

/* camera/mod.rs:123-160.  vp is column-major (glam Mat4::to_cols_array). */
/* terrain: noise 0.9.0 Perlin::new(seed) restated (see vx_oracle.c) + Chunk::generate_terrain chunk.rs:114-207 */
void vxo_noise_permutation_table(uint32_t seed, uint8_t values[256]);
double vxo_perlin2(const uint8_t values[256], double px, double py);
int32_t vxo_terrain_height(const uint8_t values[256], int32_t x, int32_t z);
int vxo_generate_terrain(const int32_t pos[3], uint32_t seed, uint8_t *voxels_out);

void vxo_frustum_from_vp(const float vp[16], float planes[24]);
/* camera/mod.rs:164-183 */
int vxo_frustum_intersects_aabb(const float planes[24], const float mn[3], const float mx[3]);
/* world.rs:118-146 filter A.  visible_out[i] in {0,1}. */
void vxo_cull_chunks(const int32_t *positions, int32_t n, const float vp[16], const float cam_pos[3],
                     int32_t view_distance, int32_t frustum_culling, uint8_t *visible_out);
/* culling.rs:40-119.  centers n x 3, order[] in/out (indices into centers; stable
 * distance sort is applied first).  Returns number kept. */
int vxo_horizon_cull(const float cam_pos[3], const float *centers, int32_t n, int32_t *order, int32_t bins,
                     float base_margin, float margin_dist_factor, float min_dist_chunks);

/* ---- rendering -------------------------------------------------------- */

typedef struct {
    uint32_t palette[4][16]; /* texture.rs:5-13 */
    uint8_t indices[4][32];
} vxo_atlas;

typedef struct {
    int32_t width, height;
    uint32_t clear_color;     /* main.rs:393 uses 0xFF87CEEB */
    int32_t backface_culling; /* rasterizer.rs:336 */
    int32_t enable_shading;   /* rasterizer.rs:340 */
    float light_dir[3];       /* shading.rs:21-31 */
    float ambient, diffuse;
    int32_t n_threads;        /* stripes = 4*n_threads, main.rs:531-534 */
    int32_t occlusion_culling; /* main.rs:112 (off), :501-526 */
    int32_t occlusion_grid_w, occlusion_grid_h; /* main.rs:46-47: 128 x 72 */
} vxo_frame_config;

typedef struct {
    const uint8_t *quads;          /* 3 bytes each */
    const uint32_t *quad_base;     /* [N] */
    const uint32_t *slice_offsets; /* [N][6][33] */
    const int32_t *face_aabb;      /* [N][6][6] */
    const int32_t *positions;      /* [N][3] chunk coordinates */
    const uint8_t *has_mesh;       /* [N] */
    int32_t n_chunks;
} vxo_mesh_batch;

void vxo_default_atlas(vxo_atlas *atlas);                    /* texture.rs:60-123 */
void vxo_default_frame_config(vxo_frame_config *cfg, int w, int h);
uint32_t vxo_shade_color_u32(uint32_t base, float light);    /* shading.rs:90-110 */
float vxo_face_light(const vxo_frame_config *cfg, int face); /* rasterizer.rs:1204-1216 */
uint32_t vxo_texture_sample(const vxo_atlas *a, int tex, uint8_t u, uint8_t v); /* texture.rs:19-38 */

/* Rasterizer::render_mesh_into_target -> render_mesh_tiny_quads(span)  rasterizer.rs:627,782.
 * rect = (x0,y0,w,h) of the PixelTarget inside the W x H framebuffer
 * (FrameSlice: x0 = 0, w = W;  FrameTile: arbitrary).  color/depth are W*H. */
void vxo_render_mesh(const vxo_mesh_batch *mb, int32_t mesh_id, const float vp[16], const vxo_frame_config *cfg,
                     const vxo_atlas *atlas, const int32_t rect[4], uint32_t *color, float *depth);

/* Rasterizer::render_mesh_tiny_quads rasterizer.rs:782-929 with the pub use_span_renderer switch: 0 = the barycentric
 * rasterizer render_triangle_from_clip_textured (:1881-2107). */
void vxo_render_mesh_tiny_quads(const vxo_mesh_batch *mb, int32_t mesh_id, const float vp[16], const vxo_frame_config *cfg,
                                const vxo_atlas *atlas, const int32_t rect[4], int32_t use_span_renderer, uint32_t *color,
                                float *depth);
/* Rasterizer::render_mesh_with_up rasterizer.rs:399-411 (span renderer iff |camera_up.y| >= 0.995, :377-382). */
void vxo_render_mesh_with_up(const vxo_mesh_batch *mb, int32_t mesh_id, const float vp[16], const vxo_frame_config *cfg,
                             const vxo_atlas *atlas, const float camera_up[3], uint32_t *color, float *depth);

/* main.rs:283-297 + :368-377 + render_frame :379-608 (occlusion off).
 *   mesh_ids: chunks that passed filter A and have a mesh, in caller order
 *   survivors_out: draw order after filter B + sorts;  returns survivor count
 *   color/depth are cleared first (framebuffer.rs:219). */
int vxo_render_frame(const vxo_mesh_batch *mb, const int32_t *mesh_ids, int32_t n_meshes, const float vp[16],
                     const float cam_pos[3], const vxo_frame_config *cfg, const vxo_atlas *atlas,
                     uint32_t *color, float *depth, int32_t *survivors_out);

/* ---- hyper-pipeline pieces ------------------------------------------- */

/* differential_projection.rs:37-62: origin, tangent, bitangent, normal (4 x vec4). */
void vxo_face_basis(int face, const int32_t chunk_pos[3], uint8_t slice_idx, const float vp[16], float basis[16]);
/* differential_projection.rs:69-71 */
void vxo_basis_project_point(const float basis[16], float u, float v, float out[4]);
/* differential_projection.rs:167-196 scalar path over a packet of n quads (SoA u8 arrays).
 * out = x_min[n], y_min[n], x_max[n], y_max[n], depth_near[n] (NDC). */
void vxo_project_packet(const float basis[16], const uint8_t *u_min, const uint8_t *v_min, const uint8_t *u_len,
                        const uint8_t *v_len, int n, float *x_min, float *y_min, float *x_max, float *y_max,
                        float *depth_near);
/* simd_vertex.rs:48-58 scalar path.  verts: 8 bytes each (x,y,z,...), out: n x 4 floats. */
void vxo_transform_vertices(const uint8_t *verts, int32_t n, const float offset[3], const float vp[16], float *out4);
/* the four clip-space corners of a TinyQuad exactly as rasterizer.rs:1092-1185 computes them */
void vxo_quad_clip_vertices(int face, uint8_t slice_pos, uint8_t u, uint8_t v, uint8_t w, uint8_t h,
                            const int32_t chunk_pos[3], const float vp[16], float clip[16]);


/* ---- adjacent rasterizers -------------------------- */

/* SpanWalkerRasterizer::get_block_color span_walker.rs:386-396 */
uint32_t vxo_span_walker_block_color(uint8_t block_type);
/* FrameSlice::fill_span span_walker.rs:412-441: pixels [x_start, x_end) of row y, depth test `<`. */
void vxo_fill_span(int32_t width, int32_t y, int32_t x_start, int32_t x_end, float depth, uint32_t color,
                   uint32_t *cbuf, float *dbuf);
/* SpanWalkerRasterizer::rasterize_projected_packet span_walker.rs:116-283 (scalar batch path) for one
 * ProjectedPacket (count <= 32 NDC boxes + constant depth + block type, visibility bit per quad). */
void vxo_span_walk_packet(const float *x_min, const float *y_min, const float *x_max, const float *y_max,
                          const float *depth_near, const uint8_t *block_type, uint32_t visibility_mask, int32_t count,
                          int32_t width, int32_t height, uint32_t *cbuf, float *dbuf);
/* MacroTileBins::add_mesh macrotile.rs:179-224: 1 binned (tiles = tx0,ty0,tx1,ty1 inclusive), 2 large primitive,
 * 0 off-screen. */
int vxo_macrotile_bin(int32_t min_x, int32_t min_y, int32_t max_x, int32_t max_y, int32_t fb_w, int32_t fb_h,
                      int32_t tiles[4]);
/* render_frame_macrotile macrotile_renderer.rs:51-170 (see vx_oracle.c). */
int vxo_render_frame_macrotile(const vxo_mesh_batch *mb, const int32_t *mesh_ids, int32_t n_meshes, const float vp[16],
                               const vxo_frame_config *cfg, const vxo_atlas *atlas, uint32_t *color, float *tile_depth,
                               int32_t *projected_out, int32_t *kind_out);

#ifdef __cplusplus
}
#endif
#endif

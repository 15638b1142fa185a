// vx_bary.cu -- the barycentric rasterizer of the mesh path:
//   Rasterizer::render_mesh_with_up            rasterizer.rs:399-411 (+ is_camera_level :377-382)
//   Rasterizer::render_mesh_tiny_quads(.., use_span_renderer = false)   :782-929
//   render_tiny_quad :932-1071, render_triangle_from_clip_textured :1881-2107
//
// The reference walks a triangle's clipped bounding box row by row and advances the three edge functions by one
// rounded f32 add per pixel / per row, so a pixel's coverage and depth depend on the whole chain from the box's
// top-left pixel.  Here every (triangle, row) is one warp task: the row's start values are the exact fast-forward
// of the per-row chain (vx_accum_jump), every lane enters the per-pixel chain the same way at its own piece of the
// row and walks it with real adds.  Depth test + colour write is the 64-bit key min of the frame path:
//   [ order-preserving depth : 32 | draw sequence : 23 | face : 3 | type : 2 | texel nibble : 4 ]
// with the pixel's previous depth entering as [depth | 0] (wins every tie, like `depth < stored` loses it).
#include "vx_common.cuh"
#include "vx_jump.h"
#include "vx_math.cuh"
#include "vx_scan.cuh"

// vx_frame.cu: the payload -> ARGB table (atlas palette x face light) and the atlas nibble indices of the context
int vx_frame_tables(VxContext *ctx, const VxFrameConfig &cfg, const uint32_t **d_lut, const uint8_t **d_tex_idx);

namespace {

constexpr int BY_THREADS = 256;

struct BaryRec { // one clipped triangle; slot = quad * 4 + triangle * 2 + clip piece = draw order
    int32_t min_x, max_x, min_y, max_y; // pixel box (min_y > max_y: nothing to draw)
    float w_row[3];                      // edge functions at the centre of pixel (min_x, min_y)
    float dx[3], dy[3];                  // their increments per pixel / per row
    float inv_area;
    float z[3], iw[3], uw[3], vw[3];
    uint32_t lo_base;                    // (slot << 9) | face << 6 | type << 4
    uint32_t pad;
};
static_assert(sizeof(BaryRec) == 112, "BaryRec layout");

struct BaryParams {
    VxMat4 vp;
    int32_t W, H, rx0, ry0, rw, rh;
    int32_t backface;
    uint32_t qbase, qcount;
    float off[3];
    const uint8_t *quads;
    const uint32_t *slice_offsets; // [6][33] of this mesh
};

struct ClipB {
    float4 p;
    float u, v;
};

// intersect_near_textured rasterizer.rs:2628-2641
__device__ __forceinline__ ClipB intersect_near_b(const ClipB &a, const ClipB &b) {
    const float t = (VX_NEAR_W_EPS - a.p.w) / (b.p.w - a.p.w);
    ClipB r;
    r.p.x = a.p.x + (b.p.x - a.p.x) * t;
    r.p.y = a.p.y + (b.p.y - a.p.y) * t;
    r.p.z = a.p.z + (b.p.z - a.p.z) * t;
    r.p.w = a.p.w + (b.p.w - a.p.w) * t;
    r.u = a.u + (b.u - a.u) * t;
    r.v = a.v + (b.v - a.v) * t;
    return r;
}

// Rasterizer::edge_function rasterizer.rs:2556-2558
__device__ __forceinline__ float edge_fn(float ax, float ay, float bx, float by, float cx, float cy) {
    return (cx - ax) * (by - ay) - (cy - ay) * (bx - ax);
}

// render_triangle_from_clip_textured :1928-2038 for one clipped triangle
__device__ __forceinline__ void bary_setup_triangle(const BaryParams &P, const ClipB &a, const ClipB &b, const ClipB &c, BaryRec &r) {
    r.min_y = 1; r.max_y = 0;
    const ClipB *tv[3] = {&a, &b, &c};
    float nx[3], ny[3], nz[3], px[3], py[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        nx[i] = tv[i]->p.x / tv[i]->p.w;
        ny[i] = tv[i]->p.y / tv[i]->p.w;
        nz[i] = tv[i]->p.z / tv[i]->p.w;
    }
    if (P.backface) { // :1943-1951
        const float v01x = nx[1] - nx[0], v01y = ny[1] - ny[0], v02x = nx[2] - nx[0], v02y = ny[2] - ny[0];
        const float cross_z = v01x * v02y - v01y * v02x;
        if (cross_z <= 0.0f) return;
    }
    const float fw = (float)P.W, fh = (float)P.H;
#pragma unroll
    for (int i = 0; i < 3; ++i) { // ndc_to_screen :2546-2551
        px[i] = (nx[i] + 1.0f) * 0.5f * fw;
        py[i] = (1.0f - ny[i]) * 0.5f * fh;
    }
    int min_x = vx_f2i(floorf(fminf(fminf(px[0], px[1]), px[2])));
    int max_x = vx_f2i(ceilf(fmaxf(fmaxf(px[0], px[1]), px[2])));
    int min_y = vx_f2i(floorf(fminf(fminf(py[0], py[1]), py[2])));
    int max_y = vx_f2i(ceilf(fmaxf(fmaxf(py[0], py[1]), py[2])));
    min_x = max(min_x, 0); max_x = min(max_x, P.W - 1); // :1968-1973
    min_y = max(min_y, 0); max_y = min(max_y, P.H - 1);
    min_x = max(min_x, P.rx0); max_x = min(max_x, P.rx0 + P.rw - 1); // :1976-1983
    min_y = max(min_y, P.ry0); max_y = min(max_y, P.ry0 + P.rh - 1);
    if (min_x > max_x || min_y > max_y) return;
    const float area = edge_fn(px[0], py[0], px[1], py[1], px[2], py[2]);
    if (area <= 0.0f) return;
    if (area < 0.1f) return; // MIN_TRIANGLE_AREA :1997-2001
    r.inv_area = 1.0f / area;
    r.dx[0] = py[2] - py[1]; r.dy[0] = px[1] - px[2]; // :2006-2011
    r.dx[1] = py[0] - py[2]; r.dy[1] = px[2] - px[0];
    r.dx[2] = py[1] - py[0]; r.dy[2] = px[0] - px[1];
    const ClipB *uvv[3] = {&a, &b, &c};
#pragma unroll
    for (int i = 0; i < 3; ++i) { // :2018-2028
        r.z[i] = nz[i];
        r.iw[i] = 1.0f / tv[i]->p.w;
        r.uw[i] = uvv[i]->u * r.iw[i];
        r.vw[i] = uvv[i]->v * r.iw[i];
    }
    const float sx = (float)min_x + 0.5f, sy = (float)min_y + 0.5f; // :2031-2037
    r.w_row[0] = edge_fn(px[1], py[1], px[2], py[2], sx, sy);
    r.w_row[1] = edge_fn(px[2], py[2], px[0], py[0], sx, sy);
    r.w_row[2] = edge_fn(px[0], py[0], px[1], py[1], sx, sy);
    r.min_x = min_x; r.max_x = max_x; r.min_y = min_y; r.max_y = max_y;
}

// one thread per quad: unpack, four clip-space corners (render_tiny_quad :932-1052), two triangles, near clip
// (clip_triangle_near_textured :2645-2697), up to four triangle records in draw order
__global__ void __launch_bounds__(BY_THREADS) bary_setup_kernel(BaryParams P, BaryRec *__restrict__ recs, uint32_t *__restrict__ n_tasks) {
    const uint32_t q = blockIdx.x * BY_THREADS + threadIdx.x;
    if (q >= P.qcount) return;
    const uint8_t *qp = P.quads + 3 * (size_t)(P.qbase + q);
    const uint32_t b0 = qp[0], b1 = qp[1], b2 = qp[2];
    int face = 0;
#pragma unroll
    for (int ff = 1; ff < 6; ++ff) face += (P.slice_offsets[ff * 33] <= q) ? 1 : 0;
    int lo = 0, hi = 31;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.slice_offsets[face * 33 + mid] <= q) lo = mid; else hi = mid - 1;
    }
    const int slice = lo, axis = face >> 1;
    const int spos = (face & 1) ? slice : slice + 1; // :896-900
    const int u = b0 & 0x1F, v = ((b0 >> 5) & 7) | ((b1 & 3) << 3);
    const int w = ((b1 >> 2) & 0x3F) + 1, h = (b2 & 0x3F) + 1;
    const uint32_t type = (b2 >> 6) & 3;
    ClipB cv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int cu = ((kCornerU[face] >> i) & 1) ? u + w : u;
        const int cvv = ((kCornerV[face] >> i) & 1) ? v + h : v;
        int lx, ly, lz;
        if (axis == 0) { lx = spos; ly = cu; lz = cvv; }
        else if (axis == 1) { lx = cu; ly = spos; lz = cvv; }
        else { lx = cu; ly = cvv; lz = spos; }
        cv[i].p = vx_mul_point(P.vp, P.off[0] + (float)lx, P.off[1] + (float)ly, P.off[2] + (float)lz); // :1041-1050
        cv[i].u = (float)cu;
        cv[i].v = (float)cvv;
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) { // tris (0,1,2), (0,2,3) :1053
        ClipB poly[4];
        int pn = 0;
        const ClipB *in[3] = {&cv[0], &cv[t == 0 ? 1 : 2], &cv[t == 0 ? 2 : 3]};
        const ClipB *prev = in[2];
        bool prev_in = prev->p.w >= VX_NEAR_W_EPS;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const ClipB *cur = in[i];
            const bool cur_in = cur->p.w >= VX_NEAR_W_EPS;
            if (prev_in && cur_in) poly[pn++] = *cur;
            else if (prev_in && !cur_in) poly[pn++] = intersect_near_b(*prev, *cur);
            else if (!prev_in && cur_in) {
                poly[pn++] = intersect_near_b(*prev, *cur);
                poly[pn++] = *cur;
            }
            prev = cur;
            prev_in = cur_in;
        }
#pragma unroll
        for (int piece = 0; piece < 2; ++piece) {
            const uint32_t slot = q * 4u + (uint32_t)(t * 2 + piece);
            BaryRec r;
            r.min_x = 0; r.max_x = -1; r.min_y = 1; r.max_y = 0;
            r.pad = 0;
            if (piece == 0 ? pn >= 3 : pn == 4) bary_setup_triangle(P, poly[0], poly[piece == 0 ? 1 : 2], poly[piece == 0 ? 2 : 3], r);
            r.lo_base = (slot << 9) | ((uint32_t)face << 6) | (type << 4);
            recs[slot] = r;
            n_tasks[slot] = r.min_y <= r.max_y ? (uint32_t)(r.max_y - r.min_y + 1) : 0u;
        }
    }
}

__global__ void __launch_bounds__(BY_THREADS) bary_init_keys_kernel(const float *__restrict__ depth, size_t npx, unsigned long long *__restrict__ keys) {
    for (size_t p = (size_t)blockIdx.x * BY_THREADS + threadIdx.x; p < npx; p += (size_t)gridDim.x * BY_THREADS) {
        const float d = depth[p];
        keys[p] = d == d ? ((unsigned long long)vx_ord(d + 0.0f) << 32) : 0ull; // a stored NaN rejects every fragment
    }
}

// one warp per (triangle, row)
__global__ void __launch_bounds__(BY_THREADS) bary_fill_kernel(const BaryRec *__restrict__ recs, const unsigned long long *__restrict__ task_base, int n_slots,
                                                             const unsigned long long *__restrict__ total_tasks, int W, const uint8_t *__restrict__ tex_idx,
                                                             unsigned long long *__restrict__ keys) {
    const unsigned long long total = *total_tasks;
    const int lane = threadIdx.x & 31;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * (BY_THREADS / 32);
    for (unsigned long long t = (unsigned long long)blockIdx.x * (BY_THREADS / 32) + (threadIdx.x >> 5); t < total; t += n_warps) {
        const int slot = vx_task_owner(task_base, n_slots, t);
        const BaryRec r = recs[slot];
        const uint32_t row = (uint32_t)(t - task_base[slot]);
        const int y = r.min_y + (int)row;
        // the row's start values: `row` steps of w_row += dy (:2099-2101); lanes 0..2 take one edge each
        float wr = 0.0f;
        if (lane < 3) wr = vx_accum_jump(r.w_row[lane], r.dy[lane], row);
        const float w0_row = __shfl_sync(0xffffffffu, wr, 0), w1_row = __shfl_sync(0xffffffffu, wr, 1), w2_row = __shfl_sync(0xffffffffu, wr, 2);
        const int width = r.max_x - r.min_x + 1;
        const int per = (width + 31) >> 5;
        const int n0 = lane * per;
        if (n0 < width) {
            const int n1 = min(n0 + per, width);
            // enter the per-pixel chain w += dx (:2093-2095) at pixel n0
            float w0 = vx_accum_jump(w0_row, r.dx[0], (uint32_t)n0);
            float w1 = vx_accum_jump(w1_row, r.dx[1], (uint32_t)n0);
            float w2 = vx_accum_jump(w2_row, r.dx[2], (uint32_t)n0);
            const uint32_t type = (r.lo_base >> 4) & 3u;
            unsigned long long *krow = keys + (size_t)y * W + r.min_x;
            for (int n = n0; n < n1; ++n) {
                if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) { // :2050
                    const float bw0 = w0 * r.inv_area, bw1 = w1 * r.inv_area, bw2 = w2 * r.inv_area;
                    const float depth = bw0 * r.z[0] + bw1 * r.z[1] + bw2 * r.z[2]; // :2058
                    if (depth == depth) {
                        const float inv_w_interp = bw0 * r.iw[0] + bw1 * r.iw[1] + bw2 * r.iw[2]; // :2064-2070
                        const float uu = (bw0 * r.uw[0] + bw1 * r.uw[1] + bw2 * r.uw[2]) / inv_w_interp;
                        const float vv = (bw0 * r.vw[0] + bw1 * r.vw[1] + bw2 * r.vw[2]) / inv_w_interp;
                        const uint32_t tex_u = (uint32_t)(vx_f2i(uu * 8.0f) & 7), tex_v = (uint32_t)(vx_f2i(vv * 8.0f) & 7); // :2073-2074
                        const uint32_t pixel_idx = (tex_v << 3) | tex_u; // texture.rs:19-38
                        const uint32_t byte = __ldg(tex_idx + type * 32 + (pixel_idx >> 1));
                        const uint32_t nib = (pixel_idx & 1) ? (byte & 0xFu) : ((byte >> 4) & 0xFu);
                        const unsigned long long key = ((unsigned long long)vx_ord(depth + 0.0f) << 32) | (unsigned long long)(r.lo_base | nib);
                        atomicMin(krow + n, key);
                    }
                }
                w0 = vx_add_rn(w0, r.dx[0]);
                w1 = vx_add_rn(w1, r.dx[1]);
                w2 = vx_add_rn(w2, r.dx[2]);
            }
        }
    }
}

__global__ void __launch_bounds__(BY_THREADS) bary_resolve_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ lut, size_t npx,
                                                                uint32_t *__restrict__ color, float *__restrict__ depth) {
    for (size_t p = (size_t)blockIdx.x * BY_THREADS + threadIdx.x; p < npx; p += (size_t)gridDim.x * BY_THREADS) {
        const unsigned long long k = keys[p];
        const uint32_t lo = (uint32_t)k;
        if (lo) { // payloads are >= 16 (block type >= 1); 0 = the pixel kept its previous content
            depth[p] = vx_unord((uint32_t)(k >> 32));
            color[p] = __ldg(lut + (lo & 511u));
        }
    }
}

} // namespace

extern "C" {

int vx_render_mesh_tiny_quads(VxContext *ctx, const VxMeshBatch *batch, int32_t mesh_id, const float vp[16], const VxFrameConfig *cfg,
                              const int32_t rect[4], int32_t use_span_renderer, uint32_t *color_inout, float *depth_inout) {
    if (use_span_renderer) return vx_render_mesh(ctx, batch, mesh_id, vp, cfg, rect, color_inout, depth_inout);
    if (!ctx || !batch || !vp || !cfg || !rect || !color_inout || !depth_inout || mesh_id < 0 || mesh_id >= batch->n_chunks)
        return vx_fail(ctx, VX_ERR_INVALID, "vx_render_mesh_tiny_quads: bad argument");
    const int W = cfg->width, H = cfg->height;
    if (W <= 0 || H <= 0 || W > 16384 || H > 16384) return vx_fail(ctx, VX_ERR_INVALID, "bad framebuffer size");
    if (rect[0] < 0 || rect[1] < 0 || rect[2] <= 0 || rect[3] <= 0 || rect[0] + rect[2] > W || rect[1] + rect[3] > H)
        return vx_fail(ctx, VX_ERR_INVALID, "bad target rect");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t *d_lut = nullptr;
    const uint8_t *d_tex = nullptr;
    int rc = vx_frame_tables(ctx, *cfg, &d_lut, &d_tex);
    if (rc != VX_OK) return rc;
    uint32_t qbase = 0, qcount = 0;
    uint8_t has = 0;
    int32_t pos[3];
    VX_CUDA(ctx, cudaMemcpyAsync(&qbase, batch->quad_base.as<uint32_t>() + mesh_id, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(&qcount, batch->quad_count.as<uint32_t>() + mesh_id, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(&has, batch->has_mesh.as<uint8_t>() + mesh_id, 1, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(pos, batch->positions.as<int32_t>() + 3 * (size_t)mesh_id, 12, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!has || qcount == 0) return VX_OK; // mesh.is_empty() :789
    if (qcount > (1u << 21)) return vx_fail(ctx, VX_ERR_CAPACITY, "more than 2^21 quads in one mesh");

    const size_t npx = (size_t)W * (size_t)H, n_slots = 4 * (size_t)qcount;
    const size_t off_cnt = (sizeof(BaryRec) * n_slots + 255) & ~(size_t)255;
    const size_t off_base = (off_cnt + 4 * n_slots + 255) & ~(size_t)255;
    const size_t off_total = (off_base + 8 * n_slots + 255) & ~(size_t)255;
    VX_CUDA(ctx, ctx->tmp_b.reserve(8 * npx));
    VX_CUDA(ctx, ctx->tmp_c.reserve(off_total + 256));
    VX_CUDA(ctx, ctx->tmp_d.reserve(8 * npx));
    uint32_t *d_color = ctx->tmp_b.as<uint32_t>();
    float *d_depth = reinterpret_cast<float *>(d_color + npx);
    char *base = ctx->tmp_c.as<char>();
    BaryRec *recs = reinterpret_cast<BaryRec *>(base);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(base + off_cnt);
    unsigned long long *tb = reinterpret_cast<unsigned long long *>(base + off_base);
    unsigned long long *total = reinterpret_cast<unsigned long long *>(base + off_total);
    unsigned long long *keys = ctx->tmp_d.as<unsigned long long>();
    VX_CUDA(ctx, cudaMemcpyAsync(d_color, color_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_depth, depth_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));

    BaryParams P;
    memset(&P, 0, sizeof(P));
    memcpy(P.vp.m, vp, sizeof(float) * 16);
    P.W = W; P.H = H;
    P.rx0 = rect[0]; P.ry0 = rect[1]; P.rw = rect[2]; P.rh = rect[3];
    P.backface = cfg->backface_culling ? 1 : 0;
    P.qbase = qbase; P.qcount = qcount;
    for (int k = 0; k < 3; ++k) P.off[k] = (float)(pos[k] * VX_CHUNK_SIZE); // mesh.rs:483-485
    P.quads = batch->quads.as<uint8_t>();
    P.slice_offsets = batch->slice_offsets.as<uint32_t>() + (size_t)mesh_id * 6 * 33;

    const size_t px_blocks = (npx + BY_THREADS - 1) / BY_THREADS, px_cap = (size_t)ctx->num_sms * 8;
    const int px_grid = (int)(px_blocks < px_cap ? px_blocks : px_cap);
    bary_init_keys_kernel<<<px_grid, BY_THREADS, 0, ctx->stream>>>(d_depth, npx, keys);
    VX_CHECK_LAUNCH(ctx);
    bary_setup_kernel<<<(qcount + BY_THREADS - 1) / BY_THREADS, BY_THREADS, 0, ctx->stream>>>(P, recs, cnt);
    VX_CHECK_LAUNCH(ctx);
    vx_scan_counts_kernel<<<1, VX_SCAN_THREADS, 0, ctx->stream>>>(cnt, (int)n_slots, tb, total);
    VX_CHECK_LAUNCH(ctx);
    bary_fill_kernel<<<ctx->num_sms * 8, BY_THREADS, 0, ctx->stream>>>(recs, tb, (int)n_slots, total, W, d_tex, keys);
    VX_CHECK_LAUNCH(ctx);
    bary_resolve_kernel<<<px_grid, BY_THREADS, 0, ctx->stream>>>(keys, d_lut, npx, d_color, d_depth);
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(color_inout, d_color, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(depth_inout, d_depth, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_render_mesh_with_up(VxContext *ctx, const VxMeshBatch *batch, int32_t mesh_id, const float vp[16], const VxFrameConfig *cfg,
                           const float camera_up[3], uint32_t *color_inout, float *depth_inout) {
    if (!cfg || !camera_up) return vx_fail(ctx, VX_ERR_INVALID, "vx_render_mesh_with_up: bad argument");
    const int32_t rect[4] = {0, 0, cfg->width, cfg->height}; // split_into_stripes(1): the whole framebuffer
    const int level = fabsf(camera_up[1]) >= 0.995f;         // is_camera_level :377-382
    return vx_render_mesh_tiny_quads(ctx, batch, mesh_id, vp, cfg, rect, level, color_inout, depth_inout);
}

} // extern "C"

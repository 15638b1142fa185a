// vx_spanwalk.cu -- the reference's flat-colour span walker on the device:
//   SpanWalkerRasterizer::rasterize_projected_packet  span_walker.rs:116-283
//   FrameSlice::fill_span                             span_walker.rs:412-441
//
// The reference walks quads one after the other, rows top to bottom, and depth-tests every pixel with `<`, so a
// pixel ends up with the colour of the FIRST quad (submission order) among those of minimal depth that cover it.
// Batching quads in groups of eight (TrapezoidBatch) only changes the order in which rows of different quads are
// visited, never the order of two fragments of one pixel.  Here every fragment is a 64-bit key
//   [ order-preserving depth : 32 | 1 + submission index : 32 ]
// merged with atomicMin; the pixel's previous content enters as [depth | 0], which wins every tie exactly like
// `depth < stored` loses it.  Five launches: key init, setup (screen boxes + work count per quad), scan, fill, resolve.
// vx_hyper_pipeline_render swaps the setup for the Hyper-Pipeline front end (face packets -> PacketPipeline, one thread per quad).
#include "vx_common.cuh"
#include "vx_math.cuh"
#include "vx_scan.cuh"

namespace {

constexpr int SW_THREADS = 256;
constexpr int SW_ROWS = 8;         // rows of one fill task

struct SpanRec {
    int32_t y0, y1;  // rows y0 .. y1 inclusive (y0 > y1: nothing)
    int32_t xs, xe;  // columns [xs, xe)
    float depth;
    uint32_t color;
};

struct SpanCtl {
    unsigned long long total_tasks;
};

__device__ __forceinline__ uint32_t block_color(uint8_t t) { // span_walker.rs:386-396, block_type.rs:70-78
    return t == 1 ? 0x00FF00FFu : t == 2 ? 0x8B4513FFu : t == 3 ? 0x808080FFu : 0u;
}

// rasterize_projected_packet :144-175 + the row / column rules of rasterize_batch_scalar :237-260 and fill_span
// :422-428 for one visible quad: NDC box -> rows y0..y1 and columns [xs, xe) (y0 > y1: nothing to draw)
__device__ __forceinline__ void span_from_box(float x_min, float y_min, float x_max, float y_max, float depth, int W, int H, SpanRec &r) {
    const float vp_w = (float)W, vp_h = (float)H;
    const float EPSILON = 0.001f; // span_walker.rs:142
    const float sx_min = fmaxf((x_min + 1.0f) * 0.5f * vp_w, 0.0f);               // :151
    const float sy_min = fmaxf((1.0f - y_max) * 0.5f * vp_h, 0.0f);               // :152
    const float sx_max = fminf((x_max + 1.0f) * 0.5f * vp_w + EPSILON, vp_w);     // :155
    const float sy_max = fminf((1.0f - y_min) * 0.5f * vp_h + EPSILON, vp_h);     // :156
    const bool outside = sx_min >= vp_w || sy_min >= vp_h || sx_max <= 0.0f || sy_max <= 0.0f; // :159-165
    if (outside || !(depth == depth)) return; // a NaN depth never passes `depth < stored`
    // rows whose centre lies in [sy_min, sy_max) (:237-247, update_active_mask :76-84), inside the frame
    int ya = vx_f2i(ceilf(sy_min - 0.5f));
    if (ya < 0) ya = 0;
    while (ya > 0 && (float)(ya - 1) + 0.5f >= sy_min) --ya;
    while (ya < H && (float)ya + 0.5f < sy_min) ++ya;
    int yb = vx_f2i(ceilf(sy_max - 0.5f)) - 1;
    if (yb > H - 1) yb = H - 1;
    while (yb + 1 <= H - 1 && (float)(yb + 1) + 0.5f < sy_max) ++yb;
    while (yb >= 0 && !((float)yb + 0.5f < sy_max)) --yb;
    int xs = vx_f2i(roundf(sx_min)), xe = vx_f2i(roundf(sx_max)); // :253-254
    xs = min(max(xs, 0), W - 1);                                    // fill_span :422-424
    xe = min(max(xe, 0), W);
    if (ya <= yb && xs < xe) {
        r.y0 = ya; r.y1 = yb; r.xs = xs; r.xe = xe;
    }
}

// mode 0: projected quads (NDC boxes).  in_f = x_min, y_min, x_max, y_max, depth_near (n each); in_b = block type, visible
// mode 1: explicit spans.               in_i = y, x_start, x_end (n each); in_f = depth; in_u = colour
__global__ void __launch_bounds__(SW_THREADS) spanwalk_setup_kernel(int mode, const float *__restrict__ in_f, const uint8_t *__restrict__ in_b,
                                                                  const int32_t *__restrict__ in_i, const uint32_t *__restrict__ in_u,
                                                                  int n, int W, int H, SpanRec *__restrict__ recs, uint32_t *__restrict__ n_tasks) {
    const int i = blockIdx.x * SW_THREADS + threadIdx.x;
    if (i >= n) return;
    SpanRec r;
    r.y0 = 1; r.y1 = 0; r.xs = 0; r.xe = 0; r.depth = 0.0f; r.color = 0u;
    const size_t N = (size_t)n;
    if (mode == 0) {
        const bool visible = in_b[N + i] != 0;
        const float depth = in_f[4 * N + i];
        if (visible) span_from_box(in_f[i], in_f[N + i], in_f[2 * N + i], in_f[3 * N + i], depth, W, H, r);
        r.depth = depth;
        r.color = block_color(in_b[i]);
    } else {
        const int y = in_i[i];
        int xs = in_i[N + i], xe = in_i[2 * N + i];
        xs = min(max(xs, 0), W - 1);
        xe = min(max(xe, 0), W);
        r.depth = in_f[i];
        r.color = in_u[i];
        if (xs < xe && r.depth == r.depth) { // y is validated on the host
            r.y0 = y; r.y1 = y; r.xs = xs; r.xe = xe;
        }
    }
    recs[i] = r;
    n_tasks[i] = r.y0 <= r.y1 ? (uint32_t)((r.y1 - r.y0) / SW_ROWS + 1) : 0u;
}

// ---- Hyper-Pipeline front end: ChunkFacePackets::from_chunk_mesh (face_packets.rs:122-174) ->
//      PacketPipeline::process_chunk_packets (packet_pipeline.rs:69-142) for a list of meshes, one thread per quad.
//      Quads reach the span walker in mesh-list order, then face, slice and list order -- the order of the batch's
//      quad stream -- so the global quad index is the submission index.
struct HyperParams {
    VxMat4 vp;
    int32_t W, H, n_meshes;
    const int32_t *ids;                 // [n_meshes] chunk index in the batch
    const unsigned long long *qstart;   // [n_meshes] first global quad of mesh i (exclusive scan of the quad counts)
    const unsigned long long *total;    // number of quads
    const uint8_t *quads;
    const uint32_t *quad_base, *slice_offsets;
    const int32_t *positions;
    uint32_t *n_visible;
};

__global__ void __launch_bounds__(SW_THREADS) hyper_counts_kernel(const int32_t *__restrict__ ids, int n, const uint8_t *__restrict__ has_mesh,
                                                                const uint32_t *__restrict__ quad_count, uint32_t *__restrict__ cnt) {
    const int i = blockIdx.x * SW_THREADS + threadIdx.x;
    if (i < n) cnt[i] = has_mesh[ids[i]] ? quad_count[ids[i]] : 0u;
}

__global__ void __launch_bounds__(SW_THREADS) hyper_setup_kernel(HyperParams P, SpanRec *__restrict__ recs, uint32_t *__restrict__ n_tasks) {
    const unsigned long long total = *P.total;
    for (unsigned long long g = (unsigned long long)blockIdx.x * SW_THREADS + threadIdx.x; g < total; g += (unsigned long long)gridDim.x * SW_THREADS) {
        const int mi = vx_task_owner(P.qstart, P.n_meshes, g);
        const uint32_t q = (uint32_t)(g - P.qstart[mi]);
        const int chunk = P.ids[mi];
        const uint32_t *so = P.slice_offsets + (size_t)chunk * 198;
        int face = 0;
#pragma unroll
        for (int ff = 1; ff < 6; ++ff) face += (so[ff * 33] <= q) ? 1 : 0;
        // the packet of this quad = 32 consecutive quads of the face list (FacePacketBuilder::push :83-100); the whole
        // packet is projected with the basis of ITS FIRST quad's slice (packet_pipeline.rs:93-95), whatever slice the
        // other quads of the packet lie in -- the reference's behaviour, kept as is
        const uint32_t first = so[face * 33] + ((q - so[face * 33]) & ~31u);
        int lo = 0, hi = 31;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (so[face * 33 + mid] <= first) lo = mid; else hi = mid - 1;
        }
        const float s = (float)((face & 1) ? lo : lo + 1); // axis_pos face_packets.rs:148-152
        // FaceBasis::from_face_direction differential_projection.rs:37-62 over face_coordinate_system :231-290
        const int axis = face >> 1;
        const float cw[3] = {(float)P.positions[3 * chunk] * (float)VX_CHUNK_SIZE, (float)P.positions[3 * chunk + 1] * (float)VX_CHUNK_SIZE,
                             (float)P.positions[3 * chunk + 2] * (float)VX_CHUNK_SIZE};
        const float o[3] = {cw[0] + (axis == 0 ? s : 0.0f), cw[1] + (axis == 1 ? s : 0.0f), cw[2] + (axis == 2 ? s : 0.0f)};
        float t[3] = {0, 0, 0}, b[3] = {0, 0, 0}, nn[3] = {0, 0, 0};
        switch (face) {
        case 0: t[1] = 1; b[2] = 1; nn[0] = 1; break;
        case 1: t[1] = 1; b[2] = -1; nn[0] = -1; break;
        case 2: t[0] = 1; b[2] = 1; nn[1] = 1; break;
        case 3: t[0] = 1; b[2] = -1; nn[1] = -1; break;
        case 4: t[0] = 1; b[1] = 1; nn[2] = 1; break;
        default: t[0] = -1; b[1] = 1; nn[2] = -1; break;
        }
        const float4 origin = vx_mul_vec4(P.vp, o[0], o[1], o[2], 1.0f);
        const float4 tangent = vx_mul_vec4(P.vp, t[0], t[1], t[2], 0.0f);
        const float4 bitangent = vx_mul_vec4(P.vp, b[0], b[1], b[2], 0.0f);
        const float4 normal = vx_mul_vec4(P.vp, nn[0], nn[1], nn[2], 0.0f);
        const uint8_t *qp = P.quads + 3 * (size_t)(P.quad_base[chunk] + q);
        const uint32_t b0 = qp[0], b1 = qp[1], b2 = qp[2];
        SpanRec r;
        r.y0 = 1; r.y1 = 0; r.xs = 0; r.xe = 0; r.depth = 0.0f;
        r.color = block_color((uint8_t)((b2 >> 6) & 3));
        if (normal.z < 0.0f) { // is_front_facing :78-82, per packet
            const float u0 = (float)(b0 & 0x1F), v0 = (float)(((b0 >> 5) & 7) | ((b1 & 3) << 3));
            const float u1 = u0 + (float)(((b1 >> 2) & 0x3F) + 1), v1 = v0 + (float)((b2 & 0x3F) + 1);
            float nd[4][3];
#pragma unroll
            for (int c = 0; c < 4; ++c) { // project_single_scalar :167-196 (corners 00, 10, 01, 11)
                const float u = (c & 1) ? u1 : u0, v = (c & 2) ? v1 : v0;
                const float px = (origin.x + u * tangent.x) + v * bitangent.x; // project_point :69-71
                const float py = (origin.y + u * tangent.y) + v * bitangent.y;
                const float pz = (origin.z + u * tangent.z) + v * bitangent.z;
                const float pw = (origin.w + u * tangent.w) + v * bitangent.w;
                nd[c][0] = px / pw; nd[c][1] = py / pw; nd[c][2] = pz / pw; // perspective_divide :412-414
            }
            const float x_min = fminf(fminf(fminf(nd[0][0], nd[1][0]), nd[2][0]), nd[3][0]);
            const float y_min = fminf(fminf(fminf(nd[0][1], nd[1][1]), nd[2][1]), nd[3][1]);
            const float x_max = fmaxf(fmaxf(fmaxf(nd[0][0], nd[1][0]), nd[2][0]), nd[3][0]);
            const float y_max = fmaxf(fmaxf(fmaxf(nd[0][1], nd[1][1]), nd[2][1]), nd[3][1]);
            const float z = fminf(fminf(fminf(nd[0][2], nd[1][2]), nd[2][2]), nd[3][2]);
            r.depth = z;
            // frustum mask, test_aabb_inside packet_pipeline.rs:279-293 with screen box [-1, 1]^2 x [0, 1]
            const bool inside = x_max >= -1.0f && x_min <= 1.0f && y_max >= -1.0f && y_min <= 1.0f && z >= 0.0f && z <= 1.0f;
            if (inside) {
                atomicAdd(P.n_visible, 1u);
                span_from_box(x_min, y_min, x_max, y_max, z, P.W, P.H, r);
            }
        }
        recs[g] = r;
        n_tasks[g] = r.y0 <= r.y1 ? (uint32_t)((r.y1 - r.y0) / SW_ROWS + 1) : 0u;
    }
}

__global__ void __launch_bounds__(SW_THREADS) spanwalk_init_keys_kernel(const float *__restrict__ depth, size_t npx, unsigned long long *__restrict__ keys) {
    for (size_t p = (size_t)blockIdx.x * SW_THREADS + threadIdx.x; p < npx; p += (size_t)gridDim.x * SW_THREADS) {
        const float d = depth[p];
        // a stored NaN rejects every fragment (`x < NaN` is false): key 0 is below every fragment key
        keys[p] = d == d ? ((unsigned long long)vx_ord(d + 0.0f) << 32) : 0ull;
    }
}

// one warp per task = up to SW_ROWS rows of one quad; tasks are taken round-robin by a persistent grid
__global__ void __launch_bounds__(SW_THREADS) spanwalk_fill_kernel(const SpanRec *__restrict__ recs, const unsigned long long *__restrict__ task_base,
                                                                 int n, const SpanCtl *__restrict__ ctl, int W, unsigned long long *__restrict__ keys) {
    const unsigned long long total = ctl->total_tasks;
    const int lane = threadIdx.x & 31;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * (SW_THREADS / 32);
    for (unsigned long long t = (unsigned long long)blockIdx.x * (SW_THREADS / 32) + (threadIdx.x >> 5); t < total; t += n_warps) {
        const int lo = vx_task_owner(task_base, n, t);
        const SpanRec r = recs[lo];
        const int chunk = (int)(t - task_base[lo]);
        const int ya = r.y0 + chunk * SW_ROWS, yb = min(ya + SW_ROWS - 1, r.y1);
        const int w = r.xe - r.xs;
        const unsigned long long key = ((unsigned long long)vx_ord(r.depth + 0.0f) << 32) | (unsigned long long)((uint32_t)lo + 1u);
        const int count = (yb - ya + 1) * w;
        for (int k = lane; k < count; k += 32) {
            const int y = ya + k / w, x = r.xs + k % w;
            atomicMin(&keys[(size_t)y * W + x], key);
        }
    }
}

__global__ void __launch_bounds__(SW_THREADS) spanwalk_resolve_kernel(const unsigned long long *__restrict__ keys, const SpanRec *__restrict__ recs, size_t npx,
                                                                    uint32_t *__restrict__ color, float *__restrict__ depth) {
    for (size_t p = (size_t)blockIdx.x * SW_THREADS + threadIdx.x; p < npx; p += (size_t)gridDim.x * SW_THREADS) {
        const uint32_t seq = (uint32_t)keys[p];
        if (seq) {
            const SpanRec r = recs[seq - 1];
            depth[p] = r.depth; // the quad's own bits (keeps a -0.0)
            color[p] = r.color;
        }
    }
}

struct SpanScratch {
    SpanRec *recs;
    uint32_t *cnt;
    unsigned long long *tb;
    SpanCtl *ctl;
    unsigned long long *keys;
    int px_grid;
};

// scratch: records | task counts | task bases | ctl in tmp_c, keys in tmp_d
int span_scratch(VxContext *ctx, size_t N, size_t npx, SpanScratch &sc) {
    const size_t off_cnt = (sizeof(SpanRec) * N + 255) & ~(size_t)255;
    const size_t off_base = (off_cnt + 4 * N + 255) & ~(size_t)255;
    const size_t off_ctl = (off_base + 8 * N + 255) & ~(size_t)255;
    VX_CUDA(ctx, ctx->tmp_c.reserve(off_ctl + 256));
    VX_CUDA(ctx, ctx->tmp_d.reserve(8 * npx));
    char *base = ctx->tmp_c.as<char>();
    sc.recs = reinterpret_cast<SpanRec *>(base);
    sc.cnt = reinterpret_cast<uint32_t *>(base + off_cnt);
    sc.tb = reinterpret_cast<unsigned long long *>(base + off_base);
    sc.ctl = reinterpret_cast<SpanCtl *>(base + off_ctl);
    sc.keys = ctx->tmp_d.as<unsigned long long>();
    const size_t px_blocks = (npx + SW_THREADS - 1) / SW_THREADS, px_cap = (size_t)ctx->num_sms * 8;
    sc.px_grid = (int)(px_blocks < px_cap ? px_blocks : px_cap);
    return VX_OK;
}

// scan of the task counts, fill, resolve (the records are in place)
int span_walk_finish(VxContext *ctx, const SpanScratch &sc, int32_t n, int32_t W, size_t npx, uint32_t *d_color, float *d_depth) {
    vx_scan_counts_kernel<<<1, VX_SCAN_THREADS, 0, ctx->stream>>>(sc.cnt, n, sc.tb, &sc.ctl->total_tasks);
    VX_CHECK_LAUNCH(ctx);
    spanwalk_fill_kernel<<<ctx->num_sms * 8, SW_THREADS, 0, ctx->stream>>>(sc.recs, sc.tb, n, sc.ctl, W, sc.keys);
    VX_CHECK_LAUNCH(ctx);
    spanwalk_resolve_kernel<<<sc.px_grid, SW_THREADS, 0, ctx->stream>>>(sc.keys, sc.recs, npx, d_color, d_depth);
    VX_CHECK_LAUNCH(ctx);
    return VX_OK;
}

int span_walk_device(VxContext *ctx, int mode, const float *d_f, const uint8_t *d_b, const int32_t *d_i, const uint32_t *d_u, int32_t n,
                     int32_t W, int32_t H, uint32_t *d_color, float *d_depth) {
    const size_t npx = (size_t)W * (size_t)H;
    SpanScratch sc;
    int rc = span_scratch(ctx, (size_t)n, npx, sc);
    if (rc != VX_OK) return rc;
    spanwalk_init_keys_kernel<<<sc.px_grid, SW_THREADS, 0, ctx->stream>>>(d_depth, npx, sc.keys);
    VX_CHECK_LAUNCH(ctx);
    spanwalk_setup_kernel<<<(n + SW_THREADS - 1) / SW_THREADS, SW_THREADS, 0, ctx->stream>>>(mode, d_f, d_b, d_i, d_u, n, W, H, sc.recs, sc.cnt);
    VX_CHECK_LAUNCH(ctx);
    return span_walk_finish(ctx, sc, n, W, npx, d_color, d_depth);
}

} // namespace

extern "C" {

int vx_span_walk_quads_device(VxContext *ctx, const float *d_boxes, const uint8_t *d_types, int32_t n, int32_t width, int32_t height,
                              uint32_t *d_color, float *d_depth) {
    if (!ctx || n < 0 || width <= 0 || height <= 0 || !d_color || !d_depth || (n > 0 && (!d_boxes || !d_types)))
        return vx_fail(ctx, VX_ERR_INVALID, "vx_span_walk_quads_device: bad argument");
    if (n == 0) return VX_OK;
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    return span_walk_device(ctx, 0, d_boxes, d_types, nullptr, nullptr, n, width, height, d_color, d_depth);
}

int vx_span_walk_quads(VxContext *ctx, const float *x_min, const float *y_min, const float *x_max, const float *y_max,
                       const float *depth_near, const uint8_t *block_type, const uint8_t *visible, int32_t n, int32_t width,
                       int32_t height, uint32_t *color_inout, float *depth_inout) {
    if (!ctx || n < 0 || width <= 0 || height <= 0 || !color_inout || !depth_inout)
        return vx_fail(ctx, VX_ERR_INVALID, "vx_span_walk_quads: bad argument");
    if (n == 0) return VX_OK;
    if (!x_min || !y_min || !x_max || !y_max || !depth_near || !block_type) return vx_fail(ctx, VX_ERR_INVALID, "vx_span_walk_quads: null array");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)n, npx = (size_t)width * (size_t)height;
    const size_t off_b = (20 * N + 255) & ~(size_t)255;
    VX_CUDA(ctx, ctx->tmp_a.reserve(off_b + 2 * N));
    VX_CUDA(ctx, ctx->tmp_b.reserve(8 * npx));
    float *d_f = ctx->tmp_a.as<float>();
    uint8_t *d_b = ctx->tmp_a.as<uint8_t>() + off_b;
    const float *src[5] = {x_min, y_min, x_max, y_max, depth_near};
    for (int k = 0; k < 5; ++k) VX_CUDA(ctx, cudaMemcpyAsync(d_f + k * N, src[k], 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_b, block_type, N, cudaMemcpyHostToDevice, ctx->stream));
    if (visible) VX_CUDA(ctx, cudaMemcpyAsync(d_b + N, visible, N, cudaMemcpyHostToDevice, ctx->stream));
    else VX_CUDA(ctx, cudaMemsetAsync(d_b + N, 1, N, ctx->stream));
    uint32_t *d_color = ctx->tmp_b.as<uint32_t>();
    float *d_depth = reinterpret_cast<float *>(d_color + npx);
    VX_CUDA(ctx, cudaMemcpyAsync(d_color, color_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_depth, depth_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = span_walk_device(ctx, 0, d_f, d_b, nullptr, nullptr, n, width, height, d_color, d_depth);
    if (rc != VX_OK) return rc;
    VX_CUDA(ctx, cudaMemcpyAsync(color_inout, d_color, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(depth_inout, d_depth, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_fill_spans(VxContext *ctx, const int32_t *y, const int32_t *x_start, const int32_t *x_end, const float *depth,
                  const uint32_t *color, int32_t n, int32_t width, int32_t height, uint32_t *color_inout, float *depth_inout) {
    if (!ctx || n < 0 || width <= 0 || height <= 0 || !color_inout || !depth_inout)
        return vx_fail(ctx, VX_ERR_INVALID, "vx_fill_spans: bad argument");
    if (n == 0) return VX_OK;
    if (!y || !x_start || !x_end || !depth || !color) return vx_fail(ctx, VX_ERR_INVALID, "vx_fill_spans: null array");
    for (int32_t i = 0; i < n; ++i) // the reference indexes row y unchecked and would panic
        if (y[i] < 0 || y[i] >= height) return vx_fail(ctx, VX_ERR_INVALID, "vx_fill_spans: row outside the framebuffer");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)n, npx = (size_t)width * (size_t)height;
    VX_CUDA(ctx, ctx->tmp_a.reserve(20 * N));
    VX_CUDA(ctx, ctx->tmp_b.reserve(8 * npx));
    int32_t *d_i = ctx->tmp_a.as<int32_t>();
    float *d_f = reinterpret_cast<float *>(d_i + 3 * N);
    uint32_t *d_u = reinterpret_cast<uint32_t *>(d_i + 4 * N);
    VX_CUDA(ctx, cudaMemcpyAsync(d_i, y, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_i + N, x_start, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_i + 2 * N, x_end, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_f, depth, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_u, color, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *d_color = ctx->tmp_b.as<uint32_t>();
    float *d_depth = reinterpret_cast<float *>(d_color + npx);
    VX_CUDA(ctx, cudaMemcpyAsync(d_color, color_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_depth, depth_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = span_walk_device(ctx, 1, d_f, nullptr, d_i, d_u, n, width, height, d_color, d_depth);
    if (rc != VX_OK) return rc;
    VX_CUDA(ctx, cudaMemcpyAsync(color_inout, d_color, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(depth_inout, d_depth, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_hyper_pipeline_render(VxContext *ctx, const VxMeshBatch *batch, const int32_t *mesh_ids, int32_t n_meshes, const float vp[16],
                             int32_t width, int32_t height, uint32_t *color_inout, float *depth_inout, int32_t *n_visible_quads) {
    if (!ctx || !batch || !vp || n_meshes < 0 || width <= 0 || height <= 0 || !color_inout || !depth_inout || (n_meshes > 0 && !mesh_ids))
        return vx_fail(ctx, VX_ERR_INVALID, "vx_hyper_pipeline_render: bad argument");
    if (n_visible_quads) *n_visible_quads = 0;
    if (n_meshes == 0) return VX_OK;
    for (int32_t i = 0; i < n_meshes; ++i)
        if (mesh_ids[i] < 0 || mesh_ids[i] >= batch->n_chunks) return vx_fail(ctx, VX_ERR_INVALID, "mesh id out of range");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t M = (size_t)n_meshes, npx = (size_t)width * (size_t)height;
    // tmp_a: ids | quad counts | first global quad per mesh | total, visible counter
    const size_t off_cnt = (4 * M + 255) & ~(size_t)255;
    const size_t off_qs = (off_cnt + 4 * M + 255) & ~(size_t)255;
    const size_t off_tot = (off_qs + 8 * M + 255) & ~(size_t)255;
    VX_CUDA(ctx, ctx->tmp_a.reserve(off_tot + 256));
    VX_CUDA(ctx, ctx->tmp_b.reserve(8 * npx));
    char *a = ctx->tmp_a.as<char>();
    int32_t *d_ids = reinterpret_cast<int32_t *>(a);
    uint32_t *d_cnt = reinterpret_cast<uint32_t *>(a + off_cnt);
    unsigned long long *d_qs = reinterpret_cast<unsigned long long *>(a + off_qs);
    unsigned long long *d_total = reinterpret_cast<unsigned long long *>(a + off_tot);
    uint32_t *d_visible = reinterpret_cast<uint32_t *>(a + off_tot + 16);
    VX_CUDA(ctx, cudaMemcpyAsync(d_ids, mesh_ids, 4 * M, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemsetAsync(a + off_tot, 0, 32, ctx->stream));
    hyper_counts_kernel<<<(n_meshes + SW_THREADS - 1) / SW_THREADS, SW_THREADS, 0, ctx->stream>>>(d_ids, n_meshes, batch->has_mesh.as<uint8_t>(),
                                                                                             batch->quad_count.as<uint32_t>(), d_cnt);
    VX_CHECK_LAUNCH(ctx);
    vx_scan_counts_kernel<<<1, VX_SCAN_THREADS, 0, ctx->stream>>>(d_cnt, n_meshes, d_qs, d_total);
    VX_CHECK_LAUNCH(ctx);
    unsigned long long total = 0;
    VX_CUDA(ctx, cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (total == 0) return VX_OK;
    if (total >= (1ull << 31)) return vx_fail(ctx, VX_ERR_CAPACITY, "more than 2^31 quads in the mesh list");
    const int32_t n = (int32_t)total;
    SpanScratch sc;
    int rc = span_scratch(ctx, (size_t)n, npx, sc);
    if (rc != VX_OK) return rc;
    uint32_t *d_color = ctx->tmp_b.as<uint32_t>();
    float *d_depth = reinterpret_cast<float *>(d_color + npx);
    VX_CUDA(ctx, cudaMemcpyAsync(d_color, color_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_depth, depth_inout, 4 * npx, cudaMemcpyHostToDevice, ctx->stream));
    spanwalk_init_keys_kernel<<<sc.px_grid, SW_THREADS, 0, ctx->stream>>>(d_depth, npx, sc.keys);
    VX_CHECK_LAUNCH(ctx);
    HyperParams P;
    memset(&P, 0, sizeof(P));
    memcpy(P.vp.m, vp, sizeof(float) * 16);
    P.W = width; P.H = height; P.n_meshes = n_meshes;
    P.ids = d_ids; P.qstart = d_qs; P.total = d_total;
    P.quads = batch->quads.as<uint8_t>();
    P.quad_base = batch->quad_base.as<uint32_t>();
    P.slice_offsets = batch->slice_offsets.as<uint32_t>();
    P.positions = batch->positions.as<int32_t>();
    P.n_visible = d_visible;
    const size_t q_blocks = ((size_t)n + SW_THREADS - 1) / SW_THREADS, q_cap = (size_t)ctx->num_sms * 16;
    hyper_setup_kernel<<<(int)(q_blocks < q_cap ? q_blocks : q_cap), SW_THREADS, 0, ctx->stream>>>(P, sc.recs, sc.cnt);
    VX_CHECK_LAUNCH(ctx);
    rc = span_walk_finish(ctx, sc, n, width, npx, d_color, d_depth);
    if (rc != VX_OK) return rc;
    uint32_t visible = 0;
    VX_CUDA(ctx, cudaMemcpyAsync(&visible, d_visible, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(color_inout, d_color, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(depth_inout, d_depth, 4 * npx, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_visible_quads) *n_visible_quads = (int32_t)visible;
    return VX_OK;
}

} // extern "C"

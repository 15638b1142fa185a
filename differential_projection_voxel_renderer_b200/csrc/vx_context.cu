// vx_context.cu -- context lifetime, defaults, error strings.
#include "vx_common.cuh"

extern "C" {

int vx_context_create(int device_id, VxContext **out) {
    if (!out) return VX_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return VX_ERR_NO_DEVICE; // no CPU fallback by design
    if (device_id < 0 || device_id >= n) return VX_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return VX_ERR_NO_DEVICE;
    if (prop.major < 10) return VX_ERR_NO_DEVICE; // sm_100a cubin only
    if (cudaSetDevice(device_id) != cudaSuccess) return VX_ERR_NO_DEVICE;
    VxContext *ctx = new VxContext();
    ctx->device = device_id;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return VX_ERR_CUDA;
    }
    vx_default_atlas(&ctx->atlas);
    ctx->atlas_dirty = true;
    *out = ctx;
    return VX_OK;
}

void vx_context_destroy(VxContext *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    vx_frame_scratch_destroy(ctx);
    ctx->tmp_a.release();
    ctx->tmp_b.release();
    ctx->tmp_c.release();
    ctx->tmp_d.release();
    ctx->pinned.release();
    ctx->multi_ptrs.release();
    ctx->multi_status.release();
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *vx_error_string(int code) {
    switch (code) {
    case VX_OK: return "ok";
    case VX_ERR_INVALID: return "invalid argument";
    case VX_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
    case VX_ERR_CUDA: return "CUDA runtime error";
    case VX_ERR_CAPACITY: return "capacity exceeded";
    case VX_ERR_OOM: return "out of device memory";
    default: return "unknown error";
    }
}

const char *vx_last_error(const VxContext *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int vx_device_synchronize(VxContext *ctx) {
    if (!ctx) return VX_ERR_INVALID;
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

void *vx_context_stream(VxContext *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int64_t vx_context_launch_count(const VxContext *ctx) { return ctx ? ctx->launches : 0; }

// ---- defaults -------------------------------------------------------------------------------

void vx_default_frame_config(VxFrameConfig *cfg, int32_t width, int32_t height) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->width = width;
    cfg->height = height;
    cfg->clear_color = 0xFF87CEEBu; // main.rs:393
    cfg->backface_culling = 1;      // rasterizer.rs:366
    cfg->enable_shading = 1;        // rasterizer.rs:368
    // Vec3(0.4, 1.0, 0.3).normalize() as spelled out at rasterizer.rs:1206-1208
    cfg->light_dir[0] = 0.35634832f;
    cfg->light_dir[1] = 0.8908708f;
    cfg->light_dir[2] = 0.2672612f;
    cfg->ambient = 0.35f;
    cfg->diffuse = 0.65f;
    cfg->differential_projection = 0;
    cfg->occlusion_culling = 0;     // main.rs:112
    cfg->occlusion_grid_w = 128;    // main.rs:46-47
    cfg->occlusion_grid_h = 72;
}

static uint32_t rgb565_to_argb32(uint16_t c) { // texture.rs:42-54
    const uint32_t r = (c >> 11) & 0x1F, g = (c >> 5) & 0x3F, b = c & 0x1F;
    return 0xFF000000u | (((r << 3) | (r >> 2)) << 16) | (((g << 2) | (g >> 4)) << 8) | ((b << 3) | (b >> 2));
}

void vx_default_atlas(VxAtlas *a) {
    if (!a) return;
    memset(a, 0, sizeof(*a));
    // 0: magenta/black checkerboard (texture.rs:81-101)
    a->palette[0][0] = rgb565_to_argb32(0xF81F);
    a->palette[0][1] = rgb565_to_argb32(0x0000);
    for (int i = 0; i < 64; ++i) {
        const uint8_t ci = (uint8_t)(((i % 8) + (i / 8)) % 2);
        a->indices[0][i / 2] |= (i % 2 == 0) ? (uint8_t)(ci << 4) : ci;
    }
    // 1..3: two-tone noise, shared LCG index pattern (texture.rs:103-123)
    const uint16_t base[4] = {0, 0x03E0, 0x8A22, 0x8410}, dark[4] = {0, 0x02E0, 0x71C2, 0x73AE};
    for (int t = 1; t < 4; ++t) {
        for (int i = 0; i < 16; ++i) a->palette[t][i] = rgb565_to_argb32((i % 2 == 0) ? base[t] : dark[t]);
        uint32_t seed = 12345;
        for (int i = 0; i < 32; ++i) {
            seed = seed * 1103515245u + 12345u;
            a->indices[t][i] = (uint8_t)(seed >> 16);
        }
    }
}

int vx_set_atlas(VxContext *ctx, const VxAtlas *atlas) {
    if (!ctx || !atlas) return vx_fail(ctx, VX_ERR_INVALID, "vx_set_atlas: bad argument");
    ctx->atlas = *atlas;
    ctx->atlas_dirty = true;
    return VX_OK;
}

} // extern "C"

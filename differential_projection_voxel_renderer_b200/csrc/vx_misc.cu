// vx_misc.cu -- stand-alone entry points around the frame path: chunk visibility (filter A), FaceBasis /
// packet projection (the reference's "Hyper-Pipeline" projection stage), legacy vertex transform and a
// vertex dump used by the parity tests.                      (compiled with -fmad=false, see vx_math.cuh)
#include "vx_common.cuh"

#include <cmath>

#include <math_constants.h>
#include "vx_math.cuh"

namespace {

// World::get_visible_chunks_frustum world.rs:118-146
__global__ void cull_chunks_kernel(const int32_t *positions, int32_t n, VxMat4 vp, float cx, float cy, float cz,
                                   int32_t view_distance, int32_t frustum, uint8_t *visible) {
    __shared__ float planes[6][4];
    if (threadIdx.x < 6) vx_frustum_plane(vp, threadIdx.x, planes[threadIdx.x]);
    __syncthreads();
    const int32_t cc[3] = {vx_f2i(floorf(cx / (float)VX_CHUNK_SIZE)), vx_f2i(floorf(cy / (float)VX_CHUNK_SIZE)),
                           vx_f2i(floorf(cz / (float)VX_CHUNK_SIZE))};
    const float vd_sq = (float)(view_distance * view_distance);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t p[3] = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
        visible[i] = vx_chunk_visible(p, cc, vd_sq, frustum != 0, planes) ? 1 : 0;
    }
}

// face_coordinate_system differential_projection.rs:231-290 + FaceBasis::from_face_direction :37-62.
// Note the mirrored (bi)tangents of the negative faces: this is the reference's FaceBasis API verbatim.
__global__ void face_basis_kernel(const int32_t *faces, const int32_t *chunk_pos, const uint8_t *slice_idx, int32_t n,
                                  VxMat4 vp, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int face = faces[i];
    const float cw[3] = {(float)chunk_pos[3 * i] * (float)VX_CHUNK_SIZE, (float)chunk_pos[3 * i + 1] * (float)VX_CHUNK_SIZE,
                         (float)chunk_pos[3 * i + 2] * (float)VX_CHUNK_SIZE};
    const float s = (float)slice_idx[i];
    const int axis = face >> 1;
    float o[3] = {cw[0] + (axis == 0 ? s : 0.0f), cw[1] + (axis == 1 ? s : 0.0f), cw[2] + (axis == 2 ? s : 0.0f)};
    float t[3] = {0, 0, 0}, b[3] = {0, 0, 0}, nn[3] = {0, 0, 0};
    switch (face) {
    case 0: t[1] = 1; b[2] = 1; nn[0] = 1; break;
    case 1: t[1] = 1; b[2] = -1; nn[0] = -1; break;
    case 2: t[0] = 1; b[2] = 1; nn[1] = 1; break;
    case 3: t[0] = 1; b[2] = -1; nn[1] = -1; break;
    case 4: t[0] = 1; b[1] = 1; nn[2] = 1; break;
    default: t[0] = -1; b[1] = 1; nn[2] = -1; break;
    }
    float4 *dst = reinterpret_cast<float4 *>(out + 16 * (size_t)i);
    dst[0] = vx_mul_vec4(vp, o[0], o[1], o[2], 1.0f);
    dst[1] = vx_mul_vec4(vp, t[0], t[1], t[2], 0.0f);
    dst[2] = vx_mul_vec4(vp, b[0], b[1], b[2], 0.0f);
    dst[3] = vx_mul_vec4(vp, nn[0], nn[1], nn[2], 0.0f);
}

struct Basis {
    float v[16];
};

// project_point :69-71 (origin + u*tangent + v*bitangent, left to right) + project_single_scalar :167-196
__global__ void project_packet_kernel(Basis bs, const uint8_t *u_min, const uint8_t *v_min, const uint8_t *u_len,
                                      const uint8_t *v_len, int32_t n, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float u0 = (float)u_min[i], v0 = (float)v_min[i];
    const float u1 = u0 + (float)u_len[i], v1 = v0 + (float)v_len[i];
    float nd[4][3];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float u = (c & 1) ? u1 : u0, v = (c & 2) ? v1 : v0;
        float p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = (bs.v[k] + u * bs.v[4 + k]) + v * bs.v[8 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) nd[c][k] = p[k] / p[3]; // perspective_divide :412-414
    }
    out[0 * (size_t)n + i] = fminf(fminf(fminf(nd[0][0], nd[1][0]), nd[2][0]), nd[3][0]);
    out[1 * (size_t)n + i] = fminf(fminf(fminf(nd[0][1], nd[1][1]), nd[2][1]), nd[3][1]);
    out[2 * (size_t)n + i] = fmaxf(fmaxf(fmaxf(nd[0][0], nd[1][0]), nd[2][0]), nd[3][0]);
    out[3 * (size_t)n + i] = fmaxf(fmaxf(fmaxf(nd[0][1], nd[1][1]), nd[2][1]), nd[3][1]);
    out[4 * (size_t)n + i] = fminf(fminf(fminf(nd[0][2], nd[1][2]), nd[2][2]), nd[3][2]);
}

// decompress_and_transform_vertices_scalar simd_vertex.rs:48-58
__global__ void transform_vertices_kernel(const VxVertex *verts, int32_t n, float ox, float oy, float oz, VxMat4 vp,
                                          float4 *out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 raw = *reinterpret_cast<const uint2 *>(verts + i); // 8-byte vertex, one 64-bit load
        const float x = (float)(raw.x & 0xFF), y = (float)((raw.x >> 8) & 0xFF), z = (float)((raw.x >> 16) & 0xFF);
        out[i] = vx_mul_point(vp, ox + x, oy + y, oz + z);
    }
}

// the four clip-space corners of every quad of one mesh, as the raster setup computes them
__global__ void project_mesh_vertices_kernel(const uint8_t *quads, const uint32_t *slice_offsets, uint32_t qbase,
                                             uint32_t qcount, float ox, float oy, float oz, VxMat4 vp, int differential,
                                             float4 *out) {
    __shared__ uint32_t so[198];
    __shared__ float4 origin[3][33];
    for (int i = threadIdx.x; i < 198; i += blockDim.x) so[i] = slice_offsets[i];
    for (int i = threadIdx.x; i < 99; i += blockDim.x) {
        const int axis = i / 33, s = i % 33;
        origin[axis][s] = vx_mul_point(vp, ox + (axis == 0 ? (float)s : 0.0f), oy + (axis == 1 ? (float)s : 0.0f),
                                       oz + (axis == 2 ? (float)s : 0.0f));
    }
    __syncthreads();
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < qcount; q += gridDim.x * blockDim.x) {
        int face = 0;
        for (int ff = 1; ff < 6; ++ff) face += (so[ff * 33] <= q) ? 1 : 0;
        int lo = 0, hi = 31;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (so[face * 33 + mid] <= q) lo = mid; else hi = mid - 1;
        }
        const int axis = face >> 1, spos = (face & 1) ? lo : lo + 1;
        const uint8_t *qp = quads + 3 * (size_t)(qbase + q);
        const uint32_t b0 = qp[0], b1 = qp[1], b2 = qp[2];
        const int u = b0 & 0x1F, v = ((b0 >> 5) & 7) | ((b1 & 3) << 3);
        const int w = ((b1 >> 2) & 0x3F) + 1, h = (b2 & 0x3F) + 1;
        for (int i = 0; i < 4; ++i) {
            const int cu = ((kCornerU[face] >> i) & 1) ? u + w : u;
            const int cv = ((kCornerV[face] >> i) & 1) ? v + h : v;
            int lx, ly, lz;
            if (axis == 0) { lx = spos; ly = cu; lz = cv; }
            else if (axis == 1) { lx = cu; ly = spos; lz = cv; }
            else { lx = cu; ly = cv; lz = spos; }
            float4 p;
            if (!differential) p = vx_mul_point(vp, ox + (float)lx, oy + (float)ly, oz + (float)lz);
            else {
                const float4 o = origin[axis][spos];
                const int ta = axis == 0 ? 1 : 0, ba = axis == 2 ? 1 : 2;
                const float fu = (float)cu, fv = (float)cv;
                p.x = fmaf(fu, vp.m[ta * 4 + 0], fmaf(fv, vp.m[ba * 4 + 0], o.x));
                p.y = fmaf(fu, vp.m[ta * 4 + 1], fmaf(fv, vp.m[ba * 4 + 1], o.y));
                p.z = fmaf(fu, vp.m[ta * 4 + 2], fmaf(fv, vp.m[ba * 4 + 2], o.z));
                p.w = fmaf(fu, vp.m[ta * 4 + 3], fmaf(fv, vp.m[ba * 4 + 3], o.w));
            }
            out[4 * (size_t)q + i] = p;
        }
    }
}

VxMat4 to_mat(const float vp[16]) {
    VxMat4 m;
    memcpy(m.m, vp, sizeof(float) * 16);
    return m;
}


// ------------------------------------------------------------------------------------------------
// culling::apply_horizon_culling (culling.rs:40-119).  One CTA.  The reference is a serial front-to-back sweep with a
// running horizon per angular bin; bins are independent, so after the stable distance sort every bin is swept by its
// own thread.  The angle of every candidate is an input: the reference calls the platform libm's atan2f
// (culling.rs:86, Rust std on Linux = the C library's atan2f), whose last bit no device routine is specified to
// reproduce, and one ulp decides the bin of a mesh that sits on a bin boundary (lattice-aligned chunk centres do, e.g.
// on the diagonals).  vx_horizon_cull therefore evaluates atan2f for its n candidates on the host -- the same call
// the reference makes -- and hands the kernel the angles; everything else runs here.
// ------------------------------------------------------------------------------------------------
constexpr int HZ_THREADS = 1024;

struct HorizonArgs {
    const float *centers; // [*][3]
    int32_t *order;       // [n] in: candidate ids, out: kept ids front-to-back
    int32_t n, bins;
    float cam[3];
    float base_margin, margin_dist_factor, min_dist_chunks;
    const float *angle_in; // [n] atan2f(center.z - cam.z, center.x - cam.x) of order[i] (host libm, see above)
    float *angle;         // [n] scratch: the same in sorted order
    float *key;           // [n] scratch: distance_sq in input order
    int32_t *sorted;      // [n] scratch: ids sorted by distance
    int32_t *bin_of;      // [n] scratch: bin of sorted[i], -1 = always kept
    float *slope, *margin, *top; // [n] scratch
    uint8_t *keep;        // [n] scratch
    int32_t *n_kept;
};

__global__ void __launch_bounds__(HZ_THREADS) horizon_cull_kernel(HorizonArgs a) {
    __shared__ uint32_t warp_sums[HZ_THREADS / 32];
    __shared__ uint32_t s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // distance_sq of every candidate (main.rs:286-290 form: |center - cam|^2, left to right)
    for (int i = tid; i < a.n; i += HZ_THREADS) {
        const float *c = a.centers + 3 * (size_t)a.order[i];
        const float dx = c[0] - a.cam[0], dy = c[1] - a.cam[1], dz = c[2] - a.cam[2];
        a.key[i] = dx * dx + dy * dy + dz * dz;
    }
    __syncthreads();
    // stable sort by distance_sq (partial_cmp, ties keep input order): rank = elements ordered before mine
    for (int i = tid; i < a.n; i += HZ_THREADS) {
        const float k = a.key[i];
        int rank = 0;
        for (int j = 0; j < a.n; ++j) {
            const float kj = a.key[j];
            rank += (kj < k || (kj == k && j < i)) ? 1 : 0;
        }
        a.sorted[rank] = a.order[i];
        a.angle[rank] = a.angle_in[i];
    }
    __syncthreads();
    // per-mesh quantities
    const float chunk_size = (float)VX_CHUNK_SIZE, half_chunk = chunk_size * 0.5f;
    const float PI = 3.14159265358979323846f;
    for (int i = tid; i < a.n; i += HZ_THREADS) {
        const float *c = a.centers + 3 * (size_t)a.sorted[i];
        const float tx = c[0] - a.cam[0], tz = c[2] - a.cam[2];
        const float dist_xz = sqrtf(tx * tx + tz * tz);
        int32_t bin = -1;
        float slope = 0.0f, margin = 0.0f, top = 0.0f;
        if (!(dist_xz < 1e-3f)) {
            const float dist_chunks = dist_xz / chunk_size;
            if (!(dist_chunks < a.min_dist_chunks)) {
                const float angle = a.angle[i];
                const float bin_f = (angle + PI) / (2.0f * PI) * (float)a.bins;
                long b = (long)vx_f2i(floorf(bin_f));
                if (b < 0) b += a.bins;
                bin = (int32_t)(b % a.bins);
                slope = (c[1] - a.cam[1]) / dist_xz;
                margin = a.base_margin * (1.0f + dist_chunks * a.margin_dist_factor);
                top = (c[1] + half_chunk - a.cam[1]) / dist_xz;
            }
        }
        a.bin_of[i] = bin;
        a.slope[i] = slope;
        a.margin[i] = margin;
        a.top[i] = top;
        a.keep[i] = bin < 0 ? 1 : 0;
    }
    __syncthreads();
    // one thread per angular bin sweeps the sorted list with its running horizon
    for (int b = tid; b < a.bins; b += HZ_THREADS) {
        float horizon = -CUDART_INF_F;
        for (int i = 0; i < a.n; ++i) {
            if (a.bin_of[i] != b) continue;
            const float s = a.slope[i];
            const bool cull = s >= 0.0f && (s + a.margin[i]) < horizon;
            if (!cull) {
                a.keep[i] = 1;
                if (a.top[i] > horizon) horizon = a.top[i];
            }
        }
    }
    __syncthreads();
    // stable compaction of the kept meshes
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < a.n; base += HZ_THREADS) {
        const int i = base + tid;
        const uint32_t k = (i < a.n && a.keep[i]) ? 1u : 0u;
        uint32_t inc = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < HZ_THREADS / 32; ++w) {
            const uint32_t c = warp_sums[w];
            if (w < warp) before += c;
            total += c;
        }
        if (k) a.order[s_run + before + inc - 1] = a.sorted[i];
        __syncthreads();
        if (tid == 0) s_run += total;
        __syncthreads();
    }
    if (tid == 0) *a.n_kept = (int32_t)s_run;
}


// ------------------------------------------------------------------------------------------------
// Terrain generation (Chunk::generate_terrain, chunk.rs:114-207): heightfield
// h = trunc(noise(x * scale, z * scale) * amplitude) with `noise 0.9.0` perlin_2d in f64 (restated in
// the CPU test restatement (oracle/) from the crate's published source), Grass at y == h, Dirt for h-3 < y < h, Stone below, Air above;
// chunks entirely above the terrain are Uniform(Air), chunks more than 10 below it Uniform(Stone) (chunk.rs:127-134).
// One CTA per chunk.  The arithmetic follows the oracle operation by operation (this translation unit is compiled
// -fmad=false), so the voxels equal the oracle's and the host generator's bit for bit.  The permutation table
// (PermutationTable::new(seed)) and the four gradients are inputs.
// ------------------------------------------------------------------------------------------------
struct TerrainArgs {
    const int32_t *positions; // [n][3]
    int32_t n;
    const int32_t *perm;      // [512]
    double grad[8][2];
    double scale, amplitude;
    uint8_t *voxels;          // [n][32768]
    uint8_t *flags;           // [n]
};

__device__ __forceinline__ double terrain_fade(double t) { // map_quintic: clamp to [0, 1], then t^3 (t (6 t - 15) + 10)
    const double x = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    return x * x * x * (x * (x * 6.0 - 15.0) + 10.0);
}

__global__ void __launch_bounds__(256) generate_terrain_kernel(TerrainArgs a) {
    __shared__ int32_t h[32][32]; // [z][x]
    __shared__ int32_t s_min, s_max;
    const int tid = threadIdx.x;
    for (int chunk = blockIdx.x; chunk < a.n; chunk += gridDim.x) {
        const int cx = a.positions[3 * chunk], cy = a.positions[3 * chunk + 1], cz = a.positions[3 * chunk + 2];
        if (tid == 0) {
            s_min = INT32_MAX;
            s_max = INT32_MIN;
        }
        __syncthreads();
        int mn = INT32_MAX, mx = INT32_MIN;
        for (int c = tid; c < 1024; c += 256) {
            const int lz = c >> 5, lx = c & 31;
            // noise 0.9.0 perlin_2d (core/perlin.rs), restated in the CPU restatement under oracle/
            const double x = (double)(cx * VX_CHUNK_SIZE + lx) * a.scale;
            const double y = (double)(cz * VX_CHUNK_SIZE + lz) * a.scale;
            const double fx = floor(x), fy = floor(y);
            const long long xi0 = (long long)fx, yi0 = (long long)fy;
            const double xf = x - (double)xi0, yf = y - (double)yi0;
            double g[2][2];
#pragma unroll
            for (int ox = 0; ox < 2; ++ox)
#pragma unroll
                for (int oy = 0; oy < 2; ++oy) {
                    const double qx = xf - (double)ox, qy = yf - (double)oy;
                    const int hh = a.perm[a.perm[(int)((xi0 + ox) & 255)] ^ (int)((yi0 + oy) & 255)] & 3; // NoiseHasher::hash
                    g[ox][oy] = a.grad[hh][0] * qx + a.grad[hh][1] * qy; // gradients are (+-1, +-1): exactly +-qx +- qy
                }
            const double u = terrain_fade(xf), v = terrain_fade(yf);
            const double k0 = g[0][0], k1 = g[1][0] - g[0][0], k2 = g[0][1] - g[0][0], k3 = g[0][0] + g[1][1] - g[1][0] - g[0][1];
            double nval = (k0 + k1 * u + k2 * v + k3 * u * v) * (2.0 / 1.4142135623730951);
            nval = nval < -1.0 ? -1.0 : (nval > 1.0 ? 1.0 : nval);
            const int hv = (int)trunc(nval * a.amplitude);
            h[lz][lx] = hv;
            mn = min(mn, hv);
            mx = max(mx, hv);
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((tid & 31) == 0) {
            atomicMin(&s_min, mn);
            atomicMax(&s_max, mx);
        }
        __syncthreads();
        const int y0 = cy * VX_CHUNK_SIZE;
        uint8_t flag = 0;
        if (y0 > s_max) flag = 1;                          // Uniform(Air)   chunk.rs:127-129
        else if (y0 + VX_CHUNK_SIZE < s_min - 10) flag = 4; // Uniform(Stone) chunk.rs:132-134
        if (tid == 0) a.flags[chunk] = flag;
        uint4 *dst = reinterpret_cast<uint4 *>(a.voxels + (size_t)chunk * VX_CHUNK_VOLUME);
        for (int r = tid; r < 1024; r += 256) { // r = z*32 + y: one 32-voxel x-row
            const int lz = r >> 5, wy = y0 + (r & 31);
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t word = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int hh = h[lz][4 * j + k];
                    uint32_t t = 3u;            // Stone
                    if (wy > hh - 3) t = 2u;    // Dirt
                    if (wy == hh) t = 1u;       // Grass
                    if (wy > hh) t = 0u;        // Air
                    word |= t << (8 * k);
                }
                w[j] = flag ? 0u : word;
            }
            dst[2 * r] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[2 * r + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------
// ChunkFacePackets::from_chunk_mesh (face_packets.rs:122-174): the quads of a mesh regrouped per face into SoA
// packets of up to 32 (FacePacket32 :13-25), in list order (slice by slice), axis_pos = slice + 1 for positive faces.
// One thread per quad; unused lanes of the last packet of a face stay zero (FacePacket32::new).
// ------------------------------------------------------------------------------------------------
__global__ void face_packets_kernel(const uint8_t *quads, const uint32_t *slice_offsets, uint32_t qbase, uint32_t qcount,
                                    VxFacePacket32 *out, uint32_t cap_packets) {
    __shared__ uint32_t so[198];
    __shared__ uint32_t pbase[7]; // first packet of each face
    for (int i = threadIdx.x; i < 198; i += blockDim.x) so[i] = slice_offsets[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int f = 0; f < 6; ++f) {
            pbase[f] = run;
            run += (so[f * 33 + 32] - so[f * 33] + 31u) / 32u;
        }
        pbase[6] = run;
    }
    __syncthreads();
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < qcount; q += gridDim.x * blockDim.x) {
        int face = 0;
        for (int ff = 1; ff < 6; ++ff) face += (so[ff * 33] <= q) ? 1 : 0;
        int lo = 0, hi = 31;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (so[face * 33 + mid] <= q) lo = mid; else hi = mid - 1;
        }
        const uint32_t in_face = q - so[face * 33];
        const uint32_t pk = pbase[face] + in_face / 32u, lane = in_face % 32u;
        if (pk >= cap_packets) continue;
        const uint8_t *qp = quads + 3 * (size_t)(qbase + q);
        const uint32_t b0 = qp[0], b1 = qp[1], b2 = qp[2];
        VxFacePacket32 &P = out[pk];
        P.u_min[lane] = (uint8_t)(b0 & 0x1F);
        P.v_min[lane] = (uint8_t)(((b0 >> 5) & 7) | ((b1 & 3) << 3)); // TinyQuad accessors mesh.rs:309-341
        P.u_len[lane] = (uint8_t)(((b1 >> 2) & 0x3F) + 1);
        P.v_len[lane] = (uint8_t)((b2 & 0x3F) + 1);
        P.axis_pos[lane] = (uint8_t)((face & 1) ? lo : lo + 1);
        P.block_type[lane] = (uint8_t)((b2 >> 6) & 3);
        const uint32_t face_quads = so[face * 33 + 32] - so[face * 33];
        if (lane == 0) P.len = (uint8_t)min(32u, face_quads - (in_face / 32u) * 32u);
    }
}

// vx_div_fast vs the `/` operator on pseudo-random operand pairs.  counters: [0] mismatching quotients among pairs the
// guard accepted, [1] pairs the guard sent to the fallback, [2] pairs tested.
__global__ void selftest_division_kernel(unsigned long long seed, unsigned long long n, int mode, unsigned long long *counters) {
    unsigned long long bad = 0, fb = 0, cnt = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull ^ seed; // splitmix64
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x ^= x >> 31;
        uint32_t ua = (uint32_t)x, ub = (uint32_t)(x >> 32);
        if (mode == 1) { // magnitudes of the raster path: exponents within +-40 of 1.0, numerator sometimes 0
            ua = (ua & 0x807FFFFFu) | ((87u + (ua >> 23) % 80u) << 23);
            ub = (ub & 0x807FFFFFu) | ((87u + (ub >> 23) % 80u) << 23);
            if ((x & 0xFF000000000000ull) == 0) ua &= 0x80000000u;
        } else if (mode == 2) { // few mantissa bits: exact and tie-prone quotients
            ua = (ua & 0xFFF80000u);
            ub = (ub & 0xFFF80000u);
        }
        else if (mode == 3) { // texel division: denominators of the raster path, numerators down to zero / subnormals
            ub = (ub & 0x807FFFFFu) | ((87u + (ub >> 23) % 103u) << 23);
            ua = (ua & 0x807FFFFFu) | (((ua >> 23) % 190u) << 23);
        }
        const float a = __uint_as_float(ua), b = __uint_as_float(ub);
        bool ok = true;
        const float ref = a / b;
        cnt++;
        if (mode == 3) {
            const float q = vx_div_texel(a, b, ok);
            if (!ok) fb++;
            else if ((vx_f2i(q * 8.0f) & 7) != (vx_f2i(ref * 8.0f) & 7)) bad++;
            continue;
        }
        const float q = vx_div_fast(a, b, ok);
        if (!ok) fb++;
        else if (__float_as_uint(q) != __float_as_uint(ref) && !(q == 0.0f && ref == 0.0f)) bad++;
    }
    atomicAdd(&counters[0], bad);
    atomicAdd(&counters[1], fb);
    atomicAdd(&counters[2], cnt);
}

} // namespace

extern "C" {

int vx_horizon_cull(VxContext *ctx, const float cam_pos[3], const float *centers, int32_t n_centers, int32_t *order_inout,
                    int32_t n, int32_t bins, float base_margin, float margin_dist_factor, float min_dist_chunks,
                    int32_t *n_kept) {
    if (!ctx || !cam_pos || !n_kept || n < 0 || n_centers < 0 || bins <= 0 || (n > 0 && (!centers || !order_inout)))
        return vx_fail(ctx, VX_ERR_INVALID, "vx_horizon_cull: bad argument");
    *n_kept = 0;
    if (n == 0) return VX_OK;
    if (n > (1 << 16)) return vx_fail(ctx, VX_ERR_CAPACITY, "vx_horizon_cull: more than 65536 meshes");
    for (int32_t i = 0; i < n; ++i)
        if (order_inout[i] < 0 || order_inout[i] >= n_centers) return vx_fail(ctx, VX_ERR_INVALID, "vx_horizon_cull: mesh id out of range");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nn = (size_t)n;
    // scratch: centers | order | key | sorted | bin_of | slope | margin | top | keep | n_kept
    const size_t off_order = sizeof(float) * 3 * (size_t)n_centers;
    const size_t off_key = off_order + 4 * nn, off_sorted = off_key + 4 * nn, off_bin = off_sorted + 4 * nn;
    const size_t off_slope = off_bin + 4 * nn, off_margin = off_slope + 4 * nn, off_top = off_margin + 4 * nn;
    const size_t off_ang_in = off_top + 4 * nn, off_ang = off_ang_in + 4 * nn;
    const size_t off_keep = off_ang + 4 * nn, off_cnt = (off_keep + nn + 15) & ~(size_t)15;
    VX_CUDA(ctx, ctx->tmp_a.reserve(off_cnt + 16));
    uint8_t *base = ctx->tmp_a.as<uint8_t>();
    // xz.y.atan2(xz.x) (culling.rs:86) with the platform's atan2f, exactly as the reference evaluates it
    std::vector<float> ang(nn);
    for (size_t i = 0; i < nn; ++i) {
        const float *c = centers + 3 * (size_t)order_inout[i];
        const float tx = c[0] - cam_pos[0], tz = c[2] - cam_pos[2];
        ang[i] = atan2f(tz, tx);
    }
    VX_CUDA(ctx, cudaMemcpyAsync(base, centers, off_order, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(base + off_order, order_inout, 4 * nn, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(base + off_ang_in, ang.data(), 4 * nn, cudaMemcpyHostToDevice, ctx->stream));
    HorizonArgs a;
    a.angle_in = reinterpret_cast<const float *>(base + off_ang_in);
    a.angle = reinterpret_cast<float *>(base + off_ang);
    a.centers = reinterpret_cast<const float *>(base);
    a.order = reinterpret_cast<int32_t *>(base + off_order);
    a.n = n;
    a.bins = bins;
    a.cam[0] = cam_pos[0]; a.cam[1] = cam_pos[1]; a.cam[2] = cam_pos[2];
    a.base_margin = base_margin; a.margin_dist_factor = margin_dist_factor; a.min_dist_chunks = min_dist_chunks;
    a.key = reinterpret_cast<float *>(base + off_key);
    a.sorted = reinterpret_cast<int32_t *>(base + off_sorted);
    a.bin_of = reinterpret_cast<int32_t *>(base + off_bin);
    a.slope = reinterpret_cast<float *>(base + off_slope);
    a.margin = reinterpret_cast<float *>(base + off_margin);
    a.top = reinterpret_cast<float *>(base + off_top);
    a.keep = base + off_keep;
    a.n_kept = reinterpret_cast<int32_t *>(base + off_cnt);
    horizon_cull_kernel<<<1, HZ_THREADS, 0, ctx->stream>>>(a);
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(n_kept, a.n_kept, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(order_inout, a.order, 4 * nn, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_generate_terrain(VxContext *ctx, const int32_t *positions, int32_t n, const VxTerrainParams *params, uint8_t *d_voxels_out,
                        uint8_t *uniform_flags_out) {
    if (!ctx || n < 0 || !params || (n > 0 && (!positions || !d_voxels_out || !uniform_flags_out)))
        return vx_fail(ctx, VX_ERR_INVALID, "vx_generate_terrain: bad argument");
    if (n == 0) return VX_OK;
    if ((reinterpret_cast<uintptr_t>(d_voxels_out) & 15) != 0) return vx_fail(ctx, VX_ERR_INVALID, "voxels must be 16-byte aligned");
    for (int i = 0; i < 512; ++i)
        if (params->perm[i] < 0 || params->perm[i] > 255) return vx_fail(ctx, VX_ERR_INVALID, "vx_generate_terrain: permutation entries must be 0..255");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nn = (size_t)n;
    const size_t off_perm = sizeof(int32_t) * 3 * nn, off_flags = off_perm + sizeof(int32_t) * 512;
    VX_CUDA(ctx, ctx->tmp_a.reserve(off_flags + nn + 16));
    uint8_t *base = ctx->tmp_a.as<uint8_t>();
    VX_CUDA(ctx, cudaMemcpyAsync(base, positions, off_perm, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(base + off_perm, params->perm, sizeof(int32_t) * 512, cudaMemcpyHostToDevice, ctx->stream));
    TerrainArgs a;
    a.positions = reinterpret_cast<const int32_t *>(base);
    a.n = n;
    a.perm = reinterpret_cast<const int32_t *>(base + off_perm);
    memcpy(a.grad, params->grad, sizeof(a.grad));
    a.scale = params->scale;
    a.amplitude = params->amplitude;
    a.voxels = d_voxels_out;
    a.flags = base + off_flags;
    const int grid = n < ctx->num_sms * 8 ? n : ctx->num_sms * 8;
    generate_terrain_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(uniform_flags_out, a.flags, nn, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_face_packets(VxContext *ctx, const VxMeshBatch *batch, int32_t mesh_id, VxFacePacket32 *packets_out, int32_t cap_packets,
                    int32_t n_packets_per_face[6]) {
    if (!ctx || !batch || !n_packets_per_face || cap_packets < 0 || (cap_packets > 0 && !packets_out))
        return vx_fail(ctx, VX_ERR_INVALID, "vx_face_packets: bad argument");
    if (mesh_id < 0 || mesh_id >= batch->n_chunks) return vx_fail(ctx, VX_ERR_INVALID, "mesh id out of range");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t so[198], qb = 0, qc = 0;
    VX_CUDA(ctx, cudaMemcpyAsync(so, batch->slice_offsets.as<uint32_t>() + (size_t)mesh_id * 198, sizeof(so), cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(&qb, batch->quad_base.as<uint32_t>() + mesh_id, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(&qc, batch->quad_count.as<uint32_t>() + mesh_id, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int32_t total = 0;
    for (int f = 0; f < 6; ++f) {
        n_packets_per_face[f] = (int32_t)((so[f * 33 + 32] - so[f * 33] + 31u) / 32u);
        total += n_packets_per_face[f];
    }
    if (total > cap_packets) return vx_fail(ctx, VX_ERR_CAPACITY, "vx_face_packets: packet array too small (sizes are in n_packets_per_face)");
    if (total == 0) return VX_OK;
    VX_CUDA(ctx, ctx->tmp_a.reserve(sizeof(VxFacePacket32) * (size_t)total));
    VX_CUDA(ctx, cudaMemsetAsync(ctx->tmp_a.ptr, 0, sizeof(VxFacePacket32) * (size_t)total, ctx->stream));
    const int blocks = (int)((qc + 255u) / 256u);
    face_packets_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, ctx->stream>>>(batch->quads.as<uint8_t>(), batch->slice_offsets.as<uint32_t>() + (size_t)mesh_id * 198,
                                                                      qb, qc, ctx->tmp_a.as<VxFacePacket32>(), (uint32_t)total);
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(packets_out, ctx->tmp_a.ptr, sizeof(VxFacePacket32) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_selftest_division(VxContext *ctx, uint64_t seed, uint64_t n_pairs, int32_t mode, uint64_t counters_out[3]) {
    if (!ctx || !counters_out) return vx_fail(ctx, VX_ERR_INVALID, "vx_selftest_division: bad argument");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    VX_CUDA(ctx, ctx->tmp_a.reserve(3 * sizeof(unsigned long long)));
    VX_CUDA(ctx, cudaMemsetAsync(ctx->tmp_a.ptr, 0, 3 * sizeof(unsigned long long), ctx->stream));
    selftest_division_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(seed, n_pairs, mode, ctx->tmp_a.as<unsigned long long>());
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(counters_out, ctx->tmp_a.ptr, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}


int vx_cull_chunks(VxContext *ctx, const int32_t *positions, int32_t n, const float vp[16], const float cam_pos[3],
                   int32_t view_distance, int32_t frustum_culling, uint8_t *visible_out) {
    if (!ctx || n < 0 || !vp || !cam_pos || (n > 0 && (!positions || !visible_out))) return vx_fail(ctx, VX_ERR_INVALID, "vx_cull_chunks: bad argument");
    if (n == 0) return VX_OK;
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    VX_CUDA(ctx, ctx->tmp_a.reserve(sizeof(int32_t) * 3 * (size_t)n));
    VX_CUDA(ctx, ctx->tmp_b.reserve((size_t)n));
    VX_CUDA(ctx, cudaMemcpyAsync(ctx->tmp_a.ptr, positions, sizeof(int32_t) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int grid = (n + 255) / 256;
    if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
    cull_chunks_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->tmp_a.as<int32_t>(), n, to_mat(vp), cam_pos[0], cam_pos[1], cam_pos[2],
                                                     view_distance, frustum_culling, ctx->tmp_b.as<uint8_t>());
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(visible_out, ctx->tmp_b.ptr, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_face_basis(VxContext *ctx, const int32_t *faces, const int32_t *chunk_pos, const uint8_t *slice_idx, int32_t n,
                  const float vp[16], float *basis_out) {
    if (!ctx || n < 0 || !vp || (n > 0 && (!faces || !chunk_pos || !slice_idx || !basis_out))) return vx_fail(ctx, VX_ERR_INVALID, "vx_face_basis: bad argument");
    if (n == 0) return VX_OK;
    for (int32_t i = 0; i < n; ++i)
        if (faces[i] < 0 || faces[i] > 5) return vx_fail(ctx, VX_ERR_INVALID, "face index out of range");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)n;
    VX_CUDA(ctx, ctx->tmp_a.reserve(4 * N));
    VX_CUDA(ctx, ctx->tmp_b.reserve(12 * N));
    VX_CUDA(ctx, ctx->tmp_c.reserve(N));
    VX_CUDA(ctx, ctx->tmp_d.reserve(64 * N));
    VX_CUDA(ctx, cudaMemcpyAsync(ctx->tmp_a.ptr, faces, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(ctx->tmp_b.ptr, chunk_pos, 12 * N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(ctx->tmp_c.ptr, slice_idx, N, cudaMemcpyHostToDevice, ctx->stream));
    face_basis_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->tmp_a.as<int32_t>(), ctx->tmp_b.as<int32_t>(), ctx->tmp_c.as<uint8_t>(), n,
                                                               to_mat(vp), ctx->tmp_d.as<float>());
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(basis_out, ctx->tmp_d.ptr, 64 * N, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_project_packet(VxContext *ctx, const float basis[16], const uint8_t *u_min, const uint8_t *v_min,
                      const uint8_t *u_len, const uint8_t *v_len, int32_t n, float *x_min, float *y_min,
                      float *x_max, float *y_max, float *depth_near) {
    if (!ctx || n < 0 || !basis) return vx_fail(ctx, VX_ERR_INVALID, "vx_project_packet: bad argument");
    if (n == 0) return VX_OK;
    if (!u_min || !v_min || !u_len || !v_len || !x_min || !y_min || !x_max || !y_max || !depth_near) return vx_fail(ctx, VX_ERR_INVALID, "vx_project_packet: null array");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)n;
    VX_CUDA(ctx, ctx->tmp_a.reserve(4 * N));
    VX_CUDA(ctx, ctx->tmp_b.reserve(5 * 4 * N));
    uint8_t *d_in = ctx->tmp_a.as<uint8_t>();
    VX_CUDA(ctx, cudaMemcpyAsync(d_in + 0 * N, u_min, N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_in + 1 * N, v_min, N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_in + 2 * N, u_len, N, cudaMemcpyHostToDevice, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(d_in + 3 * N, v_len, N, cudaMemcpyHostToDevice, ctx->stream));
    Basis bs;
    memcpy(bs.v, basis, sizeof(bs.v));
    project_packet_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(bs, d_in, d_in + N, d_in + 2 * N, d_in + 3 * N, n, ctx->tmp_b.as<float>());
    VX_CHECK_LAUNCH(ctx);
    float *outs[5] = {x_min, y_min, x_max, y_max, depth_near};
    for (int k = 0; k < 5; ++k)
        VX_CUDA(ctx, cudaMemcpyAsync(outs[k], ctx->tmp_b.as<float>() + k * N, 4 * N, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_transform_vertices(VxContext *ctx, const VxVertex *verts, int32_t n, const float offset[3], const float vp[16],
                          float *out4) {
    if (!ctx || n < 0 || !offset || !vp || (n > 0 && (!verts || !out4))) return vx_fail(ctx, VX_ERR_INVALID, "vx_transform_vertices: bad argument");
    if (n == 0) return VX_OK;
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)n;
    VX_CUDA(ctx, ctx->tmp_a.reserve(8 * N));
    VX_CUDA(ctx, ctx->tmp_b.reserve(16 * N));
    VX_CUDA(ctx, cudaMemcpyAsync(ctx->tmp_a.ptr, verts, 8 * N, cudaMemcpyHostToDevice, ctx->stream));
    int grid = (n + 255) / 256;
    if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
    transform_vertices_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->tmp_a.as<VxVertex>(), n, offset[0], offset[1], offset[2], to_mat(vp), ctx->tmp_b.as<float4>());
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(out4, ctx->tmp_b.ptr, 16 * N, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

int vx_project_mesh_vertices(VxContext *ctx, const VxMeshBatch *batch, int32_t mesh_id, const float vp[16],
                             int32_t differential, float *out, int64_t cap_quads) {
    if (!ctx || !batch || !vp || !out || mesh_id < 0 || mesh_id >= batch->n_chunks) return vx_fail(ctx, VX_ERR_INVALID, "vx_project_mesh_vertices: bad argument");
    VX_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t qbase = 0, qcount = 0;
    int32_t pos[3];
    VX_CUDA(ctx, cudaMemcpyAsync(&qbase, batch->quad_base.as<uint32_t>() + mesh_id, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(&qcount, batch->quad_count.as<uint32_t>() + mesh_id, 4, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaMemcpyAsync(pos, batch->positions.as<int32_t>() + 3 * (size_t)mesh_id, 12, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if ((int64_t)qcount > cap_quads) return vx_fail(ctx, VX_ERR_CAPACITY, "output too small for the mesh's quads");
    if (qcount == 0) return VX_OK;
    VX_CUDA(ctx, ctx->tmp_a.reserve(64 * (size_t)qcount));
    int grid = (int)((qcount + 127) / 128);
    project_mesh_vertices_kernel<<<grid, 128, 0, ctx->stream>>>(batch->quads.as<uint8_t>(), batch->slice_offsets.as<uint32_t>() + (size_t)mesh_id * 198,
                                                               qbase, qcount, (float)(pos[0] * VX_CHUNK_SIZE), (float)(pos[1] * VX_CHUNK_SIZE),
                                                               (float)(pos[2] * VX_CHUNK_SIZE), to_mat(vp), differential, ctx->tmp_a.as<float4>());
    VX_CHECK_LAUNCH(ctx);
    VX_CUDA(ctx, cudaMemcpyAsync(out, ctx->tmp_a.ptr, 64 * (size_t)qcount, cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VX_OK;
}

} // extern "C"

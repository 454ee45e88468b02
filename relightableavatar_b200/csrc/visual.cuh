// Image side of the path: the per-type map transforms of Visualizer.generate_image
// (lib/visualizers/base_visualizer.py:55-180), its ray -> image scatter / light-probe overlay / alpha channel (:182-202,
// lib/utils/relight_utils.py:38-52) and save_image's channel swap + 8 / 16-bit quantisation (lib/utils/data_utils.py:689-709),
// so that a frame leaves the device as finished pixels in ONE copy instead of one fp32 map per light and type.
#pragma once
#include "common.cuh"
#include "render.cuh"

// Output types RA_VIS_* (lib/config/config.py:364-378): include/ra_b200.h.  Semantic / Feature are marked deprecated in the
// reference and are not offered.
#include "../../include/ra_b200.h"

// ---- exact k-th order statistic (the reference's "simple version of percentile": topk(k)[0].max() / .min()) --------------
// Radix select over the order-preserving integer image of the floats, 8 bits per pass; every pass is one grid-wide
// histogram (shared-memory bins flushed with one atomic per bin and block) and a one-block pick.
struct KthState { unsigned prefix, maskbits, k_rem, hist[256]; float result; int n_valid; };

__device__ __forceinline__ unsigned kth_key(float v, int largest) {
    unsigned u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // ascending order of floats == ascending order of keys
    return largest ? ~u : u;
}
__device__ __forceinline__ float kth_unkey(unsigned k, int largest) {
    if (largest) k = ~k;
    k = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(k);
}
// element i of the selection domain; mode 0: a[i]; mode 1: a[i] - b[i]; mode 2: a[i] where b[i] != 0 (depth over acc.bool())
__device__ __forceinline__ bool kth_elem(int mode, const float* __restrict__ a, const float* __restrict__ b, long long i, float& v) {
    if (mode == 0) { v = a[i]; return true; }
    if (mode == 1) { v = a[i] - b[i]; return true; }
    v = a[i];
    return b[i] != 0.f;
}
__global__ void k_kth_init(KthState* s, unsigned k) {
    if (threadIdx.x < 256) s->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s->prefix = 0; s->maskbits = 0; s->k_rem = k; s->result = 0.f; s->n_valid = 0; }
}
__global__ void k_kth_hist(KthState* s, int pass, int mode, int largest, const float* __restrict__ a, const float* __restrict__ b, long long n) {
    __shared__ unsigned h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const unsigned prefix = s->prefix, maskbits = s->maskbits;
    const int shift = 24 - 8 * pass;
    int valid = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v;
        if (!kth_elem(mode, a, b, i, v)) continue;
        valid++;
        const unsigned key = kth_key(v, largest);
        if ((key & maskbits) == prefix) atomicAdd(&h[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(&s->hist[i], h[i]);
    if (pass == 0 && valid) atomicAdd(&s->n_valid, valid);
}
__global__ void k_kth_pick(KthState* s, int pass, int largest) {      // <<<1, 32>>>
    if (threadIdx.x == 0) {
        const int shift = 24 - 8 * pass;
        unsigned k = s->k_rem, cum = 0;
        int bin = 255;
        for (int i = 0; i < 256; i++) {
            if (cum + s->hist[i] >= k) { bin = i; break; }
            cum += s->hist[i];
        }
        s->k_rem = k - cum;
        s->prefix |= (unsigned)bin << shift;
        s->maskbits |= 0xffu << shift;
        if (pass == 3) s->result = kth_unkey(s->prefix, largest);
    }
    __syncwarp();
    for (int i = threadIdx.x; i < 256; i += 32) s->hist[i] = 0;
}

// ---- per-type map transform: ray-ordered (n, 3) visual map ---------------------------------------------------------------
struct VisualIn {
    const float *rgb, *acc, *norm, *depth, *shade, *albedo, *rough, *cpts, *bpts, *surf, *spec;   // (n,3) or (n); unused ones NULL
    const float* cam_R;      // (3,3) world -> camera, device
    const float* tbounds;    // (2,3) big-pose bounds, device
    const KthState *lo, *hi; // selected order statistics (Depth: both; Shading / Specular / Residual: hi only)
    float min_clip;          // cfg.min_clip
    int normalize;           // cfg.normalize_shading / cfg.normalize_specular
    int tonemap_albedo;      // cfg.tonemapping_albedo
};

__global__ void k_visual_map(int type, VisualIn in, long long n, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float3 v = make3(0.f, 0.f, 0.f);
        const float a = in.acc ? in.acc[i] : 1.f;
        switch (type) {
        case RA_VIS_RENDERING: v = make3(in.rgb[i * 3], in.rgb[i * 3 + 1], in.rgb[i * 3 + 2]); break;
        case RA_VIS_NORMAL: {      // normalize, world -> camera (n @ R^T), flip y / z, to [0,1], times acc   (:58-66)
            const float3 nw = normalize_ref(make3(in.norm[i * 3], in.norm[i * 3 + 1], in.norm[i * 3 + 2]));
            const float* R = in.cam_R;
            float3 c = make3(nw.x * R[0] + nw.y * R[1] + nw.z * R[2], nw.x * R[3] + nw.y * R[4] + nw.z * R[5], nw.x * R[6] + nw.y * R[7] + nw.z * R[8]);
            c.y *= -1.f; c.z *= -1.f;
            v = make3((c.x * 0.5f + 0.5f) * a, (c.y * 0.5f + 0.5f) * a, (c.z * 0.5f + 0.5f) * a);
        } break;
        case RA_VIS_ALPHA: v = make3(a, a, a); break;
        case RA_VIS_DEPTH: {       // percentile window over the foreground, min clipped from above by cfg.min_clip   (:101-116)
            const float dmin = fminf(in.lo->result, in.min_clip), dmax = in.hi->result;
            const float d = clampf((in.depth[i] - dmin) / (dmax - dmin), 0.f, 1.f);
            v = make3(d, d, d);
        } break;
        case RA_VIS_SHADING: case RA_VIS_SPECULAR: {      // (:118-127, :166-176)
            const float* m = type == RA_VIS_SHADING ? in.shade : in.spec;
            v = make3(m[i * 3], m[i * 3 + 1], m[i * 3 + 2]);
            if (in.normalize) { const float s = in.hi->result; v = make3(v.x / s, v.y / s, v.z / s); }
        } break;
        case RA_VIS_ALBEDO:
            v = make3(in.albedo[i * 3], in.albedo[i * 3 + 1], in.albedo[i * 3 + 2]);
            if (in.tonemap_albedo) v = make3(linear2srgb(v.x), linear2srgb(v.y), linear2srgb(v.z));
            break;
        case RA_VIS_ROUGHNESS: { const float r = in.rough[i]; v = make3(r, r, r); } break;
        case RA_VIS_SURFACE: {     // canonical (or world) surface point in the big-pose box, times acc   (:139-143)
            const float* p = in.cpts ? in.cpts : in.surf;
            const float* tb = in.tbounds;
            v = make3((p[i * 3] - tb[0]) / (tb[3] - tb[0]) * a, (p[i * 3 + 1] - tb[1]) / (tb[4] - tb[1]) * a, (p[i * 3 + 2] - tb[2]) / (tb[5] - tb[2]) * a);
        } break;
        case RA_VIS_RESIDUAL: {    // (cpts - bpts) / its 99.5th percentile, times acc   (:145-154)
            const float s = in.hi->result;
            v = make3((in.cpts[i * 3] - in.bpts[i * 3]) / s * a, (in.cpts[i * 3 + 1] - in.bpts[i * 3 + 1]) / s * a, (in.cpts[i * 3 + 2] - in.bpts[i * 3 + 2]) / s * a);
        } break;
        }
        out[i * 3] = v.x; out[i * 3 + 1] = v.y; out[i * 3 + 2] = v.z;
    }
}

// ---- ray -> image scatter, light-probe overlay, alpha channel, channel order, quantisation --------------------------------
// img = bg; img[mask_at_box] = map (:182-187); the top-left uH x uW pixels show the env-map probe sampled along `probe_dirs`
// (add_light_probe, relight_utils.py:38-52); alpha[mask_at_box] = acc (:193-200).  save_image (data_utils.py:689-709):
// BGR order for cv2, png -> (v * 65535).clip(0, 65535) as uint16, jpg -> 3 channels, (v * 255).clip(0, 255) as uint8.
struct AssembleArgs {
    const unsigned char* mask; int n, W;
    const int* blk_off;
    const float* map;        // (P,3) in ray order
    const float* acc;        // (P) or NULL (no alpha source: alpha 0)
    float bg;
    int C;                   // output channels: 3 or 4 (cfg.store_alpha_channel)
    int bgr;                 // swap R / B (save_image)
    const float* probe; int eh, ew;            // overlay source or NULL
    const float* probe_dirs; int uH, uW;       // (uH,uW,3) world-space directions (gen_light_dir)
    float* out_f; unsigned char* out_u8; unsigned short* out_u16;
};

__global__ void k_assemble2(AssembleArgs a) {
    __shared__ int wcnt[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool m = i < a.n && a.mask[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0) wcnt[wid] = __popc(bal);
    __syncthreads();
    int base = a.blk_off[blockIdx.x];
    for (int w = 0; w < wid; w++) base += wcnt[w];
    const int ray = base + __popc(bal & ((1u << lane) - 1u));
    if (i >= a.n) return;
    float v[4] = {a.bg, a.bg, a.bg, 0.f};
    if (m) {
        v[0] = a.map[(size_t)ray * 3]; v[1] = a.map[(size_t)ray * 3 + 1]; v[2] = a.map[(size_t)ray * 3 + 2];
        if (a.acc) v[3] = a.acc[ray];
    }
    const int y = i / a.W, x = i % a.W;
    if (a.probe && y < a.uH && x < a.uW) {
        const float* d = a.probe_dirs + ((size_t)y * a.uW + x) * 3;
        const float3 c = envmap_fetch(a.probe, a.eh, a.ew, make3(d[0], d[1], d[2]));
        v[0] = c.x; v[1] = c.y; v[2] = c.z;
    }
    if (a.bgr) { const float t = v[0]; v[0] = v[2]; v[2] = t; }
    for (int c = 0; c < a.C; c++) {
        const size_t o = (size_t)i * a.C + c;
        if (a.out_f) a.out_f[o] = v[c];
        if (a.out_u8) a.out_u8[o] = (unsigned char)clampf(v[c] * 255.f, 0.f, 255.f);
        if (a.out_u16) a.out_u16[o] = (unsigned short)clampf(v[c] * 65535.f, 0.f, 65535.f);
    }
}

// rotate_envmap's shift_image for an image of any size (the floor's attached env-map image, relight_utils.py:74-75,103):
// out[r] = image shifted by step * (j0 + r) texels along the longitude, wrap-around sample position, border clamp.
__global__ void k_shift_image(const float* __restrict__ img, int H, int W, double step, int j0, int n_rot, float* out) {
    const long long total = (long long)n_rot * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), r = (int)(i / ((long long)W * H));
        const float shift = (float)(step * (double)(j0 + r));          // python float -> fp32 scalar added to the fp32 grid
        const float gx = fmodf((float)x + 0.5f + shift, (float)W);
        const float nx = gx / (float)W * 2.f - 1.f;
        float ix = ((nx + 1.f) * W - 1.f) / 2.f;
        ix = clampf(ix, 0.f, (float)(W - 1));
        const int x0 = (int)floorf(ix);
        const float fx = ix - x0;
        const int x1 = min(x0 + 1, W - 1);
        const float w1 = (x0 + 1 > W - 1) ? 0.f : fx;
        // the row coordinate goes through the same normalise / un-normalise round trip: (y + .5) / H * 2 - 1 -> iy
        const float ny = ((float)y + 0.5f) / (float)H * 2.f - 1.f;
        float iy = ((ny + 1.f) * H - 1.f) / 2.f;
        iy = clampf(iy, 0.f, (float)(H - 1));
        const int y0 = (int)floorf(iy);
        const float fy = iy - y0;
        const int y1 = min(y0 + 1, H - 1);
        const float wy1 = (y0 + 1 > H - 1) ? 0.f : fy;
        const float* p00 = img + ((size_t)y0 * W + x0) * 3; const float* p01 = img + ((size_t)y0 * W + x1) * 3;
        const float* p10 = img + ((size_t)y1 * W + x0) * 3; const float* p11 = img + ((size_t)y1 * W + x1) * 3;
        for (int c = 0; c < 3; c++)
            out[(size_t)i * 3 + c] = (p00[c] * (1.f - fx) + p01[c] * w1) * (1.f - fy) + (p10[c] * (1.f - fx) + p11[c] * w1) * wy1;
    }
}

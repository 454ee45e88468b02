// Ground-plane shading: the vis_ground_shading branch of the reference renderer.
// Reference: sphere_tracing_renderer.py:463-548 (render_ground), :1079-1111 (Renderer.render ground branch), :395-451
// (alpha_blend / blend_output_); novel_light_sphere_tracing.py:69-98 (floor re-shade per env-map);
// mesh_utils.py:710-738 (moller_trumbore), net_utils.py:392-396 (compute_ground_tris).
// The floor's soft shadows reuse the shadow-ray list and k_trace_shadow with the cfg.env_lvis parameters
// (16 iterations, dist_th 5 mm: nearly every query is answered by the SMPL distance alone).
#pragma once
#include "render.cuh"

struct GroundCfg {
    float normal[3], origin[3], albedo[3];
    int attach_envmap;
    float shading_albedo, multiplier, env_r, near_offset, bbox_margin;
    int tonemap;
};

// image pixel -> ray index (or -1): the `inds` of the reference (mask.nonzero() order); also acc_g = 1 - acc_human
__global__ void k_ground_pix2ray(const unsigned char* __restrict__ mask, int n, const int* __restrict__ blk_off, const float* __restrict__ acc_ray,
                                 int* pix2ray, float* acc_g) {
    __shared__ int wcnt[32];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool m = i < n && mask[i] != 0;
    unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0) wcnt[wid] = __popc(bal);
    __syncthreads();
    int base = blk_off[blockIdx.x];
    for (int w = 0; w < wid; w++) base += wcnt[w];
    int ray = base + __popc(bal & ((1u << lane) - 1u));
    if (i >= n) return;
    pix2ray[i] = m ? ray : -1;
    if (acc_g) acc_g[i] = m ? 1.f - acc_ray[ray] : 1.f;
}

// per light: radiance of the probe along normalize(xyz_l) (the same for every floor pixel) times nothing else; (L,3)
__global__ void k_ground_light_table(const float* __restrict__ lxyz, int L, const float* __restrict__ probe, int eh, int ew, float* table) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
        float3 d = normalize_ref(make3(lxyz[l * 3], lxyz[l * 3 + 1], lxyz[l * 3 + 2]));
        float3 li = envmap_fetch(probe, eh, ew, d);
        table[l * 3] = li.x; table[l * 3 + 1] = li.y; table[l * 3 + 2] = li.z;
    }
}

// per pixel: plane hit (moller_trumbore's t against the ground triangle), surf, depth, far-field blend weight, albedo
__global__ void k_ground_setup(GroundCfg g, const float* __restrict__ ray_o, const float* __restrict__ ray_d, long long p0, long long n,
                               const float* __restrict__ albedo_img, int ih, int iw,
                               float* surf, float* depth, float* norm, float* albedo, float* rough, float* weight) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        long long i = p0 + k;
        float3 o = make3(ray_o[i * 3], ray_o[i * 3 + 1], ray_o[i * 3 + 2]);
        float3 d = make3(ray_d[i * 3], ray_d[i * 3 + 1], ray_d[i * 3 + 2]);
        float3 nn = normalize_ref(make3(g.normal[0], g.normal[1], g.normal[2]));
        float3 og = make3(g.origin[0], g.origin[1], g.origin[2]);
        // N = E1 x E2 = |a|^2 n with the reference's random tangent; |a|^2 := 1 (it only scales the 1e-8, see the oracle)
        float invdet = 1.f / -((d.x * nn.x + d.y * nn.y + d.z * nn.z) + 1e-8f);
        float t = ((o.x - og.x) * nn.x + (o.y - og.y) * nn.y + (o.z - og.z) * nn.z) * invdet;
        float3 s = o + d * t;
        surf[i * 3] = s.x; surf[i * 3 + 1] = s.y; surf[i * 3 + 2] = s.z;
        if (depth) depth[i] = clampf(t, -g.env_r, g.env_r);
        if (norm) { norm[i * 3] = nn.x; norm[i * 3 + 1] = nn.y; norm[i * 3 + 2] = nn.z; }
        if (rough) rough[i] = 1.f;
        float3 e = s - og;
        float dist = (t <= 0.f) ? 1e9f : sqrtf(e.x * e.x + e.y * e.y + e.z * e.z);
        weight[i] = clampf((dist - g.env_r) / g.env_r, 0.f, 1.f);
        float3 al = make3(g.albedo[0], g.albedo[1], g.albedo[2]);
        if (g.attach_envmap) al = envmap_fetch(albedo_img, ih, iw, d);
        albedo[i * 3] = al.x; albedo[i * 3 + 1] = al.y; albedo[i * 3 + 2] = al.z;
    }
}

// visibility of the (pixel, light) pairs before tracing: 1 where the light is in front of the floor and the pixel shows floor (rays that
// miss the body's box stay at 1, traced rays overwrite theirs), else 0.  Pixel-major, coalesced: k_ground_rays walks the pairs in packet
// order (32 pixels of one light per warp), where these stores would be 2 KB apart.
__global__ void k_ground_vis_init(GroundCfg g, const float* __restrict__ acc_g, long long p0, long long n, const float* __restrict__ ldir, int L, float* lvis) {
    const float3 nn = normalize_ref(make3(g.normal[0], g.normal[1], g.normal[2]));
    const long long total = n * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = p0 + i / L; const int l = (int)(i % L);
        const float dt = ldir[l * 3] * nn.x + ldir[l * 3 + 1] * nn.y + ldir[l * 3 + 2] * nn.z;
        lvis[(size_t)pix * L + l] = (dt > 0.f && acc_g[pix] > 0.f) ? 1.f : 0.f;
    }
}

// light_visibility set-up for the floor (:265-329): per (pixel, light) front-facing & box tests, shadow-ray append.
// `pad_chunks0`: how many pixel chunks the human pass already grew wbounds by; the floor's own chunk index comes from the
// pixel index (`chunk_actual` = the reference's equalised chunk size over the H*W pixels).
__global__ void k_ground_rays(const FrameConst* __restrict__ fc, GroundCfg g, const float* __restrict__ surf, const float* __restrict__ acc_g,
                              long long p0, long long n, const float* __restrict__ ldir, int L, int pad_chunks0, int chunk_actual,
                              ShadowRays sr, int* n_shadow, int packets,
                              int tile_w /* image width when a packet is an 8 x 4 pixel tile (batch = whole groups of 4 rows), 0: 32 consecutive pixels */) {
    const int lane = threadIdx.x & 31;
    const float3 nn = normalize_ref(make3(g.normal[0], g.normal[1], g.normal[2]));
    if (packets) {
        // One warp per TILE of 32 pixels (lane = pixel), all lights: a packet = the same light for the tile's pixels, a whole aligned
        // 32-entry block of the ray list (lanes without a ray carry fg = -1, see k_shadow_gen).  The packets of 32 lights are counted first
        // and the list grows by one atomic per such group (one atomic per packet -- 0.8 M on one counter at 512^2 -- took 2 ms).  NOT one
        // atomic per tile: that lays each tile's ~100 packets out contiguously, the tracer's blocks then hold either only expensive packets
        // (tiles under the body) or only cheap ones, and the floor pass measured 4 ms SLOWER; groups of a few packets from the ~10 k warps
        // in flight interleave in the list and balance the tracer's blocks.
        const long long ntile = (n + 31) >> 5;
        const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long tile = warp0; tile < ntile; tile += nwarps) {
            long long k = tile * 32 + lane;
            if (tile_w) {
                const long long per_row = tile_w >> 3;
                k = ((tile / per_row) * 4 + (lane >> 3)) * tile_w + (tile % per_row) * 8 + (lane & 7);
            }
            const long long pix = p0 + k;
            const bool floor_px = k < n && acc_g[pix] > 0.f;
            if (!__any_sync(0xffffffffu, floor_px)) continue;
            float3 o = make3(0.f, 0.f, 0.f);
            float bmin[3] = {0.f, 0.f, 0.f}, bmax[3] = {0.f, 0.f, 0.f};
            if (floor_px) {
                const float pad = g.bbox_margin * (float)(pad_chunks0 + 1 + (int)(pix / chunk_actual));
                for (int a = 0; a < 3; a++) { bmin[a] = fc->wb[a] - pad; bmax[a] = fc->wb[3 + a] + pad; }
                o = make3(surf[pix * 3], surf[pix * 3 + 1], surf[pix * 3 + 2]);
            }
            auto test = [&](int l, float& nr, float& fr) -> bool {       // this lane's ray towards light l: does it cross the box?
                const float3 dl = make3(__ldg(&ldir[l * 3]), __ldg(&ldir[l * 3 + 1]), __ldg(&ldir[l * 3 + 2]));
                if (!(dl.x * nn.x + dl.y * nn.y + dl.z * nn.z > 0.f) || !floor_px) return false;
                aabb_near_far(bmin, bmax, o, dl, nr, fr);
                nr = fmaxf(nr, g.near_offset); fr = fmaxf(fr, g.near_offset);
                return nr < fr;
            };
            for (int l0 = 0; l0 < L; l0 += 32) {           // groups of 32 lights: see the note on list order above
                const int l1 = min(l0 + 32, L);
                int n_pk = 0;
                for (int l = l0; l < l1; l++) {
                    float nr, fr;
                    if (__any_sync(0xffffffffu, test(l, nr, fr))) n_pk++;
                }
                if (n_pk == 0) continue;
                int base = 0;
                if (lane == 0) base = atomicAdd(n_shadow, 32 * n_pk);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int l = l0; l < l1; l++) {
                    float nr = 0.f, fr = 0.f;
                    const bool trace = test(l, nr, fr);
                    const unsigned any = __ballot_sync(0xffffffffu, trace);
                    if (!any) continue;
                    const int slot = base + lane;
                    base += 32;
                    if (slot >= sr.cap) { if (trace) atomicAdd(sr.dropped, 1); continue; }
                    sr.fg[slot] = trace ? (int)pix : -1; sr.light[slot] = (unsigned short)l; sr.near_[slot] = nr; sr.far_[slot] = fr;
                }
            }
        }
        return;
    }
    // legacy order: one warp = 32 lights of one pixel, rays appended compacted
    const long long total = n * L;
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < total; base += (long long)gridDim.x * blockDim.x) {
        const int l = (int)((base + lane) % L);
        const long long k = (base + lane) / L;
        const bool valid = base + lane < total;
        bool trace = false;
        const long long pix = p0 + k;
        float nr = 0.f, fr = 0.f;
        if (valid) {
            float3 dl = make3(ldir[l * 3], ldir[l * 3 + 1], ldir[l * 3 + 2]);
            float dt = dl.x * nn.x + dl.y * nn.y + dl.z * nn.z;
            if (dt > 0.f && acc_g[pix] > 0.f) {
                float pad = g.bbox_margin * (float)(pad_chunks0 + 1 + (int)(pix / chunk_actual));
                float bmin[3] = {fc->wb[0] - pad, fc->wb[1] - pad, fc->wb[2] - pad};
                float bmax[3] = {fc->wb[3] + pad, fc->wb[4] + pad, fc->wb[5] + pad};
                float3 o = make3(surf[pix * 3], surf[pix * 3 + 1], surf[pix * 3 + 2]);
                aabb_near_far(bmin, bmax, o, dl, nr, fr);
                nr = fmaxf(nr, g.near_offset); fr = fmaxf(fr, g.near_offset);
                trace = nr < fr;          // (misses the box: stays visible, k_ground_vis_init)
            }
        }
        shadow_append(sr, n_shadow, false, trace, (int)pix, l, nr, fr);
    }
}

// One warp per floor pixel.  first_pass: blend the traced visibility towards 1 in the far field, replace ldot by n.l for
// EVERY light (the reference recomputes it unmasked), store both maps, then the Lambertian light sum.
// Re-shade (first_pass == 0): the stored maps are used as they are (novel_light_sphere_tracing.py:69-98).
__global__ void k_ground_shade(GroundCfg g, int first_pass, long long p0, long long n, const float* __restrict__ weight,
                               const float* __restrict__ ldir, const float* __restrict__ larea, int L, const float* __restrict__ light_table,
                               float* lvis, float* ldot, const float* __restrict__ albedo, float shade_scale, float shade_map_mult,
                               float* rgb, float* shade, float* spec) {
    const float PI = 3.14159265358979323846f;
    int lane = threadIdx.x & 31;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float3 nn = normalize_ref(make3(g.normal[0], g.normal[1], g.normal[2]));
    for (long long k = warp; k < n; k += nwarps) {
        long long i = p0 + k;
        float w = first_pass ? weight[i] : 0.f;
        float cs[3] = {0.f, 0.f, 0.f};
        for (int l = lane; l < L; l += 32) {
            size_t o = (size_t)i * L + l;
            float lv = lvis[o], ld;
            if (first_pass) {
                lv = lv * (1.f - w) + 1.f * w;
                ld = ldir[l * 3] * nn.x + ldir[l * 3 + 1] * nn.y + ldir[l * 3 + 2] * nn.z;
                lvis[o] = lv; ldot[o] = ld;
            } else ld = ldot[o];
            float s = lv * ld * larea[l];
            cs[0] += s * light_table[l * 3]; cs[1] += s * light_table[l * 3 + 1]; cs[2] += s * light_table[l * 3 + 2];
        }
        for (int c = 0; c < 3; c++) cs[c] = warp_sum(cs[c]);
        if (lane == 0)
            for (int c = 0; c < 3; c++) {
                if (rgb) rgb[i * 3 + c] = tone(albedo[i * 3 + c] / PI * cs[c], first_pass ? g.tonemap : 1);      // the re-shade always maps (novel_light_sphere_tracing.py:94)
                float sh = cs[c] * shade_scale / PI;
                if (shade) shade[i * 3 + c] = sh * shade_map_mult;
                if (spec) spec[i * 3 + c] = sh / 20.f;
            }
    }
}

// floor albedo of a novel env-map (ground_attach_envmap): the probe's colour along the view ray
__global__ void k_ground_albedo(const float* __restrict__ ray_d, long long n, const float* __restrict__ img, int ih, int iw, float* albedo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float3 al = envmap_fetch(img, ih, iw, make3(ray_d[i * 3], ray_d[i * 3 + 1], ray_d[i * 3 + 2]));
        albedo[i * 3] = al.x; albedo[i * 3 + 1] = al.y; albedo[i * 3 + 2] = al.z;
    }
}

// blend_output_ / alpha_blend (:395-451): out = ground * acc_g + scatter(human) * (1 - acc_g), image-sized, C channels.
// human == nullptr: alpha_times (key only in the ground dict).  ground == nullptr: the acc_map rule (target = 0).
// human_premul: the human map is already multiplied by acc_h = 1 - acc_g (the plain maps of ra_render_relight).
__global__ void k_ground_blend(const int* __restrict__ pix2ray, const float* __restrict__ acc_g, const float* __restrict__ ground,
                               const float* __restrict__ human, int human_premul, int C, long long n, float* out) {
    long long total = n * C;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
        long long i = j / C; int c = (int)(j % C);
        float a = acc_g[i];
        float v = ground ? ground[j] * a : 0.f;
        int ray = pix2ray[i];
        if (human && ray >= 0) v += human[(size_t)ray * C + c] * (human_premul ? 1.f : (1.f - a));
        out[j] = v;
    }
}

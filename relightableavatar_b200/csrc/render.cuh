// Ray-level kernels: fixed-iteration sphere tracing (surface + soft shadows), shadow-ray generation,
// surface sampling / blending, microfacet light sum.  No host synchronisation anywhere: every
// data-dependent size is a device-side counter and every kernel is a grid-stride loop over it.
// Reference: sphere_tracing_renderer.py:103-216 (tracing), :265-344 (visibility), :551-784 (render_human),
// relight_utils.py:106-127, 179-192, 468-633 (env-map, sRGB, BRDF).
#pragma once
#include "common.cuh"
#include "hdq.cuh"

struct TraceCfg {
    int iters; float tan_i, relax, offset, eps; int skip;
    float th, blend_radius;
};

// Distance-query work list shared by all tracing stages: in-shell points waiting for the MLPs.
struct QueryList {
    float* bpts;      // [cap][3]
    float* net;       // [cap] network sdf written by the MLP stage
    int* count;       // device counter
};

struct SurfState {   // SoA over the P rays of a render call
    float *t, *occ, *d0, *cd, *dt, *st, *off, *rlx, *q_smpl;
    int* q_slot;
};

struct Counters {
    int* n_fg; int* n_shadow; int* n_attr;
    unsigned long long* n_queries; unsigned long long* n_inshell;
};

__device__ __forceinline__ void count_queries(const Counters& c, bool valid, bool ins) {
    unsigned mv = __ballot_sync(0xffffffffu, valid), mi = __ballot_sync(0xffffffffu, ins);
    if ((threadIdx.x & 31) == 0) {
        if (mv) atomicAdd(c.n_queries, (unsigned long long)__popc(mv));
        if (mi) atomicAdd(c.n_inshell, (unsigned long long)__popc(mi));
    }
}

// One launch = finish tracing iteration `it-1` with the MLP results, then start iteration `it`
// (or finalise when it == iters).  Hard (surface) mode of A.4.
__global__ void __launch_bounds__(256, RA_TRACE_MINBLOCKS) k_trace_surface(int it, TraceCfg cfg, const FrameConst* __restrict__ fc, SortedVerts sv, int nverts,
                                const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                                const float* __restrict__ near_, const float* __restrict__ far_, int P,
                                SurfState s, QueryList q, Counters cnt,
                                // finalisation outputs (it == iters)
                                float* surf, float* acc, float* depth, int* fg_ray, int packets) {
    for (int base = blockIdx.x * blockDim.x; base < P; base += gridDim.x * blockDim.x) {      // block-uniform trip count
        int i = base + threadIdx.x;
        bool valid = i < P;
        float3 o = make3(0, 0, 0), d = make3(0, 0, 1);
        float nr = 0.f, fr = 0.f, t = 0.f, occ = 1.f, d0 = 1e9f, cd = 1e9f, dt = 1e9f, st = 0.f, off = cfg.offset, rlx = cfg.relax;
        // A ray whose front did not move (t clamped at near / far: every ray that misses the body ends up parked at
        // `far`) would query the very same point again and get the very same distance: reuse it (exact), skip the query.
        bool parked = false;
        float d_keep = 0.f;
        if (valid) {
            o = make3(ray_o[i * 3], ray_o[i * 3 + 1], ray_o[i * 3 + 2]);
            d = make3(ray_d[i * 3], ray_d[i * 3 + 1], ray_d[i * 3 + 2]);
            nr = near_[i]; fr = far_[i];
            if (it == 0) { t = nr; st = fr; }
            else {
                t = s.t[i]; occ = s.occ[i]; d0 = s.d0[i]; cd = s.cd[i]; dt = s.dt[i]; st = s.st[i]; off = s.off[i]; rlx = s.rlx[i];
                int slot = s.q_slot[i];
                float smpl = s.q_smpl[i];
                float d1 = (slot >= 0) ? hdq_blend(q.net[slot], smpl, cfg.th, true) : smpl;
                int pi = it - 1;
                float tanv = 1.0f / cfg.tan_i;
                if (pi >= cfg.skip) {
                    float c = fmaxf(d1, 0.f) / fmaxf(fmaxf(t, nr), cfg.eps) / (tanv * 2.f);
                    if (c < occ) occ = c;
                }
                float d1u = fabsf(d1), d0u = fabsf(d0);
                if (signf(d0) != signf(d1)) {
                    st = t - dt * clampf(d1u / (d0u + d1u + cfg.eps), 0.f, 1.f);
                    off = 0.f; rlx = 0.f;
                }
                if (d1u < cd) { cd = d1u; st = t; }
                dt = d1 + rlx * d1 + off;
                const float t_old = t;
                t = fmaxf(fminf(t + dt, fr), nr);
                d0 = d1;
                parked = (t == t_old);
                d_keep = d1;
            }
        }
        if (it < cfg.iters) {
            HdqFront f; f.in_shell = false; f.smpl = 0.f;
            const bool ask = valid && !parked;
            hdq_front<false>(fc, sv, nverts, o + d * t, ask, cfg.th, cfg.blend_radius, f, 12, packets != 0);   // a warp = 32 neighbouring pixels of an image row: a packet
            bool ins = ask && f.in_shell;
            count_queries(cnt, ask, ins);
            int slot = warp_append(q.count, ins);
            if (ins) { q.bpts[slot * 3] = f.bpts.x; q.bpts[slot * 3 + 1] = f.bpts.y; q.bpts[slot * 3 + 2] = f.bpts.z; }
            if (valid) {
                s.t[i] = t; s.occ[i] = occ; s.d0[i] = d0; s.cd[i] = cd; s.dt[i] = dt; s.st[i] = st; s.off[i] = off; s.rlx[i] = rlx;
                s.q_smpl[i] = parked ? d_keep : f.smpl; s.q_slot[i] = parked ? -1 : slot;     // slot -1: q_smpl is the final distance
            }
        } else {
            float a = 1.f - occ;
            bool fg = valid && (a > 0.f);
            int slot = warp_append(cnt.n_fg, fg);
            if (valid) {
                float3 sp = o + d * st;
                surf[i * 3] = sp.x; surf[i * 3 + 1] = sp.y; surf[i * 3 + 2] = sp.z;
                acc[i] = a;
                depth[i] = (sp.x - o.x) / d.x;
                if (fg) fg_ray[slot] = i;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ shadows
struct ShadowRays {      // SoA over traced (pixel, light) pairs
    int* fg; unsigned short* light;
    float *near_, *far_, *t, *occ, *d0, *q_smpl;
    int* q_slot;
    int cap;             // entries the arrays hold (256 per ray of ra_config.max_rays)
    int* dropped;        // device counter: rays that did not fit (reported by ra_get_stats; their light stays 'visible')
    int* n_rays;         // device counter: rays traced, counted by the tracer's last launch (the list counter itself counts padded packet slots)
};
// append guard: a light layout with more than 256 front-facing lights per pixel (not the antipodally symmetric 16x32 grid of
// gen_light_xyz) could generate more shadow rays than the workspace holds -- drop and count instead of writing out of bounds
__device__ __forceinline__ bool shadow_slot_ok(const ShadowRays& sr, bool trace, int slot) {
    if (trace && slot >= sr.cap) { atomicAdd(sr.dropped, 1); return false; }
    return trace;
}

// one warp appends its packet: padded = a whole aligned block of 32 entries (fg = -1 on the lanes without a ray), else compacted
__device__ __forceinline__ void shadow_append(const ShadowRays& sr, int* n_shadow, bool padded, bool trace, int f, int l, float nr, float fr) {
    const unsigned any = __ballot_sync(0xffffffffu, trace);
    if (!any) return;
    if (padded) {
        int base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(n_shadow, 32);
        const int slot = __shfl_sync(0xffffffffu, base, 0) + (threadIdx.x & 31);
        if (slot >= sr.cap) { if (trace) atomicAdd(sr.dropped, 1); return; }
        sr.fg[slot] = trace ? f : -1; sr.light[slot] = (unsigned short)l; sr.near_[slot] = nr; sr.far_[slot] = fr;
    } else {
        const int slot = warp_append(n_shadow, trace);
        if (shadow_slot_ok(sr, trace, slot)) { sr.fg[slot] = f; sr.light[slot] = (unsigned short)l; sr.near_[slot] = nr; sr.far_[slot] = fr; }
    }
}

__device__ __forceinline__ void aabb_near_far(const float* bmin, const float* bmax, float3 o, float3 d, float& near_, float& far_) {
    // get_near_far_aabb(return_raw=True), net_utils.py:1683-1712
    float dd[3] = {d.x, d.y, d.z}, oo[3] = {o.x, o.y, o.z};
    near_ = -3.0e38f; far_ = 3.0e38f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float v = dd[a];
        if (v < 1e-8f && v > -1e-16f) v = 1e-8f;
        float t0 = (bmin[a] - oo[a]) / v, t1 = (bmax[a] - oo[a]) / v;
        near_ = fmaxf(near_, fminf(t0, t1));
        far_ = fminf(far_, fmaxf(t0, t1));
    }
}

// light_visibility set-up (:265-329): per (fg pixel, light): ldot, front-facing & box tests, ray append.
// packets = 0 (the default for the human pass): one warp = 32 lights of one foreground pixel, rays appended compacted.
// packets = 1: one warp = one PACKET, the same light for 32 consecutive foreground pixels, appended as a whole aligned block of the ray
// list (lanes without a ray carry fg = -1; compacted instead if the padded list could outgrow the workspace).  That is the floor
// pass's layout (ground.cuh), where it is worth 2.2x; on the body it measured SLOWER (visibility stage 17.9 -> 19.5 ms at 512^2): the
// normals of 32 neighbouring pixels sweep most of the hemisphere (a limb is ~30 pixels wide), so nearly every (tile, light) pair holds a
// ray and the padded list has 1.9x the entries, while rays that leave from ONE pixel share their 3-NN lists during the first iterations
// (the same order with a compacted list, packets straddling warps: 18.2 ms -- no better than the pixel-major 18.0).
__global__ void k_shadow_gen(const FrameConst* __restrict__ fc, const int* __restrict__ n_fg, const int* __restrict__ fg_ray,
                             const float* __restrict__ surf /*[P][3] by ray*/, const float* __restrict__ f_norm /*[fg][3]*/,
                             const float* __restrict__ ldir /*[L][3]*/, int L, float lv_near, float bbox_margin, int chunk_actual,
                             int lay_block, int lay_world, int lay_rank,     // tile sharding: local ray -> global ray (identity when world == 1)
                             int vis_mode,       // 0: DFSS tracing; 1: cfg.local_visibility (lvis = ldot > 0); 2: cfg.no_visibility (lvis = 1)   :296-301
                             float* lvis, float* ldot, ShadowRays sr, int* n_shadow, int packets /* 0: pixel-major compact list (one warp = 32 lights of a pixel) */) {
    const int lane = threadIdx.x & 31;
    const int nfg = *n_fg;
    const long long ntile = (nfg + 31) >> 5;
    const long long total = packets ? ntile * L * 32 : (long long)nfg * L;
    const bool padded = packets && total <= (long long)sr.cap;
    __shared__ int s_wcnt[32], s_base;
    const int wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (long long bbase = (long long)blockIdx.x * blockDim.x; bbase < total; bbase += (long long)gridDim.x * blockDim.x) {      // block-uniform trip count
        const long long base = bbase + threadIdx.x - lane;
        const long long w = base >> 5;
        const int l = packets ? (int)(w % L) : (int)((base + lane) % L);
        const int f = packets ? (int)(w / L) * 32 + lane : (int)((base + lane) / L);
        const bool valid = base + lane < total && (!packets || f < nfg);
        bool trace = false;
        float nr = 0.f, fr = 0.f;
        if (valid) {
            const long long idx = (long long)f * L + l;
            int ray = fg_ray[f];
            float3 n = make3(f_norm[f * 3], f_norm[f * 3 + 1], f_norm[f * 3 + 2]);
            float3 dl = make3(ldir[l * 3], ldir[l * 3 + 1], ldir[l * 3 + 2]);
            float dt = dl.x * n.x + dl.y * n.y + dl.z * n.z;
            ldot[idx] = dt;
            float vis = 0.f;
            if (vis_mode) vis = (vis_mode == 2 || dt > 0.f) ? 1.f : 0.f;
            else if (dt > 0.f) {
                // in-place wbounds growth per pixel chunk (:1020-1022), by the ray's index in the WHOLE frame
                const int gray = (lay_world > 1) ? ((ray / lay_block) * lay_world + lay_rank) * lay_block + ray % lay_block : ray;
                float pad = bbox_margin * (float)(1 + gray / chunk_actual);
                float bmin[3] = {fc->wb[0] - pad, fc->wb[1] - pad, fc->wb[2] - pad};
                float bmax[3] = {fc->wb[3] + pad, fc->wb[4] + pad, fc->wb[5] + pad};
                float3 o = make3(surf[ray * 3], surf[ray * 3 + 1], surf[ray * 3 + 2]);
                aabb_near_far(bmin, bmax, o, dl, nr, fr);
                nr = fmaxf(nr, lv_near); fr = fmaxf(fr, lv_near);
                trace = nr < fr;
                vis = 1.f;          // outside the box: visible; traced rays overwrite this at the end
            }
            lvis[idx] = vis;
        }
        if (packets) { shadow_append(sr, n_shadow, padded, trace, f, l, nr, fr); continue; }
        // compact list: the block's warps append together with ONE atomic (a warp-level append per 32 (pixel, light) pairs was 167 k
        // atomics on one counter per frame and bound this kernel: 160 us)
        const unsigned mask = __ballot_sync(0xffffffffu, trace);
        if (lane == 0) s_wcnt[wid] = __popc(mask);
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int k = 0; k < nwarp; k++) { const int c = s_wcnt[k]; s_wcnt[k] = tot; tot += c; }
            s_base = tot ? atomicAdd(n_shadow, tot) : 0;
        }
        __syncthreads();
        const int slot = s_base + s_wcnt[wid] + __popc(mask & ((1u << lane) - 1u));
        if (shadow_slot_ok(sr, trace, slot)) { sr.fg[slot] = f; sr.light[slot] = (unsigned short)l; sr.near_[slot] = nr; sr.far_[slot] = fr; }
        __syncthreads();          // s_wcnt / s_base are reused by the next iteration
    }
}

// Soft-shadow tracing step (A.4 soft + claybook), same finish/start structure as k_trace_surface.
__global__ void __launch_bounds__(256, RA_SHADOW_MINBLOCKS) k_trace_shadow(int it, TraceCfg cfg, const FrameConst* __restrict__ fc, SortedVerts sv, int nverts,
                               const int* __restrict__ n_shadow, const int* __restrict__ fg_ray, const float* __restrict__ surf,
                               const float* __restrict__ ldir, const float* __restrict__ lsharp, int L,
                               ShadowRays sr, QueryList q, Counters cnt, float* lvis, int part, int nparts, int packets, int final_skip) {
    // rays [lo, hi) of the list: the host runs the parts on different streams so that one part's CUDA-core work overlaps
    // the other part's tensor-core MLP kernel
    const long long Nall = min(*n_shadow, sr.cap);
    const int lo = (int)(Nall * part / nparts) & ~31, N = (part + 1 == nparts) ? (int)Nall : ((int)(Nall * (part + 1) / nparts) & ~31);   // parts cut between packets
    for (int base = lo + blockIdx.x * blockDim.x; base < N; base += gridDim.x * blockDim.x) {      // block-uniform trip count
        int i = base + threadIdx.x;
        bool valid = i < N;
        float3 o = make3(0, 0, 0), d = make3(0, 0, 1);
        float nr = 0.f, fr = 0.f, t = 0.f, occ = 1.f, d0 = 1e9f;
        int f = 0, l = 0;
        bool parked = false;        // front did not move (t clamped): same point, same distance -- reuse, no query (exact)
        float d_keep = 0.f;
        if (valid) { f = sr.fg[i]; valid = f >= 0; }      // padded packets: lanes without a ray
        // A ray whose state can no longer change is FINAL (bit 15 of its light index): fully occluded, or parked for the second iteration
        // in a row (the first parked iteration still sees a new d0, from the second on every input of the update repeats).  A final ray
        // costs two loads per launch instead of ten loads, five stores and the update -- most of the floor's 26 M rays in its later iterations.
        unsigned short lraw = 0;
        bool fin = false, was_parked = false;
        if (valid) { lraw = sr.light[i]; l = lraw & 0x7fff; fin = (lraw & 0x8000) != 0; }
        if (valid && !fin) {
            int ray = fg_ray ? fg_ray[f] : f;          // floor pass: the ray list indexes image pixels directly
            o = make3(surf[ray * 3], surf[ray * 3 + 1], surf[ray * 3 + 2]);
            d = make3(ldir[l * 3], ldir[l * 3 + 1], ldir[l * 3 + 2]);
            nr = sr.near_[i]; fr = sr.far_[i];
            if (it == 0) t = nr;
            else {
                t = sr.t[i]; occ = sr.occ[i]; d0 = sr.d0[i];
                // occ only ever decreases towards 0 (min of non-negative terms): a fully occluded ray is final.
                // Skipping its remaining queries changes nothing in the result (exact early termination).
                if (occ > 0.f) {
                    int slot = sr.q_slot[i];
                    was_parked = slot == -3;
                    float smpl = sr.q_smpl[i];
                    float d1 = (slot >= 0) ? hdq_blend(q.net[slot], smpl, cfg.th, true) : smpl;
                    int pi = it - 1;
                    float tanv = 1.0f / lsharp[l];
                    float off = cfg.offset, rlx = cfg.relax;
                    if (pi >= cfg.skip) {
                        float dx0 = d0 + rlx * d0 + off;
                        float dx1 = d1 + rlx * d1 + off;
                        float dy = (dx1 * dx1) / (2.f * dx0);
                        float dx = (sqrtf(dx1 * dx1 - dy * dy) - off) / (1.f + rlx);
                        float c = fmaxf(dx, 0.f) / fmaxf(fmaxf(t - dy, nr), cfg.eps) / (tanv * 2.f);
                        bool m = (c < occ) && (dy < t) && (dx1 > 0.f) && (dx0 > 0.f) && (dx > 0.f) && (dy > 0.f) && (dy < dx0);
                        if (m) occ = c;
                        float c2 = fmaxf(d1, 0.f) / fmaxf(fmaxf(t, nr), cfg.eps) / (tanv * 2.f);
                        if (c2 < occ) occ = c2;
                    }
                    float dt = d1 + rlx * d1 + off;
                    const float t_old = t;
                    t = fmaxf(fminf(t + dt, fr), nr);
                    d0 = d1;
                    parked = (t == t_old);
                    d_keep = d1;
                }
            }
        }
        const bool alive = valid && !fin && (occ > 0.f);
        if (it < cfg.iters) {
            HdqFront hf; hf.in_shell = false; hf.smpl = 0.f;
            const bool ask = alive && !parked;
            hdq_front<false>(fc, sv, nverts, o + d * t, ask, cfg.th, cfg.blend_radius, hf, 32, packets != 0);      // shadow rays: far lanes are served as one packet, stragglers by the whole warp
            bool ins = ask && hf.in_shell;
            count_queries(cnt, ask, ins);
            int slot = warp_append(q.count, ins);
            if (ins) { q.bpts[(size_t)slot * 3] = hf.bpts.x; q.bpts[(size_t)slot * 3 + 1] = hf.bpts.y; q.bpts[(size_t)slot * 3 + 2] = hf.bpts.z; }
            if (valid && !fin) {
                sr.t[i] = t; sr.occ[i] = occ; sr.d0[i] = d0;
                sr.q_smpl[i] = (alive && parked) ? d_keep : hf.smpl; sr.q_slot[i] = (alive && parked) ? -3 : slot;      // -3: parked, q_smpl is the final distance
                if (final_skip && (!alive || (parked && was_parked))) sr.light[i] = lraw | 0x8000;
            }
        } else {
            if (valid) lvis[(size_t)f * L + l] = fin ? sr.occ[i] : occ;
            // the rays of the list are counted here, once per block and launch (ra_stats.n_shadow_rays; the list counter itself counts
            // padded packet slots) -- a second same-address atomic per appended warp made k_shadow_gen twice as slow
            const int c = __syncthreads_count(valid);
            if (threadIdx.x == 0 && c) atomicAdd(sr.n_rays, c);
        }
    }
}

// ------------------------------------------------------------------------------------------ generic point queries
// x (n,3) -> front-end; in-shell points appended to the query list.  Used by ra_query_sdf.
__global__ void k_points_front(const FrameConst* __restrict__ fc, SortedVerts sv, int nverts, const float* __restrict__ x, int n,
                               float th, float blend_radius, float* smpl, int* slot_out, QueryList q, Counters cnt) {
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {      // block-uniform trip count
        int i = base + threadIdx.x;
        bool valid = i < n;
        HdqFront f; f.in_shell = false; f.smpl = 0.f;
        hdq_front<false>(fc, sv, nverts, valid ? make3(x[i * 3], x[i * 3 + 1], x[i * 3 + 2]) : make3(0, 0, 0), valid, th, blend_radius, f);
        bool ins = valid && f.in_shell;
        count_queries(cnt, valid, ins);
        int slot = warp_append(q.count, ins);
        if (ins) { q.bpts[(size_t)slot * 3] = f.bpts.x; q.bpts[(size_t)slot * 3 + 1] = f.bpts.y; q.bpts[(size_t)slot * 3 + 2] = f.bpts.z; }
        if (valid) { smpl[i] = f.smpl; slot_out[i] = slot; }
    }
}

// The exact 3-NN on its own (row a4: pytorch3d.ops.knn_points K=3 at sample_utils.py:122): world points -> pose space -> the three
// nearest posed vertices as ORIGINAL vertex indices, nearest first, with their squared distances.
__global__ void k_points_knn(const FrameConst* __restrict__ fc, SortedVerts sv, int nverts, const float* __restrict__ x, int n, int* ids, float* d2,
                             int packets /* 1: every 32 consecutive points are one packet (the shadow tracer's search) */) {
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {      // block-uniform trip count
        const int i = base + threadIdx.x;
        const bool valid = i < n;
        float3 p = make3(0, 0, 0);
        if (valid) {
            const float3 q = make3(x[i * 3] - fc->Th[0], x[i * 3 + 1] - fc->Th[1], x[i * 3 + 2] - fc->Th[2]);
            p = make3(q.x * fc->R[0] + q.y * fc->R[3] + q.z * fc->R[6], q.x * fc->R[1] + q.y * fc->R[4] + q.z * fc->R[7],
                      q.x * fc->R[2] + q.y * fc->R[5] + q.z * fc->R[8]);
        }
        KnnOut nn;
        knn3_warp(fc, sv, nverts, p, valid, nn, packets ? 32 : 12, packets != 0);
        if (valid)
            for (int k = 0; k < 3; k++) { ids[i * 3 + k] = __float_as_int(__ldg(&sv.pos[nn.id[k]]).w); d2[i * 3 + k] = nn.d2[k]; }
    }
}

__global__ void k_points_finish(const float* __restrict__ smpl, const int* __restrict__ slot, const float* __restrict__ net, int n,
                                float th, int smooth, float* sdf) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = slot[i];
        sdf[i] = (s >= 0) ? hdq_blend(net[s], smpl[i], th, smooth != 0) : smpl[i];
    }
}

// ------------------------------------------------------------------------------------------ attribute samples
// Work list of in-shell samples that go through forward + input-gradient (net(x, v, d, batch)).
struct AttrList {
    float* bpts;     // [cap][3]
    float* mats;     // [cap][18]: bigAr (9), Rinv (9)     -- normal transform base_network.py:471-475
    float* bvds;     // [cap][3] big-pose view dirs (AniSDF colour net)      base_network.py:324-334
    int* src;        // [cap] sample index the raw row scatters to
    int* count;
};

// Sample points: mode 0: explicit (x, v) arrays; mode 1: surface samples surf + z_m * view over the fg list;
// mode 2: uniform volume samples along rays (base_renderer.py:15-31).
__global__ void k_attr_front(int mode, const FrameConst* __restrict__ fc, SortedVerts sv, int nverts, float th, float blend_radius,
                             const float* __restrict__ x, const float* __restrict__ v, long long n_explicit,
                             const int* __restrict__ n_fg, const int* __restrict__ fg_ray, const float* __restrict__ surf,
                             const float* __restrict__ ray_o, const float* __restrict__ ray_d, const float* __restrict__ near_,
                             const float* __restrict__ far_, int n_samples, float sample_range, float clip_near, float clip_far,
                             long long ray0, long long n_rays,
                             AttrList al, Counters cnt, int packets) {
    long long total = (mode == 0) ? n_explicit : (mode == 1 ? (long long)(*n_fg) * n_samples : n_rays * n_samples);
    for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {   // block-uniform
        long long i = base + threadIdx.x;
        bool valid = i < total;
        float3 px = make3(0, 0, 0), pv = make3(0, 0, 1);
        if (valid) {
            if (mode == 0) {
                px = make3(x[i * 3], x[i * 3 + 1], x[i * 3 + 2]);
                if (v) pv = make3(v[i * 3], v[i * 3 + 1], v[i * 3 + 2]);
            } else if (mode == 1) {
                int f = (int)(i / n_samples), m = (int)(i % n_samples);
                int ray = fg_ray[f];
                pv = make3(ray_d[ray * 3], ray_d[ray * 3 + 1], ray_d[ray * 3 + 2]);
                // zval = linspace(0,1,S) * 2*range - range   (sphere_tracing_renderer.py:607-613)
                float z = (n_samples == 1) ? 0.5f : (float)m / (float)(n_samples - 1);
                z = z * (2.f * sample_range) - sample_range;
                px = make3(surf[ray * 3] + z * pv.x, surf[ray * 3 + 1] + z * pv.y, surf[ray * 3 + 2] + z * pv.z);
            } else {
                long long r = ray0 + i / n_samples; int m = (int)(i % n_samples);
                pv = make3(ray_d[r * 3], ray_d[r * 3 + 1], ray_d[r * 3 + 2]);
                float tv = (float)m / (float)(n_samples - 1);
                float nr = fmaxf(near_[r], clip_near), fr = fminf(far_[r], clip_far);
                float z = nr * (1.f - tv) + fr * tv;
                px = make3(ray_o[r * 3] + pv.x * z, ray_o[r * 3 + 1] + pv.y * z, ray_o[r * 3 + 2] + pv.z * z);
            }
        }
        HdqFront f; f.in_shell = false;
        hdq_front<true>(fc, sv, nverts, px, valid, th, blend_radius, f, 12, packets != 0);      // volume samples: a warp = 32 consecutive samples of one ray
        bool ins = valid && f.in_shell;
        count_queries(cnt, false, false);
        int slot = warp_append(al.count, ins);
        if (ins) {
            size_t s = (size_t)slot;
            al.bpts[s * 3] = f.bpts.x; al.bpts[s * 3 + 1] = f.bpts.y; al.bpts[s * 3 + 2] = f.bpts.z;
#pragma unroll
            for (int k = 0; k < 9; k++) { al.mats[s * 18 + k] = f.bigAr[k]; al.mats[s * 18 + 9 + k] = f.Rinv[k]; }
            // view dir: world -> pose (v @ R), pose -> tpose (A^T), tpose -> bigpose (bigRinv^T)
            float3 p1 = make3(pv.x * fc->R[0] + pv.y * fc->R[3] + pv.z * fc->R[6],
                              pv.x * fc->R[1] + pv.y * fc->R[4] + pv.z * fc->R[7],
                              pv.x * fc->R[2] + pv.y * fc->R[5] + pv.z * fc->R[8]);
            float3 p2 = make3(f.Ar[0] * p1.x + f.Ar[3] * p1.y + f.Ar[6] * p1.z,
                              f.Ar[1] * p1.x + f.Ar[4] * p1.y + f.Ar[7] * p1.z,
                              f.Ar[2] * p1.x + f.Ar[5] * p1.y + f.Ar[8] * p1.z);
            float3 p3 = make3(f.bigRinv[0] * p2.x + f.bigRinv[3] * p2.y + f.bigRinv[6] * p2.z,
                              f.bigRinv[1] * p2.x + f.bigRinv[4] * p2.y + f.bigRinv[7] * p2.z,
                              f.bigRinv[2] * p2.x + f.bigRinv[5] * p2.y + f.bigRinv[8] * p2.z);
            al.bvds[s * 3] = p3.x; al.bvds[s * 3 + 1] = p3.y; al.bvds[s * 3 + 2] = p3.z;
            al.src[s] = (int)(mode == 2 ? i : i);
        }
    }
}

// sdf -> occ over the fixed 5 mm interval   net_utils.py:867-893
__device__ __forceinline__ float sdf_to_occ(float sdf, float beta) {
    float x = -sdf;
    float sigma = (x <= 0.f) ? (1.f / beta * (0.5f * expf(x / beta))) : (1.f / beta * (1.f - 0.5f * expf(-x / beta)));
    return 1.f - expf(-fmaxf(sigma, 0.f) * 0.005f);
}

// world normal from d sdf / d bpts: normalize -> bigA^T -> Rinv^T -> R_world -> normalize   (base_network.py:471-475)
__device__ __forceinline__ float3 normal_to_world(float3 g, const float* mats, const FrameConst* fc) {
    float3 n = normalize_ref(g);
    const float* B = mats; const float* Ri = mats + 9;
    float3 a = make3(B[0] * n.x + B[3] * n.y + B[6] * n.z, B[1] * n.x + B[4] * n.y + B[7] * n.z, B[2] * n.x + B[5] * n.y + B[8] * n.z);
    float3 b = make3(Ri[0] * a.x + Ri[3] * a.y + Ri[6] * a.z, Ri[1] * a.x + Ri[4] * a.y + Ri[7] * a.z, Ri[2] * a.x + Ri[5] * a.y + Ri[8] * a.z);
    // pose_dirs_to_world_dirs: n @ R^T
    float3 c = make3(b.x * fc->R[0] + b.y * fc->R[1] + b.z * fc->R[2], b.x * fc->R[3] + b.y * fc->R[4] + b.z * fc->R[5],
                     b.x * fc->R[6] + b.y * fc->R[7] + b.z * fc->R[8]);
    return normalize_ref(c);
}

// normals for the whole attribute chunk (needed before the AniSDF colour net input is assembled)
__global__ void k_attr_normals(const FrameConst* __restrict__ fc, const float* __restrict__ gbp, const float* __restrict__ mats,
                               float* nrm, const int* count, int row0, int rows_cap) {
    int M = min(*count - row0, rows_cap);
    if (M <= 0) return;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        float3 n = normal_to_world(make3(gbp[m * 3], gbp[m * 3 + 1], gbp[m * 3 + 2]), mats + (size_t)m * 18, fc);
        nrm[m * 3] = n.x; nrm[m * 3 + 1] = n.y; nrm[m * 3 + 2] = n.z;
    }
}

// AniSDF colour net input: [PE4(bvds) (27) | norm (3) | feat (256) | pad] = 288 wide   (base_network.py:160-161)
__global__ void k_render_input(const float* __restrict__ bvds, const float* __restrict__ nrm, const float* __restrict__ feat,
                               int ldo, float* X, int ldx, const int* count, int row0, int rows_cap) {
    int M = min(*count - row0, rows_cap);
    if (M <= 0) return;
    // one thread per (row, column): coalesced copies of the 256 feature columns (a thread per row moved 288 scattered words)
    const unsigned n_el = (unsigned)M * 288u;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
        const unsigned m = i / 288u, k = i % 288u;
        float v = 0.f;
        if (k < 27u) v = pe_feature_ref(bvds + (size_t)m * 3, 4, (int)k);
        else if (k < 30u) v = nrm[(size_t)m * 3 + (k - 27u)];
        else if (k < 286u) v = feat[(size_t)m * ldo + (k - 30u)];
        X[(size_t)m * ldx + k] = v;
    }
}

// Assemble the raw rows (relight 17 ch: cpts,bpts,resd,albedo,rough,norm,occ; AniSDF 16 ch: cpts,bpts,resd,norm,rgb,occ)
// and scatter them to their sample slots   (relight_network.py:97-104, base_network.py:504-510)
__global__ void k_attr_finish(int relight, const float* __restrict__ bpts, const float* __restrict__ cpts, const float* __restrict__ resd,
                              const float* __restrict__ out257, int ldo, const float* __restrict__ nrm,
                              const float* __restrict__ head_a /*[M][4] albedo z | rgb z*/, const float* __restrict__ head_r /*[M][4]*/,
                              float beta, float albedo_slope, float albedo_bias, float rough_slope, float rough_bias,
                              const int* __restrict__ src, float* raw, const int* count, int row0, int rows_cap) {
    int M = min(*count - row0, rows_cap);
    if (M <= 0) return;
    int C = relight ? 17 : 16;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        float* r = raw + (size_t)src[m] * C;
        for (int c = 0; c < 3; c++) { r[c] = cpts[m * 3 + c]; r[3 + c] = bpts[m * 3 + c]; r[6 + c] = resd[m * 3 + c]; }
        float occ = sdf_to_occ(out257[(size_t)m * ldo], beta);
        if (relight) {
            for (int c = 0; c < 3; c++) r[9 + c] = albedo_slope * (1.f / (1.f + expf(-head_a[m * 4 + c]))) + albedo_bias;
            r[12] = rough_slope * (1.f / (1.f + expf(-head_r[m * 4]))) + rough_bias;
            r[13] = nrm[m * 3]; r[14] = nrm[m * 3 + 1]; r[15] = nrm[m * 3 + 2];
            r[16] = occ;
        } else {
            r[9] = nrm[m * 3]; r[10] = nrm[m * 3 + 1]; r[11] = nrm[m * 3 + 2];
            for (int c = 0; c < 3; c++) r[12 + c] = 1.f / (1.f + expf(-head_a[m * 4 + c]));
            r[15] = occ;
        }
    }
}

// 3-sample blend at the surface (sphere_tracing_renderer.py:616-650): per fg pixel.
// Writes the compact per-fg attributes and scatters the acc-premultiplied maps to the P-ray outputs.
struct FgMaps { float *norm, *albedo, *rough; };   // compact [fg] arrays used by shadows / shading

struct OutMaps {
    float *rgb, *acc, *depth, *surf, *norm, *cpts, *bpts, *resd, *albedo, *rough, *shade;
};

__global__ void k_surface_blend(int relight, const int* __restrict__ n_fg, const int* __restrict__ fg_ray, const float* __restrict__ raw,
                                int n_samples, const float* __restrict__ acc_ray, const float* __restrict__ surf_ray,
                                const float* __restrict__ depth_ray, float albedo_slope, float albedo_bias, float rough_slope,
                                float rough_bias, float albedo_mult, FgMaps fm, OutMaps om) {
    int C = relight ? 17 : 16;
    int n = *n_fg;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        int ray = fg_ray[f];
        float val[16];
        for (int c = 0; c < C - 1; c++) val[c] = 0.f;
        float T = 1.f, wsum = 0.f;
        for (int m = 0; m < n_samples; m++) {
            const float* r = raw + ((size_t)f * n_samples + m) * C;
            float a = r[C - 1];
            float w = a * T;                      // w_i = a_i * prod_{j<i}(1 - a_j + 1e-8)   net_utils.py:987-991
            T *= (1.f - a + 1e-8f);
            wsum += w;
            for (int c = 0; c < C - 1; c++) val[c] += w * r[c];
        }
        float inv = wsum + 1e-8f;
        for (int c = 0; c < C - 1; c++) val[c] /= inv;
        float a = acc_ray[ray];
        float3 nrm = relight ? make3(val[13], val[14], val[15]) : make3(val[9], val[10], val[11]);
        if (nrm.x + nrm.y + nrm.z == 0.f) nrm = make3(1.f, 1.f, 1.f);
        nrm = normalize_ref(nrm);
        fm.norm[f * 3] = nrm.x; fm.norm[f * 3 + 1] = nrm.y; fm.norm[f * 3 + 2] = nrm.z;
        if (om.norm) { om.norm[ray * 3] = nrm.x * a; om.norm[ray * 3 + 1] = nrm.y * a; om.norm[ray * 3 + 2] = nrm.z * a; }
        for (int c = 0; c < 3; c++) {
            if (om.cpts) om.cpts[ray * 3 + c] = val[c] * a;
            if (om.bpts) om.bpts[ray * 3 + c] = val[3 + c] * a;
            if (om.resd) om.resd[ray * 3 + c] = val[6 + c];         // resd_map is not in blend_keys: not premultiplied
            if (om.surf) om.surf[ray * 3 + c] = surf_ray[ray * 3 + c] * a;
        }
        if (om.acc) om.acc[ray] = a;
        if (om.depth) om.depth[ray] = depth_ray[ray] * a;
        if (relight) {
            for (int c = 0; c < 3; c++) {
                float al = clampf(val[9 + c], albedo_bias, albedo_bias + albedo_slope) * albedo_mult;
                fm.albedo[f * 3 + c] = al;
                if (om.albedo) om.albedo[ray * 3 + c] = al * a;
            }
            float ro = clampf(val[12], rough_bias, rough_bias + rough_slope);
            fm.rough[f] = ro;
            if (om.rough) om.rough[ray] = ro * a;
        } else {
            for (int c = 0; c < 3; c++)
                if (om.rgb) om.rgb[ray * 3 + c] = val[12 + c] * a;
        }
    }
}

// ------------------------------------------------------------------------------------------ shading
// safe_divide (relight_utils.py:618-633): clamp |a|,|b| >= 1e-8 (0 -> +1e-8), divide, NaN/Inf -> 0, clip +-1e10
__device__ __forceinline__ float sd_clamp(float a) {
    if (a < 1e-8f && a >= 0.f) return 1e-8f;
    if (a > -1e-8f && a <= 0.f) return -1e-8f;
    return a;
}
__device__ __forceinline__ float sd_div(float a, float b) {
    float d = a / b;
    if (d != d) d = 0.f;
    if (isinf(d)) d = 0.f;
    return clampf(d, -1e10f, 1e10f);
}

// bilinear env-map fetch, grid_sample(align_corners=False, padding_mode='border')   relight_utils.py:106-127
// split into the probe-independent tap computation and the per-probe gather
struct EnvTap { int o00, o01, o10, o11; float w00, w01, w10, w11; };
__device__ __forceinline__ EnvTap envmap_tap(int H, int W, float3 d) {
    const float PI = 3.14159265358979323846f;
    float theta = acosf(d.z) - 1e-6f;
    float phi = atan2f(d.y, d.x);
    float qy = (theta / PI) * 2.f - 1.f;
    float qx = -phi / PI;
    float ix = ((qx + 1.f) * W - 1.f) / 2.f, iy = ((qy + 1.f) * H - 1.f) / 2.f;
    ix = clampf(ix, 0.f, (float)(W - 1)); iy = clampf(iy, 0.f, (float)(H - 1));
    int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
    float fx = ix - x0, fy = iy - y0;
    int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    EnvTap t;
    t.w00 = (1.f - fx) * (1.f - fy); t.w01 = fx * (1.f - fy); t.w10 = (1.f - fx) * fy; t.w11 = fx * fy;
    if (x0 + 1 > W - 1) { t.w01 = 0.f; t.w11 = 0.f; }
    if (y0 + 1 > H - 1) { t.w10 = 0.f; t.w11 = 0.f; }
    t.o00 = (y0 * W + x0) * 3; t.o01 = (y0 * W + x1) * 3; t.o10 = (y1 * W + x0) * 3; t.o11 = (y1 * W + x1) * 3;
    return t;
}
__device__ __forceinline__ float3 envmap_gather(const float* __restrict__ img, const EnvTap& t) {
    const float* p00 = img + t.o00; const float* p01 = img + t.o01; const float* p10 = img + t.o10; const float* p11 = img + t.o11;
    return make3(p00[0] * t.w00 + p01[0] * t.w01 + p10[0] * t.w10 + p11[0] * t.w11,
                 p00[1] * t.w00 + p01[1] * t.w01 + p10[1] * t.w10 + p11[1] * t.w11,
                 p00[2] * t.w00 + p01[2] * t.w01 + p10[2] * t.w10 + p11[2] * t.w11);
}
__device__ __forceinline__ float3 envmap_fetch(const float* __restrict__ img, int H, int W, float3 d) {
    EnvTap t = envmap_tap(H, W, d);
    return envmap_gather(img, t);
}

// Microfacet (cancel_cosine=True): returns the scalar specular term and the lambert cosine factor l.n / pi
// brdf_mode: 0 glossy + lambert, 1 cfg.lambert_only, 2 cfg.glossy_only (relight_utils.py:563-568)
__device__ __forceinline__ void brdf_mode_apply(int brdf_mode, float& spec, float& lambert_k) {
    if (brdf_mode == 1) spec = 0.f;
    if (brdf_mode == 2) lambert_k = 0.f;
}
__device__ __forceinline__ void microfacet_eval(float3 s2l, float3 s2c, float3 nrm, float rough, float f0, float& spec, float& lambert_k) {
    const float PI = 3.14159265358979323846f;
    float3 l = normalize_f(s2l), v = normalize_f(s2c), n = normalize_f(nrm);
    float l_dot_n = clampf(dot3(l, n), 1e-4f, 1.f);
    float v_dot_n = clampf(dot3(v, n), 1e-4f, 1.f);
    lambert_k = l_dot_n / PI;
    float3 h = normalize_f(l + v);
    float omc = 1.f - dot3(l, h);
    float f = f0 + (1.f - f0) * (omc * omc * omc * omc * omc);
    float alpha = rough * rough;
    float a2 = alpha * alpha;
    // D
    float cm = dot3(h, n);
    float chi = cm > 0.f ? 1.f : 0.f;
    float cm2 = cm * cm;
    float num = sd_clamp(1.f - cm2);
    cm2 = sd_clamp(cm2);
    float tm2 = sd_div(num, cm2);
    float den = PI * (cm2 * cm2) * ((a2 + tm2) * (a2 + tm2));
    float D = sd_div(sd_clamp(a2 * chi), sd_clamp(den));
    // G
    float cv = sd_clamp(dot3(n, v));
    float ct = sd_clamp(dot3(h, v));
    float chig = sd_div(ct, cv) > 0.f ? 1.f : 0.f;
    float cv2 = clampf(cv * cv, 0.f, 1.f);
    float numv = sd_clamp(1.f - cv2);
    cv2 = sd_clamp(cv2);
    float tv2 = clampf(sd_div(numv, cv2), 0.f, 1e10f);
    float G = sd_div(sd_clamp(chig * 2.f), sd_clamp(1.f + sqrtf(1.f + a2 * tv2)));
    spec = sd_div(sd_clamp(f * G * D), sd_clamp(4.f * 1.f * fabsf(v_dot_n)));
}

__device__ __forceinline__ float linear2srgb(float x) {
    x = clampf(x, 0.f, 1.f);
    return (x <= 0.0031308f) ? x * 12.92f : 1.055f * powf(x + 1e-7f, 1.f / 2.4f) - 0.055f;
}

// cfg.tonemapping_rendering (config.py:417; switched off for .exr / .hdr output, config.py:446-448)
__device__ __forceinline__ float tone(float x, int tonemap) { return tonemap ? linear2srgb(x) : x; }

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per fg pixel: 512-direction light sum.  premul != 0: inputs are multiplied by acc first, as the novel-light
// pass consumes the alpha_output_'d maps (novel_light_sphere_tracing.py:21-66).
// Outputs are written per RAY (scatter); rgb/shade premultiplied by acc when `out_premul`.
__global__ void k_shade(const int* __restrict__ n_fg, const int* __restrict__ fg_ray, const float* __restrict__ ray_o,
                        const float* __restrict__ surf_ray, const float* __restrict__ acc_ray, FgMaps fm,
                        const float* __restrict__ lvis, const float* __restrict__ ldot, const float* __restrict__ lxyz,
                        const float* __restrict__ larea, int L, const float* __restrict__ probe, int eh, int ew, float f0,
                        float shading_albedo, int premul, int out_premul, int tonemap, int brdf_mode, float* rgb, float* shade, float* spec) {
    const float PI = 3.14159265358979323846f;
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    int n = *n_fg;
    for (int f = warp; f < n; f += nwarps) {
        int ray = fg_ray[f];
        float a = acc_ray[ray];
        float pm = premul ? a : 1.f;
        float3 sp = make3(surf_ray[ray * 3] * pm, surf_ray[ray * 3 + 1] * pm, surf_ray[ray * 3 + 2] * pm);
        float3 ro = make3(ray_o[ray * 3], ray_o[ray * 3 + 1], ray_o[ray * 3 + 2]);
        float3 nr = make3(fm.norm[f * 3] * pm, fm.norm[f * 3 + 1] * pm, fm.norm[f * 3 + 2] * pm);
        float al[3] = {fm.albedo[f * 3] * pm, fm.albedo[f * 3 + 1] * pm, fm.albedo[f * 3 + 2] * pm};
        float rough = fm.rough[f] * pm;
        float3 s2c = normalize_ref(ro - sp);
        float cr[3] = {0, 0, 0}, cs[3] = {0, 0, 0}, cp[3] = {0, 0, 0};
        for (int l = lane; l < L; l += 32) {
            float3 s2l = normalize_ref(make3(lxyz[l * 3] - sp.x, lxyz[l * 3 + 1] - sp.y, lxyz[l * 3 + 2] - sp.z));
            float3 li = envmap_fetch(probe, eh, ew, s2l);
            float sp_term, lk;
            microfacet_eval(s2l, s2c, nr, rough, f0, sp_term, lk);
            brdf_mode_apply(brdf_mode, sp_term, lk);
            float lv = lvis[(size_t)f * L + l] * pm, ld = ldot[(size_t)f * L + l] * pm;
            float ar = larea[l];
            float lic[3] = {li.x, li.y, li.z};
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float brdf = sp_term + al[c] * lk;          // spec + albedo/pi * l.n
                float sh = lv * 1.f * ar * lic[c];          // lvis * ldot(:=1) * area * light
                cr[c] += brdf * sh;
                cs[c] += lv * ld * ar * lic[c];
                cp[c] += sp_term * (1.f * ar * lic[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) { cr[c] = warp_sum(cr[c]); cs[c] = warp_sum(cs[c]); cp[c] = warp_sum(cp[c]); }
        if (lane == 0) {
            float om = out_premul ? a : 1.f;
            for (int c = 0; c < 3; c++) {
                if (rgb) rgb[ray * 3 + c] = tone(cr[c], tonemap) * om;
                if (shade) shade[ray * 3 + c] = cs[c] * shading_albedo / PI * om;
                if (spec) spec[ray * 3 + c] = cp[c] * om;       // cfg.vis_specular_map (:739-748): spec brdf x unshadowed light, in blend_keys
            }
        }
    }
}

// ---- env-map rotation sweep -------------------------------------------------------------------------------
// rotate_envmap's shift_image on the probe (relight_utils.py:55-103): bilinear shift along the longitude axis by
// `shift` = eW / (eW*repeat) * j texels, wrap-around of the sample position, border clamp of the interpolation.
// out[r][y][x][c] for r in [0, n_rot): j = j0 + r.
__global__ void k_shift_probe(const float* __restrict__ probe, int H, int W, int repeat, int j0, int n_rot, float* out) {
    int total = n_rot * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int x = i % W, y = (i / W) % H, r = i / (W * H);
        float shift = (float)W / (float)(W * repeat) * (float)(j0 + r);
        float gx = fmodf((float)x + 0.5f + shift, (float)W);
        float nx = gx / (float)W * 2.f - 1.f;                       // the reference normalises, grid_sample un-normalises
        float ix = ((nx + 1.f) * W - 1.f) / 2.f;
        ix = clampf(ix, 0.f, (float)(W - 1));
        int x0 = (int)floorf(ix);
        float fx = ix - x0;
        int x1 = min(x0 + 1, W - 1);
        float w1 = (x0 + 1 > W - 1) ? 0.f : fx;
        const float* p0 = probe + (y * W + x0) * 3; const float* p1 = probe + (y * W + x1) * 3;
        for (int c = 0; c < 3; c++) out[(size_t)i * 3 + c] = p0[c] * (1.f - fx) + p1[c] * w1;
    }
}

// Light sum for up to 4 probes of equal size at once: geometry, BRDF and visibility terms do not depend on the probe,
// so a sweep over many env-maps (rotations) pays for them once per 4 probes and reads the (S_fg, L) visibility once.
#define RA_SHADE_MULTI 4
__global__ void k_shade_multi(const int* __restrict__ n_fg, const int* __restrict__ fg_ray, const float* __restrict__ ray_o,
                              const float* __restrict__ surf_ray, const float* __restrict__ acc_ray, FgMaps fm,
                              const float* __restrict__ lvis, const float* __restrict__ ldot, const float* __restrict__ lxyz,
                              const float* __restrict__ larea, int L, const float* __restrict__ probes, int n_probe, int eh, int ew,
                              float f0, float shading_albedo, float* rgb, float* shade, float* spec, long long P, int premul, int out_premul, int tonemap,
                              int brdf_mode) {
    const float PI = 3.14159265358979323846f;
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    int n = *n_fg;
    const int psz = eh * ew * 3;
    for (int f = warp; f < n; f += nwarps) {
        int ray = fg_ray[f];
        const float acc_f = acc_ray[ray];
        const float a = premul ? acc_f : 1.f;      // premultiplied inputs (a19) unless ground shading is on
        const float om = out_premul ? acc_f : 1.f;
        float3 sp = make3(surf_ray[ray * 3] * a, surf_ray[ray * 3 + 1] * a, surf_ray[ray * 3 + 2] * a);
        float3 ro = make3(ray_o[ray * 3], ray_o[ray * 3 + 1], ray_o[ray * 3 + 2]);
        float3 nr = make3(fm.norm[f * 3] * a, fm.norm[f * 3 + 1] * a, fm.norm[f * 3 + 2] * a);
        float al[3] = {fm.albedo[f * 3] * a, fm.albedo[f * 3 + 1] * a, fm.albedo[f * 3 + 2] * a};
        float rough = fm.rough[f] * a;
        float3 s2c = normalize_ref(ro - sp);
        float cr[RA_SHADE_MULTI][3], cs[RA_SHADE_MULTI][3], cp[RA_SHADE_MULTI][3];
#pragma unroll
        for (int e = 0; e < RA_SHADE_MULTI; e++)
            for (int c = 0; c < 3; c++) { cr[e][c] = 0.f; cs[e][c] = 0.f; cp[e][c] = 0.f; }
        for (int l = lane; l < L; l += 32) {
            float3 s2l = normalize_ref(make3(lxyz[l * 3] - sp.x, lxyz[l * 3 + 1] - sp.y, lxyz[l * 3 + 2] - sp.z));
            float sp_term, lk;
            microfacet_eval(s2l, s2c, nr, rough, f0, sp_term, lk);
            brdf_mode_apply(brdf_mode, sp_term, lk);
            float lv = lvis[(size_t)f * L + l] * a, ld = ldot[(size_t)f * L + l] * a;
            float ar = larea[l];
            const EnvTap tap = envmap_tap(eh, ew, s2l);
#pragma unroll
            for (int e = 0; e < RA_SHADE_MULTI; e++) {
                if (e < n_probe) {
                    float3 li = envmap_gather(probes + (size_t)e * psz, tap);
                    float lic[3] = {li.x, li.y, li.z};
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float brdf = sp_term + al[c] * lk;
                        float sh = lv * 1.f * ar * lic[c];
                        cr[e][c] += brdf * sh;
                        cs[e][c] += lv * ld * ar * lic[c];
                        cp[e][c] += sp_term * (1.f * ar * lic[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int e = 0; e < RA_SHADE_MULTI; e++) {
            if (e >= n_probe) break;
            for (int c = 0; c < 3; c++) { cr[e][c] = warp_sum(cr[e][c]); cs[e][c] = warp_sum(cs[e][c]); cp[e][c] = warp_sum(cp[e][c]); }
            if (lane == 0)
                for (int c = 0; c < 3; c++) {
                    size_t o = ((size_t)e * P + ray) * 3 + c;
                    if (rgb) rgb[o] = tone(cr[e][c], tonemap) * om;
                    if (shade) shade[o] = cs[e][c] * shading_albedo / PI * om;
                    if (spec) spec[o] = cp[e][c] * om;
                }
        }
    }
}

// background pixels in the novel-light pass: all inputs are zero, spec is a probe-dependent constant
__global__ void k_bg_spec(const float* __restrict__ lxyz, const float* __restrict__ larea, int L, const float* __restrict__ probe,
                          int eh, int ew, float f0, int brdf_mode, float* out3) {
    int lane = threadIdx.x & 31;
    float cp[3] = {0, 0, 0};
    float3 z = make3(0, 0, 0);
    float3 s2c = normalize_ref(z);
    for (int l = lane; l < L; l += 32) {
        float3 s2l = normalize_ref(make3(lxyz[l * 3], lxyz[l * 3 + 1], lxyz[l * 3 + 2]));
        float3 li = envmap_fetch(probe, eh, ew, s2l);
        float sp_term, lk;
        microfacet_eval(s2l, s2c, z, 0.f, f0, sp_term, lk);
        brdf_mode_apply(brdf_mode, sp_term, lk);
        cp[0] += sp_term * (larea[l] * li.x); cp[1] += sp_term * (larea[l] * li.y); cp[2] += sp_term * (larea[l] * li.z);
    }
    for (int c = 0; c < 3; c++) cp[c] = warp_sum(cp[c]);
    if (lane == 0) { out3[0] = cp[0]; out3[1] = cp[1]; out3[2] = cp[2]; }
}

__global__ void k_fill3(float* dst, const float* __restrict__ v3, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * 3; i += (long long)gridDim.x * blockDim.x) dst[i] = v3[i % 3];
}

// (fg, L) visibility / cosine maps -> (P, L) ray layout, premultiplied by acc like alpha_output_
__global__ void k_scatter_lmaps(const int* __restrict__ n_fg, const int* __restrict__ fg_ray, const float* __restrict__ acc_ray,
                                const float* __restrict__ lvis, const float* __restrict__ ldot, int L, float* lvis_map, float* ldot_map) {
    long long total = (long long)(*n_fg) * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int f = (int)(i / L), l = (int)(i % L);
        int ray = fg_ray[f];
        float a = acc_ray[ray];
        if (lvis_map) lvis_map[(size_t)ray * L + l] = lvis[i] * a;
        if (ldot_map) ldot_map[(size_t)ray * L + l] = ldot[i] * a;
    }
}

// volume rendering over n_samples raw rows per ray (base_renderer.py:72-113; net_utils.py:970-999)
// Half a warp per ray, one lane per raw channel: a sample's 16 channels are one 64-byte segment, the transmittance product runs in
// sample order in every lane (the same sequence of roundings as a single thread walking the ray).  C == 16.
__global__ void k_volume_blend(const float* __restrict__ raw, int C, int n_samples, const float* __restrict__ near_, const float* __restrict__ far_,
                               float clip_near, float clip_far, long long ray0, long long n_rays, OutMaps om) {
    const int lane = threadIdx.x & 31, c = lane & 15;
    const long long hw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4, n_hw = ((long long)gridDim.x * blockDim.x) >> 4;
    const long long n_pad = (n_rays + 1) & ~1LL;          // both halves of a warp iterate together (shuffles below)
    for (long long r = hw; r < n_pad; r += n_hw) {
        const bool ok = r < n_rays;
        const long long ray = ray0 + (ok ? r : 0);
        float val = 0.f, T = 1.f, wsum = 0.f, dep = 0.f;
        const float nr = fmaxf(near_[ray], clip_near), fr = fminf(far_[ray], clip_far);
        const float* rr = raw + (size_t)(ok ? r : 0) * n_samples * C + c;
#pragma unroll 8
        for (int m = 0; m < n_samples; m++) {
            const float x = rr[(size_t)m * C];
            const float a = __shfl_sync(0xffffffffu, x, (lane & 16) | 15);      // the alpha channel of this half's ray
            const float w = a * T;
            T *= (1.f - a + 1e-8f);
            wsum += w;
            const float tv = (float)m / (float)(n_samples - 1);
            dep += w * (nr * (1.f - tv) + fr * tv);
            val += w * x;
        }
        if (!ok) continue;
        float* dst = c < 3 ? om.cpts : (c < 6 ? om.bpts : (c < 9 ? om.resd : (c < 12 ? om.norm : (c < 15 ? om.rgb : nullptr))));
        if (dst) dst[ray * 3 + c % 3] = val;
        if (c == 15) {
            if (om.acc) om.acc[ray] = wsum;
            if (om.depth) om.depth[ray] = dep;
        }
    }
}

// ---- image assembly: ray -> image scatter with alpha channel and 8-bit quantisation ----------------------
// Visualizer.generate_image (lib/visualizers/base_visualizer.py:182-202): img = bg_brightness; img[mask_at_box] = rgb_map;
// alpha[mask_at_box] = acc_map; RGBA.  Rays are the masked pixels in row-major order, so the ray of a pixel is its rank
// among the masked pixels: per-block counts -> single-block scan -> in-block ballot rank.
__global__ void k_mask_count(const unsigned char* __restrict__ mask, int n, int* blk_cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int c = __syncthreads_count(i < n && mask[i] != 0);
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = c;
}
__global__ void k_scan_blocks(int* blk_cnt, int nb) {          // one block: exclusive scan in place
    __shared__ int carry;
    __shared__ int wsum[32];
    int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += blockDim.x) {
        int i = base + tid;
        int v = (i < nb) ? blk_cnt[i] : 0, x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            wsum[lane] = s;
        }
        __syncthreads();
        int excl = carry + (wid ? wsum[wid - 1] : 0) + x - v;
        if (i < nb) blk_cnt[i] = excl;
        __syncthreads();
        if (tid == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
}
__global__ void k_assemble(const unsigned char* __restrict__ mask, int n, const int* __restrict__ blk_off, const float* __restrict__ rgb,
                           const float* __restrict__ acc, float bg, float* out_f, unsigned char* out_u8) {
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
    float v[4] = {bg, bg, bg, 0.f};
    if (m) { v[0] = rgb[ray * 3]; v[1] = rgb[ray * 3 + 1]; v[2] = rgb[ray * 3 + 2]; v[3] = acc[ray]; }
    for (int c = 0; c < 4; c++) {
        if (out_f) out_f[(size_t)i * 4 + c] = v[c];
        if (out_u8) out_u8[(size_t)i * 4 + c] = (unsigned char)(clampf(v[c], 0.f, 1.f) * 255.f);     // (img.clip(0,1) * 255).astype(uint8)
    }
}

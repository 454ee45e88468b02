// Per-frame batch preparation on the GPU: what pose_dataset.__getitem__ computes on the CPU per frame.
//   pose  : get_lbs_params / get_blend (base_dataset.py:308-397): Rodrigues + kinematic chain -> A (net_utils.py:1163-1172,
//           i.e. smplx.lbs.batch_rodrigues / batch_rigid_transform), LBS of the rest vertices (blend_utils.py:212-218,303-313),
//           pose -> world (blend_utils.py:264-273), vertex normals (pytorch3d Meshes.verts_normals_packed, or rotated rest
//           normals when no faces are given), get_bounds (data_utils.py:1241-1248)
//   rays  : get_rays_within_bounds (data_utils.py:925-938) = get_rays (:827-845) + get_near_far (:848-875) + compaction
// The per-frame input shrinks from ~3.9 MB (posed vertices, normals, 6890 x 52 weights, rays) to < 1 KB (poses, Rh, Th, K, R, T).
#pragma once
#include "common.cuh"

struct BodyDev {
    float* tjoints = nullptr; int* parents = nullptr; float* rverts = nullptr; float* rnorm = nullptr; int* faces = nullptr;
    float* weights = nullptr; int n_faces = 0; bool ready = false;
};

__device__ __forceinline__ void rodrigues_smplx(const float* r, float* R) {
    // smplx.lbs.batch_rodrigues: angle = |r + 1e-8|, axis = r / angle, R = I + sin K + (1 - cos) K K
    float ax = r[0] + 1e-8f, ay = r[1] + 1e-8f, az = r[2] + 1e-8f;
    float angle = sqrtf(ax * ax + ay * ay + az * az);
    float x = r[0] / angle, y = r[1] / angle, z = r[2] / angle;
    float s = sinf(angle), c = 1.f - cosf(angle);
    // K = [[0,-z,y],[z,0,-x],[-y,x,0]]
    float K[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
    float KK[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) KK[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    for (int i = 0; i < 9; i++) R[i] = ((i % 4) == 0 ? 1.f : 0.f) + s * K[i] + c * KK[i];
}

// one block of >= J threads: per-joint Rodrigues in parallel, the (short, sequential) kinematic chain on thread 0.
// A[j] = G[j] with the rest joint removed: A[:3,3] = G[:3,3] - G[:3,:3] tjoint_j   (batch_rigid_transform's rel_transforms)
__global__ void k_prep_pose(const float* __restrict__ poses, const float* __restrict__ Rh, const float* __restrict__ tjoints,
                            const int* __restrict__ parents, int J, float* A, float* posed_joints, float* Rout) {
    extern __shared__ float sh[];          // J * 9 rotations, J * 12 chain transforms
    float* Rj = sh;
    float* G = sh + J * 9;
    int j = threadIdx.x;
    if (j < J) rodrigues_smplx(poses + j * 3, Rj + j * 9);
    if (j == J) rodrigues_smplx(Rh, Rout);             // cv2.Rodrigues(Rh)
    __syncthreads();
    if (j == 0) {
        for (int k = 0; k < J; k++) {
            int p = parents[k];
            float rel[3];
            for (int a = 0; a < 3; a++) rel[a] = tjoints[k * 3 + a] - (p >= 0 ? tjoints[p * 3 + a] : 0.f);
            float* g = G + k * 12;
            if (p < 0) {
                for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) g[a * 4 + b] = Rj[k * 9 + a * 3 + b]; g[a * 4 + 3] = rel[a]; }
            } else {
                const float* gp = G + p * 12;
                for (int a = 0; a < 3; a++) {
                    for (int b = 0; b < 3; b++) g[a * 4 + b] = gp[a * 4] * Rj[k * 9 + b] + gp[a * 4 + 1] * Rj[k * 9 + 3 + b] + gp[a * 4 + 2] * Rj[k * 9 + 6 + b];
                    g[a * 4 + 3] = gp[a * 4] * rel[0] + gp[a * 4 + 1] * rel[1] + gp[a * 4 + 2] * rel[2] + gp[a * 4 + 3];
                }
            }
        }
    }
    __syncthreads();
    if (j < J) {
        const float* g = G + j * 12;
        for (int a = 0; a < 3; a++) {
            for (int b = 0; b < 3; b++) A[j * 16 + a * 4 + b] = g[a * 4 + b];
            A[j * 16 + a * 4 + 3] = g[a * 4 + 3] - (g[a * 4] * tjoints[j * 3] + g[a * 4 + 1] * tjoints[j * 3 + 1] + g[a * 4 + 2] * tjoints[j * 3 + 2]);
            if (posed_joints) posed_joints[j * 3 + a] = g[a * 4 + 3];
        }
        A[j * 16 + 12] = 0.f; A[j * 16 + 13] = 0.f; A[j * 16 + 14] = 0.f; A[j * 16 + 15] = 1.f;
    }
}

__device__ __forceinline__ void atomic_min_f(int* addr, float v) {      // order-preserving int encoding of floats
    int i = __float_as_int(v);
    i = (i >= 0) ? i : i ^ 0x7fffffff;
    atomicMin(addr, i);
}
__device__ __forceinline__ void atomic_max_f(int* addr, float v) {
    int i = __float_as_int(v);
    i = (i >= 0) ? i : i ^ 0x7fffffff;
    atomicMax(addr, i);
}
__device__ __forceinline__ float decode_f(int i) { return __int_as_float((i >= 0) ? i : i ^ 0x7fffffff); }

__global__ void k_prep_bounds_init(int* mm) {       // [0..5] pose-space min/max, [6..11] world min/max
    int i = threadIdx.x;
    if (i < 12) mm[i] = ((i % 6) < 3) ? 0x7fffffff : (int)0x80000000;
}

// LBS of the rest vertices (tpose_points_to_pose_points), rest normals rotated by the blended rotation (when given),
// pose -> world, running min/max for get_bounds.
__global__ void k_prep_verts(const float* __restrict__ rverts, const float* __restrict__ rnorm, const float* __restrict__ weights,
                             const float* __restrict__ A, int N, int J, const float* __restrict__ R, const float* __restrict__ Th,
                             float* pverts, float* pnorm, float* wverts, float* wnorm, int* mm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float pv[3] = {0, 0, 0}, wv[3] = {0, 0, 0};
    bool valid = i < N;
    if (valid) {
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; k++) T[k] = 0.f;
        for (int j = 0; j < J; j++) {
            float w = weights[(size_t)i * J + j];
            if (w != 0.f) {
#pragma unroll
                for (int k = 0; k < 12; k++) T[k] += w * A[j * 16 + k];
            }
        }
        float v[3] = {rverts[i * 3], rverts[i * 3 + 1], rverts[i * 3 + 2]};
        for (int a = 0; a < 3; a++) pv[a] = T[a * 4] * v[0] + T[a * 4 + 1] * v[1] + T[a * 4 + 2] * v[2] + T[a * 4 + 3];
        for (int a = 0; a < 3; a++) wv[a] = pv[0] * R[a * 3] + pv[1] * R[a * 3 + 1] + pv[2] * R[a * 3 + 2] + Th[a];     // ppts @ R^T + Th
        for (int a = 0; a < 3; a++) { pverts[i * 3 + a] = pv[a]; if (wverts) wverts[i * 3 + a] = wv[a]; }
        if (rnorm) {
            float n[3] = {rnorm[i * 3], rnorm[i * 3 + 1], rnorm[i * 3 + 2]}, pn[3], wn[3];
            for (int a = 0; a < 3; a++) pn[a] = T[a * 4] * n[0] + T[a * 4 + 1] * n[1] + T[a * 4 + 2] * n[2];
            float inv = 1.f / (sqrtf(pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]) + 1e-12f);
            for (int a = 0; a < 3; a++) pn[a] *= inv;
            for (int a = 0; a < 3; a++) wn[a] = pn[0] * R[a * 3] + pn[1] * R[a * 3 + 1] + pn[2] * R[a * 3 + 2];
            for (int a = 0; a < 3; a++) { pnorm[i * 3 + a] = pn[a]; if (wnorm) wnorm[i * 3 + a] = wn[a]; }
        }
    }
    // warp-level min/max, one atomic per warp and axis
    for (int a = 0; a < 3; a++) {
        float lo = valid ? pv[a] : 3e38f, hi = valid ? pv[a] : -3e38f, wlo = valid ? wv[a] : 3e38f, whi = valid ? wv[a] : -3e38f;
        for (int o = 16; o; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            wlo = fminf(wlo, __shfl_xor_sync(0xffffffffu, wlo, o)); whi = fmaxf(whi, __shfl_xor_sync(0xffffffffu, whi, o));
        }
        if ((threadIdx.x & 31) == 0) { atomic_min_f(mm + a, lo); atomic_max_f(mm + 3 + a, hi); atomic_min_f(mm + 6 + a, wlo); atomic_max_f(mm + 9 + a, whi); }
    }
}

__global__ void k_prep_bounds_finish(const int* __restrict__ mm, float pad, float* pbounds, float* wbounds) {
    int i = threadIdx.x;
    if (i < 6) {
        float s = (i < 3) ? -pad : pad;
        if (pbounds) pbounds[i] = decode_f(mm[i]) + s;
        if (wbounds) wbounds[i] = decode_f(mm[6 + i]) + s;
    }
}

// pytorch3d Meshes._compute_vertex_normals: every face adds its (area-weighted) normal, taken at each of its corners,
// to its three vertices; then F.normalize(eps=1e-6).
__global__ void k_prep_face_normals(const float* __restrict__ verts, const int* __restrict__ faces, int nf, float* acc) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    int id[3] = {faces[f * 3], faces[f * 3 + 1], faces[f * 3 + 2]};
    float3 v[3];
    for (int k = 0; k < 3; k++) v[k] = make3(verts[id[k] * 3], verts[id[k] * 3 + 1], verts[id[k] * 3 + 2]);
    for (int k = 0; k < 3; k++) {          // corner k: cross(v[k+1] - v[k], v[k+2] - v[k])
        float3 a = v[(k + 1) % 3] - v[k], b = v[(k + 2) % 3] - v[k];
        float3 c = make3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
        atomicAdd(acc + id[k] * 3, c.x); atomicAdd(acc + id[k] * 3 + 1, c.y); atomicAdd(acc + id[k] * 3 + 2, c.z);
    }
}
__global__ void k_prep_normals_finish(const float* __restrict__ acc, int N, const float* __restrict__ R, float* pnorm, float* wnorm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float n[3] = {acc[i * 3], acc[i * 3 + 1], acc[i * 3 + 2]};
    float len = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-6f);
    for (int a = 0; a < 3; a++) n[a] /= len;
    for (int a = 0; a < 3; a++) pnorm[i * 3 + a] = n[a];
    if (wnorm)       // normals of the world-space mesh = posed normals rotated (R is a rotation)
        for (int a = 0; a < 3; a++) wnorm[i * 3 + a] = n[0] * R[a * 3] + n[1] * R[a * 3 + 1] + n[2] * R[a * 3 + 2];
}

// ---- rays -------------------------------------------------------------------------------------------------------------------
struct CamDev { float Kinv[9], R[9], T[3], o[3]; };

__device__ __forceinline__ bool prep_ray(const CamDev& c, const float* __restrict__ b, int px, int py, float3& o, float3& d, float& near_, float& far_) {
    // get_rays: pixel_camera = [x, y, 1] @ inv(K)^T ; pixel_world = (pixel_camera - T) @ R ; d = normalize(pixel_world - o)
    float x = (float)px, y = (float)py;
    float pc[3], pw[3];
    for (int a = 0; a < 3; a++) pc[a] = x * c.Kinv[a * 3] + y * c.Kinv[a * 3 + 1] + c.Kinv[a * 3 + 2] - c.T[a];
    for (int a = 0; a < 3; a++) pw[a] = pc[0] * c.R[a] + pc[1] * c.R[3 + a] + pc[2] * c.R[6 + a];
    o = make3(c.o[0], c.o[1], c.o[2]);
    float3 dd = make3(pw[0] - o.x, pw[1] - o.y, pw[2] - o.z);
    float len = sqrtf(dd.x * dd.x + dd.y * dd.y + dd.z * dd.z);
    d = make3(dd.x / len, dd.y / len, dd.z / len);
    // get_full_near_far: |d| is recomputed, tiny components pushed to +-1e-5
    float nd = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    float vd[3] = {d.x / nd, d.y / nd, d.z / nd}, oo[3] = {o.x, o.y, o.z};
    near_ = -3.0e38f; far_ = 3.0e38f;
    for (int a = 0; a < 3; a++) {
        float v = vd[a];
        if (v < 1e-5f && v > -1e-10f) v = 1e-5f;
        if (v > -1e-5f && v < 1e-10f) v = -1e-5f;
        float t0 = (b[a] - oo[a]) / v, t1 = (b[3 + a] - oo[a]) / v;
        near_ = fmaxf(near_, fminf(t0, t1));
        far_ = fminf(far_, fmaxf(t0, t1));
    }
    bool m = near_ < far_;
    near_ /= nd; far_ /= nd;       // get_full_near_far, then get_near_far divides by |d| once more (|d| = 1 up to rounding)
    near_ /= nd; far_ /= nd;
    return m;
}

__global__ void k_prep_rays_count(CamDev c, const float* __restrict__ wbounds, int H, int W, unsigned char* mask, int* blk_cnt) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool m = false;
    if (i < H * W) {
        float3 o, d; float nr, fr;
        m = prep_ray(c, wbounds, i % W, i / W, o, d, nr, fr);
        mask[i] = m ? 1 : 0;
    }
    unsigned bal = __ballot_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&cnt, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = cnt;
}

__global__ void k_prep_rays_write(CamDev c, const float* __restrict__ wbounds, int H, int W, const int* __restrict__ blk_off, int nb,
                                  float* ray_o, float* ray_d, float* near_out, float* far_out, int* n_rays) {
    __shared__ int wcnt[32];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float3 o = make3(0, 0, 0), d = make3(0, 0, 1); float nr = 0.f, fr = 0.f;
    bool m = false;
    if (i < H * W) m = prep_ray(c, wbounds, i % W, i / W, o, d, nr, fr);
    unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0) wcnt[wid] = __popc(bal);
    __syncthreads();
    int base = blk_off[blockIdx.x];
    for (int w = 0; w < wid; w++) base += wcnt[w];
    int ray = base + __popc(bal & ((1u << lane) - 1u));
    if (m) {
        ray_o[ray * 3] = o.x; ray_o[ray * 3 + 1] = o.y; ray_o[ray * 3 + 2] = o.z;
        ray_d[ray * 3] = d.x; ray_d[ray * 3 + 1] = d.y; ray_d[ray * 3 + 2] = d.z;
        near_out[ray] = nr; far_out[ray] = fr;
    }
    if (blockIdx.x == nb - 1 && threadIdx.x == blockDim.x - 1) {      // last thread of the grid: total = its exclusive offset + block count
        int tot = blk_off[blockIdx.x];
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += wcnt[w];
        *n_rays = tot;
    }
}

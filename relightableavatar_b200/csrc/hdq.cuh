// Hierarchical distance query front-end: world->pose, exact 3-NN to the posed SMPL-H vertices,
// signed vertex distances, geodesic filter, shell test, Gaussian-blended inverse LBS to big pose.
// Reference: base_network.py:238-336,365-383; sample_utils.py:103-162; blend_utils.py:125-165,212-329.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------ frame prep
// One block: pose-space bounds of the vertices -> uniform grid; fold the per-frame condition vectors
// into layer biases; copy R / Th / wbounds.
__global__ void k_frame_prep(FrameConst* fc, const float* __restrict__ R, const float* __restrict__ Th,
                             const float* __restrict__ pverts, int nverts, const float* __restrict__ wbounds,
                             const float* __restrict__ poses, const float* __restrict__ mat_cond,
                             int cond,                                                               // C = 3 * n_bones
                             const float* __restrict__ resd_w0, const float* __restrict__ resd_b0,   // (256,63+C)
                             const float* __restrict__ resd_w4, const float* __restrict__ resd_b4,   // (256,256+63+C)
                             const float* __restrict__ rend_w3, const float* __restrict__ rend_b3,   // (256,256+C) or null
                             int* cell_count, float cell_h, float grid2_ratio) {
    __shared__ float smin[3][32], smax[3][32];
    int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = tid; i < nverts; i += blockDim.x)
        for (int a = 0; a < 3; a++) {
            float v = pverts[i * 3 + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) { smin[a][wid] = mn[a]; smax[a][wid] = mx[a]; }
    }
    __syncthreads();
    if (tid == 0) {
        int nw = blockDim.x >> 5;
        float h = cell_h;
        float lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            lo[a] = 1e30f; hi[a] = -1e30f;
            for (int w = 0; w < nw; w++) { lo[a] = fminf(lo[a], smin[a][w]); hi[a] = fmaxf(hi[a], smax[a][w]); }
            h = fmaxf(h, (hi[a] - lo[a]) * (1.0f / (RA_MAX_GRID_DIM - 1)));
        }
        int cells = 1;
        for (int a = 0; a < 3; a++) {
            fc->g_org[a] = lo[a] - 0.5f * h;
            int d = (int)floorf((hi[a] - lo[a] + h) / h) + 1;
            d = min(max(d, 1), RA_MAX_GRID_DIM);
            fc->g_dim[a] = d;
            cells *= d;
        }
        fc->g_h = h;
        fc->g_inv_h = 1.0f / h;
        fc->g_cells = cells;
        float h2 = h * grid2_ratio;
        int cells2 = 1;
        for (int a = 0; a < 3; a++) {
            fc->g2_org[a] = lo[a] - 0.5f * h2;
            int d = (int)floorf((hi[a] - lo[a] + h2) / h2) + 1;
            d = min(max(d, 1), RA_MAX_GRID_DIM);
            fc->g2_dim[a] = d;
            cells2 *= d;
        }
        fc->g2_h = h2;
        fc->g2_inv_h = 1.0f / h2;
        fc->g2_cells = cells2;
        fc->n_occ = 0;
        for (int i = 0; i < 9; i++) fc->R[i] = R[i];
        for (int i = 0; i < 3; i++) fc->Th[i] = Th[i];
        for (int i = 0; i < 6; i++) fc->wb[i] = wbounds ? wbounds[i] : 0.f;
    }
    // the cell counters of the counting sort are all zero here: ra_create zero-fills them and k_grid_scan re-zeroes what it consumed
    // folded biases: one output row per WARP (coalesced weight reads, shuffle reduction), rows strided over the block's warps
    for (int o = wid; o < 256; o += (blockDim.x >> 5)) {
        const float* w0 = resd_w0 + (size_t)o * (63 + cond) + 63;
        const float* w4 = resd_w4 + (size_t)o * (319 + cond) + 256 + 63;
        const bool rend = rend_w3 != nullptr && mat_cond != nullptr;
        const float* w3 = rend ? rend_w3 + (size_t)o * (256 + cond) + 256 : nullptr;
        float s0 = 0.f, s4 = 0.f, s3 = 0.f;
        for (int k = lane; k < cond; k += 32) {
            float c = poses[k];
            s0 += w0[k] * c;
            s4 += w4[k] * c;
            if (rend) s3 += w3[k] * mat_cond[k];
        }
        for (int m = 16; m; m >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, m); s4 += __shfl_xor_sync(0xffffffffu, s4, m); s3 += __shfl_xor_sync(0xffffffffu, s3, m);
        }
        if (lane == 0) {
            fc->resd_b0[o] = resd_b0[o] + s0;
            fc->resd_b4[o] = resd_b4[o] + s4;
            if (rend) fc->rend_b3[o] = rend_b3[o] + s3;
        }
    }
}

struct GridRef { float org[3]; float h, inv_h; int dim[3]; int cells; };
__device__ __forceinline__ GridRef grid_ref(const FrameConst* fc, int level) {
    GridRef g;
    if (level == 0) {
        g.org[0] = fc->g_org[0]; g.org[1] = fc->g_org[1]; g.org[2] = fc->g_org[2]; g.h = fc->g_h; g.inv_h = fc->g_inv_h;
        g.dim[0] = fc->g_dim[0]; g.dim[1] = fc->g_dim[1]; g.dim[2] = fc->g_dim[2]; g.cells = fc->g_cells;
    } else {
        g.org[0] = fc->g2_org[0]; g.org[1] = fc->g2_org[1]; g.org[2] = fc->g2_org[2]; g.h = fc->g2_h; g.inv_h = fc->g2_inv_h;
        g.dim[0] = fc->g2_dim[0]; g.dim[1] = fc->g2_dim[1]; g.dim[2] = fc->g2_dim[2]; g.cells = fc->g2_cells;
    }
    return g;
}
__device__ __forceinline__ int cell_of(const GridRef& g, float3 p, int& cx, int& cy, int& cz) {
    cx = min(max((int)floorf((p.x - g.org[0]) * g.inv_h), 0), g.dim[0] - 1);
    cy = min(max((int)floorf((p.y - g.org[1]) * g.inv_h), 0), g.dim[1] - 1);
    cz = min(max((int)floorf((p.z - g.org[2]) * g.inv_h), 0), g.dim[2] - 1);
    return (cz * g.dim[1] + cy) * g.dim[0] + cx;
}

// level 0: vertices in input order; level 1: vertices taken from the level-0 sorted array (so ids refer to it)
__global__ void k_grid_count(const FrameConst* fc, int level, const float* __restrict__ pverts, const float4* __restrict__ spos,
                             int nverts, int* cell_count, int* vert_cell) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nverts) return;
    GridRef g = grid_ref(fc, level);
    int cx, cy, cz;
    float3 p = level == 0 ? make3(pverts[i * 3], pverts[i * 3 + 1], pverts[i * 3 + 2]) : make3(spos[i].x, spos[i].y, spos[i].z);
    int c = cell_of(g, p, cx, cy, cz);
    vert_cell[i] = c;
    atomicAdd(&cell_count[c], 1);
}

// Counting sort, scatter step, made DETERMINISTIC: k_grid_order drops the vertex ids of a cell into its run in whatever order the
// atomics retire; the fill kernels then place vertex i at the rank of i among the ids of its run.  The order of the vertices inside a
// cell decides which of two exactly equidistant vertices the 3-NN keeps -- without this a frame could differ from the same frame
// rendered by another handle in one visibility entry out of millions (observed: the first handle of a process vs. the later ones).
__global__ void k_grid_order(int nverts, const int* __restrict__ vert_cell, const int* __restrict__ cell_start, int* cell_fill, int* order) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nverts) return;
    int c = vert_cell[i];
    order[cell_start[c] + atomicAdd(&cell_fill[c], 1)] = i;
}
__device__ __forceinline__ int grid_rank_slot(int i, const int* __restrict__ vert_cell, const int* __restrict__ cell_start, const int* __restrict__ order) {
    const int c = vert_cell[i];
    const int s = cell_start[c], e = cell_start[c + 1];
    int rank = 0;
    for (int k = s; k < e; k++) rank += (order[k] < i);
    return s + rank;
}
__global__ void k_grid_fill2(const float4* __restrict__ spos, int nverts, const int* __restrict__ vert_cell,
                             const int* __restrict__ cell_start, const int* __restrict__ order, float4* pos2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nverts) return;
    int dst = grid_rank_slot(i, vert_cell, cell_start, order);
    float4 v = spos[i];
    pos2[dst] = make_float4(v.x, v.y, v.z, __int_as_float(i));
}

// Far-field hierarchy of the exact 3-NN: the occupied coarse cells with the tight bounding box of their vertices, grouped
// into super cells (blocks of f x f x f coarse cells, f >= 4 so that there are at most RA_MAX_SUP of them) that carry the union
// box and the range of their cells.  A far query tests the ~10-30 super boxes first and descends only into those that can hold
// a closer vertex, instead of testing every one of the ~300 coarse boxes (twice).  One block (the coarse grid has a few
// hundred to a few thousand cells).
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v)); else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v)); else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__global__ void __launch_bounds__(1024) k_grid_occ(FrameConst* fc, const int* __restrict__ cell_start2, const float4* __restrict__ pos2, SortedVerts sv) {
    __shared__ int s_cnt[RA_MAX_SUP], s_start[RA_MAX_SUP], s_fill[RA_MAX_SUP];
    __shared__ float s_lo[RA_MAX_SUP][3], s_hi[RA_MAX_SUP][3];
    __shared__ int s_n, s_nsup;
    const int n = fc->g2_cells;
    const int d0 = fc->g2_dim[0], d1 = fc->g2_dim[1], d2 = fc->g2_dim[2];
    int f = 4;
    while (((d0 + f - 1) / f) * ((d1 + f - 1) / f) * ((d2 + f - 1) / f) > RA_MAX_SUP) f *= 2;
    const int e0 = (d0 + f - 1) / f, e1 = (d1 + f - 1) / f, e2 = (d2 + f - 1) / f, nsup = e0 * e1 * e2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = tid; i < nsup; i += blockDim.x) {
        s_cnt[i] = 0; s_fill[i] = 0;
        for (int a = 0; a < 3; a++) { s_lo[i][a] = 3e38f; s_hi[i][a] = -3e38f; }
    }
    if (tid == 0) { s_n = 0; s_nsup = 0; }
    __syncthreads();
    for (int c = warp; c < n; c += nwarps) {           // one warp per coarse cell: tight box of its vertices
        const int s = cell_start2[c], e = cell_start2[c + 1];
        if (e <= s) continue;
        float3 lo = make3(3e38f, 3e38f, 3e38f), hi = make3(-3e38f, -3e38f, -3e38f);
        for (int v = s + lane; v < e; v += 32) {
            float4 q = pos2[v];
            lo = make3(fminf(lo.x, q.x), fminf(lo.y, q.y), fminf(lo.z, q.z));
            hi = make3(fmaxf(hi.x, q.x), fmaxf(hi.y, q.y), fmaxf(hi.z, q.z));
        }
        for (int m = 16; m; m >>= 1) {
            lo = make3(fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, m)), fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, m)), fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, m)));
            hi = make3(fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, m)), fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, m)), fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, m)));
        }
        if (lane == 0) {
            const int cx = c % d0, cy = (c / d0) % d1, cz = c / (d0 * d1);
            const int sup = ((cz / f) * e1 + cy / f) * e0 + cx / f;
            const int idx = atomicAdd(&s_n, 1);
            if (idx < RA_MAX_OCC) {
                sv.occ_tmp[2 * idx] = make_float4(lo.x, lo.y, lo.z, __int_as_float(s));
                sv.occ_tmp[2 * idx + 1] = make_float4(hi.x, hi.y, hi.z, __int_as_float(e));
                sv.occ_sup[idx] = sup;
                atomicAdd(&s_cnt[sup], 1);
                atomic_min_f(&s_lo[sup][0], lo.x); atomic_min_f(&s_lo[sup][1], lo.y); atomic_min_f(&s_lo[sup][2], lo.z);
                atomic_max_f(&s_hi[sup][0], hi.x); atomic_max_f(&s_hi[sup][1], hi.y); atomic_max_f(&s_hi[sup][2], hi.z);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {                                      // ranges of the occupied super cells (a few dozen entries)
        int run = 0, k = 0;
        for (int i = 0; i < nsup; i++) {
            s_start[i] = run;
            if (s_cnt[i] > 0) {
                sv.sup_lo[k] = make_float4(s_lo[i][0], s_lo[i][1], s_lo[i][2], __int_as_float(run));
                sv.sup_hi[k] = make_float4(s_hi[i][0], s_hi[i][1], s_hi[i][2], __int_as_float(run + s_cnt[i]));
                k++;
            }
            run += s_cnt[i];
        }
        fc->n_occ = s_n;             // > RA_MAX_OCC: the consumers fall back to brute force
        fc->n_sup = k;
    }
    __syncthreads();
    const int nocc = min(s_n, RA_MAX_OCC);
    for (int i = tid; i < nocc; i += blockDim.x) {       // group the cell boxes by super cell
        const int sup = sv.occ_sup[i];
        const int dst = s_start[sup] + atomicAdd(&s_fill[sup], 1);
        sv.occ_lo[dst] = sv.occ_tmp[2 * i];
        sv.occ_hi[dst] = sv.occ_tmp[2 * i + 1];
    }
}

// single block exclusive scan over g_cells (+1) entries; also resets the fill cursors
__global__ void k_grid_scan(const FrameConst* fc, int level, int* cell_count, int* cell_start, int* cell_fill) {
    __shared__ int carry;
    __shared__ int wsum[32];
    int n = level == 0 ? fc->g_cells : fc->g2_cells;
    int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 4 * blockDim.x) {        // four consecutive cells per thread
        const int i0 = base + tid * 4;
        int v[4], t = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = (i0 + k < n) ? cell_count[i0 + k] : 0; t += v[k]; }
        int x = t;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        int excl = carry + (wid ? wsum[wid - 1] : 0) + x - t;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) { cell_start[i0 + k] = excl; cell_fill[i0 + k] = 0; cell_count[i0 + k] = 0; }
            excl += v[k];
        }
        __syncthreads();
        if (tid == blockDim.x - 1) carry = excl;
        __syncthreads();
    }
    if (tid == 0) cell_start[n] = carry;
}

// scatter vertices into cell order and pre-blend the per-vertex skinning transforms:
// T_v = sum_j weights[v][j] * A_j (and big_A_j).  sum_k w_k T_{nn_k} == sum_j (sum_k w_k W[nn_k][j]) A_j
// (blend_transform, blend_utils.py:212-218) up to fp32 re-association.
__global__ void k_grid_fill(const FrameConst* fc, const float* __restrict__ pverts, const float* __restrict__ pnorm,
                            const float* __restrict__ tverts, const float* __restrict__ weights,
                            const float* __restrict__ A, const float* __restrict__ bigA, int nverts, int nbones,
                            const int* __restrict__ vert_cell, const int* __restrict__ cell_start, const int* __restrict__ order,
                            SortedVerts sv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nverts) return;
    int dst = grid_rank_slot(i, vert_cell, cell_start, order);
    sv.pos[dst] = make_float4(pverts[i * 3], pverts[i * 3 + 1], pverts[i * 3 + 2], __int_as_float(i));
    sv.nrm[dst] = make_float4(pnorm[i * 3], pnorm[i * 3 + 1], pnorm[i * 3 + 2], 0.f);
    sv.tv[dst] = make_float4(tverts[i * 3], tverts[i * 3 + 1], tverts[i * 3 + 2], 0.f);
    float T[24];
#pragma unroll
    for (int k = 0; k < 24; k++) T[k] = 0.f;
    for (int j = 0; j < nbones; j++) {
        float w = weights[(size_t)i * nbones + j];
        if (w != 0.f) {
#pragma unroll
            for (int k = 0; k < 12; k++) {
                T[k] += w * A[j * 16 + k];
                T[12 + k] += w * bigA[j * 16 + k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 24; k++) sv.T[(size_t)dst * 24 + k] = T[k];
}

// ------------------------------------------------------------------------------------------ neighbourhood lists
// Enumerate, for the cube shell of radius r around cell (cx,cy,cz) (r == 1: the whole 3x3x3 block), the contiguous
// runs of the cell-sorted vertex array: full x-runs on the z / y faces, the two end cells on the inner rows.
template <typename F>
__device__ __forceinline__ void nb_for_each_run(const GridRef& g, const int* __restrict__ cell_start, int cx, int cy, int cz, int r, F&& f) {
    const int dx_ = g.dim[0], dy_ = g.dim[1], dz_ = g.dim[2];
    int z0 = max(cz - r, 0), z1 = min(cz + r, dz_ - 1);
    int y0 = max(cy - r, 0), y1 = min(cy + r, dy_ - 1);
    int x0 = max(cx - r, 0), x1 = min(cx + r, dx_ - 1);
    for (int z = z0; z <= z1; z++) {
        bool zf = (r == 1) || (z == cz - r) || (z == cz + r);
        for (int y = y0; y <= y1; y++) {
            bool yf = zf || (y == cy - r) || (y == cy + r);
            int row = (z * dy_ + y) * dx_;
            if (yf) {
                f(__ldg(&cell_start[row + x0]), __ldg(&cell_start[row + x1 + 1]));
            } else {
                if (cx - r >= 0) f(__ldg(&cell_start[row + cx - r]), __ldg(&cell_start[row + cx - r + 1]));
                if (cx + r < dx_) f(__ldg(&cell_start[row + cx + r]), __ldg(&cell_start[row + cx + r + 1]));
            }
        }
    }
}

// counts of all levels, one thread per cell: cnt[level * (RA_MAX_CELLS + 1) + cell]
__global__ void k_nb_count(const FrameConst* fc, const int* __restrict__ cell_start, int* cnt, unsigned char* mask) {
    GridRef g = grid_ref(fc, 0);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.cells; c += gridDim.x * blockDim.x) {
        int cx = c % g.dim[0], cy = (c / g.dim[0]) % g.dim[1], cz = c / (g.dim[0] * g.dim[1]);
        unsigned m = 0;
        for (int lv = 0; lv < RA_NB_LEVELS; lv++) {
            int n = 0;
            nb_for_each_run(g, cell_start, cx, cy, cz, lv + 1, [&](int s, int e) { n += e - s; });
            cnt[lv * (RA_MAX_CELLS + 1) + c] = n;
            if (n > 0) m |= 1u << lv;
        }
        mask[c] = (unsigned char)m;
    }
}

// exclusive scans of the RA_NB_LEVELS count arrays (one block per level)
__global__ void k_nb_scan(const FrameConst* fc, const int* __restrict__ cnt, SortedVerts sv) {
    __shared__ int carry;
    __shared__ int wsum[32];
    const int lv = blockIdx.x;
    const int n = fc->g_cells;
    const int* in = cnt + lv * (RA_MAX_CELLS + 1);
    int* out = sv.nb_start[lv];
    int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 4 * blockDim.x) {        // four consecutive cells per thread
        const int i0 = base + tid * 4;
        int v[4], t = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; t += v[k]; }
        int x = t;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            wsum[lane] = s;
        }
        __syncthreads();
        int excl = carry + (wid ? wsum[wid - 1] : 0) + x - t;
#pragma unroll
        for (int k = 0; k < 4; k++) { if (i0 + k < n) out[i0 + k] = excl; excl += v[k]; }
        __syncthreads();
        if (tid == blockDim.x - 1) carry = excl;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

// one warp per (cell, level): copy the runs into the cell's contiguous list.  Lanes enumerate the rows of the (2r+1)^2 square
// in parallel (most rows of a shell hold no vertex: no serial chain of dependent cell_start loads through them), a warp scan
// turns the row counts into offsets, and only the non-empty rows are copied (cooperatively).
__global__ void k_nb_fill(const FrameConst* fc, const int* __restrict__ cell_start, const float4* __restrict__ pos, SortedVerts sv) {
    GridRef g = grid_ref(fc, 0);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int dx_ = g.dim[0], dy_ = g.dim[1], dz_ = g.dim[2];
    for (int w = warp; w < g.cells * RA_NB_LEVELS; w += nwarps) {
        const int c = w / RA_NB_LEVELS, lv = w % RA_NB_LEVELS;
        int dst = sv.nb_start[lv][c];
        if (sv.nb_start[lv][c + 1] == dst) continue;
        const int cx = c % dx_, cy = (c / dx_) % dy_, cz = c / (dx_ * dy_);
        const int r = lv + 1, side = 2 * r + 1;
        float4* out = sv.nb_pos[lv];
        for (int j0 = 0; j0 < side * side; j0 += 32) {
            const int j = j0 + lane;
            int s0 = 0, e0 = 0, s1 = 0, e1 = 0;          // up to two runs per row
            if (j < side * side) {
                const int z = cz - r + j / side, y = cy - r + j % side;
                if (z >= 0 && z < dz_ && y >= 0 && y < dy_) {
                    const int row = (z * dy_ + y) * dx_;
                    const bool face = (r == 1) || z == cz - r || z == cz + r || y == cy - r || y == cy + r;
                    if (face) {
                        s0 = __ldg(&cell_start[row + max(cx - r, 0)]); e0 = __ldg(&cell_start[row + min(cx + r, dx_ - 1) + 1]);
                    } else {
                        if (cx - r >= 0) { s0 = __ldg(&cell_start[row + cx - r]); e0 = __ldg(&cell_start[row + cx - r + 1]); }
                        if (cx + r < dx_) { s1 = __ldg(&cell_start[row + cx + r]); e1 = __ldg(&cell_start[row + cx + r + 1]); }
                    }
                }
            }
            const int n0 = e0 - s0, n1 = e1 - s1, cnt = n0 + n1;
            int incl = cnt;
            for (int o = 1; o < 32; o <<= 1) { int y2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y2; }
            // one lane per ENTRY of these 32 rows (rows in order, a row's first run before its second): the entry's row by binary search
            // over the rows' running counts.  (A loop over the non-empty rows with the lanes striding over each 1-5 vertex run left most
            // lanes idle: 150 us per frame for ~5 M entries.)
            const int excl = incl - cnt;
            const int T = __shfl_sync(0xffffffffu, incl, 31);
            for (int e0 = 0; e0 < T; e0 += 32) {
                const int e = e0 + lane;
                int r = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, r + step - 1);
                    if (v <= e) r += step;
                }
                r = min(r, 31);
                const int bex = __shfl_sync(0xffffffffu, excl, r), bs0 = __shfl_sync(0xffffffffu, s0, r), bn0 = __shfl_sync(0xffffffffu, n0, r);
                const int bs1 = __shfl_sync(0xffffffffu, s1, r);
                if (e < T) {
                    const int k = e - bex;
                    const int src = (k < bn0) ? bs0 + k : bs1 + (k - bn0);
                    const float4 q = __ldg(&pos[src]);
                    out[dst + e] = make_float4(q.x, q.y, q.z, __int_as_float(src));
                }
            }
            dst += T;
        }
    }
}

// ------------------------------------------------------------------------------------------ exact 3-NN
#ifdef RA_KNN_STATS
__device__ unsigned long long g_knn_stats[12];   // [0] near-path queries, [1] far-path queries, [2] far cells scanned, [3] far verts scanned,
                                                // [4] near verts scanned, [5] near rings visited, [6] packets searched, [7] packets refused (too spread out),
                                                // [8] far lanes served by packets, [9] cells / [10] vertices scanned by packets
#define KNN_STAT(i, v) atomicAdd(&g_knn_stats[i], (unsigned long long)(v))
#else
#define KNN_STAT(i, v)
#endif
__device__ __forceinline__ void knn_insert(KnnOut& o, float d2, int id) {
    if (d2 < o.d2[2]) {
        if (d2 < o.d2[1]) {
            o.d2[2] = o.d2[1]; o.id[2] = o.id[1];
            if (d2 < o.d2[0]) {
                o.d2[1] = o.d2[0]; o.id[1] = o.id[0];
                o.d2[0] = d2; o.id[0] = id;
            } else { o.d2[1] = d2; o.id[1] = id; }
        } else { o.d2[2] = d2; o.id[2] = id; }
    }
}

// squared L2 as (dx^2 + dy^2) + dz^2 with separately rounded products (matches the oracle's elementwise form)
__device__ __forceinline__ float dist2_ref(float3 p, float4 v) {
    float dx = p.x - v.x, dy = p.y - v.y, dz = p.z - v.z;
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ void knn_range(KnnOut& o, float3 p, const float4* __restrict__ pos, int s, int e) {
    for (int v = s; v < e; v++) knn_insert(o, dist2_ref(p, __ldg(&pos[v])), v);
}

__device__ __forceinline__ float bbox_dist2(float3 p, float4 lo, float4 hi) {
    float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.f);
    float dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.f);
    float dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// Exact 3 nearest vertices (K=3 of pytorch3d.ops.knn_points at sample_utils.py:122).
// Near the body: expanding shells over the fine grid (cell ~4 cm).  Farther away: branch-and-bound over the list of
// occupied coarse cells (tight vertex bounding boxes): nearest cell first, then every cell whose box is closer than
// the current 3rd-best distance.  Both are exact; brute force only if the cell list overflowed.
// near phase: the cell's neighbourhood lists, level by level (3x3x3 block, then the radius-2 and radius-3 shells); after each
// level the result is final if the 3rd best distance is provably inside the explored block.  Returns true when final
// (o may hold seed candidates otherwise).
__device__ __forceinline__ bool knn3_near(const FrameConst* __restrict__ fc, const SortedVerts& sv, float3 p, KnnOut& o) {
    o.d2[0] = o.d2[1] = o.d2[2] = 3.0e38f;
    o.id[0] = o.id[1] = o.id[2] = -1;      // empty slots: never equal to a real sorted-vertex id (knn3_far de-duplicates against them)
    const GridRef g = grid_ref(fc, 0);
    int cx, cy, cz;
    const int c = cell_of(g, p, cx, cy, cz);
    const int dx_ = g.dim[0], dy_ = g.dim[1], dz_ = g.dim[2];
    const float h = g.h;
    const unsigned lvmask = __ldg(&sv.nb_mask[c]);
    if (lvmask == 0) return false;          // nothing within RA_NB_LEVELS cells: nothing to certify, straight to the far phase
#pragma unroll 1
    for (int lv = 0; lv < RA_NB_LEVELS; lv++) {
        const int r = lv + 1;
        if ((lvmask >> lv) & 1u) {          // (an empty level adds no candidates; the larger explored block may still certify the result)
            const int s = __ldg(&sv.nb_start[lv][c]), e = __ldg(&sv.nb_start[lv][c + 1]);
            const float4* __restrict__ lst = sv.nb_pos[lv];
            KNN_STAT(4, e - s); KNN_STAT(5, 1);
            int v = s;
            for (; v + 4 <= e; v += 4) {          // four independent 16 B loads in flight per lane (the scan is L2-latency bound)
                const float4 q0 = __ldg(&lst[v]), q1 = __ldg(&lst[v + 1]), q2 = __ldg(&lst[v + 2]), q3 = __ldg(&lst[v + 3]);
                const float e0 = dist2_ref(p, q0), e1 = dist2_ref(p, q1), e2 = dist2_ref(p, q2), e3 = dist2_ref(p, q3);
                if (fminf(fminf(e0, e1), fminf(e2, e3)) < o.d2[2]) {
                    knn_insert(o, e0, __float_as_int(q0.w)); knn_insert(o, e1, __float_as_int(q1.w));
                    knn_insert(o, e2, __float_as_int(q2.w)); knn_insert(o, e3, __float_as_int(q3.w));
                }
            }
            for (; v < e; v++) {
                float4 q = __ldg(&lst[v]);
                knn_insert(o, dist2_ref(p, q), __float_as_int(q.w));
            }
        }
        // distance from p to the nearest face of the explored block that still has unexplored cells behind it
        float bound = 3.0e38f;
        if (cx - r > 0) bound = fminf(bound, p.x - (g.org[0] + (cx - r) * h));
        if (cx + r < dx_ - 1) bound = fminf(bound, (g.org[0] + (cx + r + 1) * h) - p.x);
        if (cy - r > 0) bound = fminf(bound, p.y - (g.org[1] + (cy - r) * h));
        if (cy + r < dy_ - 1) bound = fminf(bound, (g.org[1] + (cy + r + 1) * h) - p.y);
        if (cz - r > 0) bound = fminf(bound, p.z - (g.org[2] + (cz - r) * h));
        if (cz + r < dz_ - 1) bound = fminf(bound, (g.org[2] + (cz + r + 1) * h) - p.z);
        // 1e-4 relative safety margin against fp32 rounding of the face coordinates
        if (bound > 0.f && o.d2[2] <= bound * bound * 0.9999f) return true;
        if (bound == 3.0e38f) return true;   // whole grid explored
    }
    return false;
}

// far phase of ONE lane: branch-and-bound over the two-level box hierarchy (super cells -> occupied coarse cells -> vertices),
// seeded with whatever the near phase found.  Used when most lanes of a warp are far from the body (rays entering the box: the
// lanes walk nearly the same boxes); stragglers of mostly-near warps are served by the whole warp instead (knn3_warp).
__device__ void knn3_far(const FrameConst* __restrict__ fc, const SortedVerts& sv, int nverts, float3 p, KnnOut& o) {
    // The candidates found so far are real vertices: o.d2[2] (if finite) is an upper bound of the true 3rd distance.
    if (fc->n_occ > RA_MAX_OCC) { o.d2[0] = o.d2[1] = o.d2[2] = 3.0e38f; knn_range(o, p, sv.pos, 0, nverts); return; }
    const int ns = fc->n_sup;
    auto scan = [&](int c) {
        int s = __float_as_int(__ldg(&sv.occ_lo[c]).w), e = __float_as_int(__ldg(&sv.occ_hi[c]).w);
        KNN_STAT(2, 1); KNN_STAT(3, e - s);
        for (int v = s; v < e; v++) {
            float4 q = __ldg(&sv.pos2[v]);
            float d2 = dist2_ref(p, q);
            int id = __float_as_int(q.w);
            if (d2 < o.d2[2] && id != o.id[0] && id != o.id[1] && id != o.id[2]) knn_insert(o, d2, id);
        }
    };
    int bc = -1;
    if (o.d2[2] >= 3.0e38f) {          // fewer than 3 candidates yet: the nearest cell of the nearest super cell first
        float best = 3.0e38f; int bs = 0;
        for (int s = 0; s < ns; s++) {
            float lb = bbox_dist2(p, __ldg(&sv.sup_lo[s]), __ldg(&sv.sup_hi[s]));
            if (lb < best) { best = lb; bs = s; }
        }
        const int c0 = __float_as_int(__ldg(&sv.sup_lo[bs]).w), c1 = __float_as_int(__ldg(&sv.sup_hi[bs]).w);
        best = 3.0e38f;
        for (int c = c0; c < c1; c++) {
            float lb = bbox_dist2(p, __ldg(&sv.occ_lo[c]), __ldg(&sv.occ_hi[c]));
            if (lb < best) { best = lb; bc = c; }
        }
        scan(bc);
    }
    for (int s = 0; s < ns; s++) {
        const float4 slo = __ldg(&sv.sup_lo[s]), shi = __ldg(&sv.sup_hi[s]);
        if (bbox_dist2(p, slo, shi) * 0.9999f > o.d2[2]) continue;      // 1e-4 relative slack: boxes round independently of dist2_ref
        const int c0 = __float_as_int(slo.w), c1 = __float_as_int(shi.w);
        for (int c = c0; c < c1; c++) {
            if (c == bc) continue;
            float lb = bbox_dist2(p, __ldg(&sv.occ_lo[c]), __ldg(&sv.occ_hi[c]));
            if (lb * 0.9999f <= o.d2[2]) scan(c);
        }
    }
}

__device__ __forceinline__ float warp_min_f(float v) {
    for (int m = 16; m; m >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
// warp arg-min of (key, index): smallest key, ties to the smallest index
__device__ __forceinline__ int warp_argmin(float key, int idx) {
    for (int m = 16; m; m >>= 1) {
        float ok = __shfl_xor_sync(0xffffffffu, key, m); int oi = __shfl_xor_sync(0xffffffffu, idx, m);
        if (ok < key || (ok == key && oi < idx)) { key = ok; idx = oi; }
    }
    return idx;
}

// Packet form of the far phase for COHERENT warps (the lanes' points lie close together: parallel rays of neighbouring pixels).
// The far lanes search together: the box hierarchy is walked ONCE for the packet -- lanes test 32 boxes at a time against the
// sphere (centre of the packet's points, radius = largest current 3rd-neighbour bound + packet radius) -- and every vertex of a
// qualifying cell is loaded once (warp-uniform address) and offered to all far lanes, each keeping its own exact top 3 with the
// same distance arithmetic as the other paths.  Exact: a vertex within lane l's true 3rd distance r_l of p_l is within
// r_l + |p_l - c| <= sqrt(Bmax) + rho of the centre c, so its cell's box passes the sphere test; bounds only ever shrink towards
// the true values because they are distances of real vertices.  Against one-query-per-warp this removes the per-query box walk,
// bootstrap and 5-round shuffle merge (about 4/5 of its instructions).  Returns false (nothing done) when the packet is too
// spread out to share one sphere; MUST be called by all 32 lanes.
__device__ __forceinline__ float warp_max_nonneg(float v) { return __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(v))); }
__device__ int g_pkt_min = RA_PKT_MIN;
__device__ float g_pkt_rho = RA_PKT_RHO;       // largest packet radius served as one search (env RA_PKT_RHO overrides it for experiments)
__device__ bool knn3_packet(const FrameConst* __restrict__ fc, const SortedVerts& sv, float3 p, unsigned need, KnnOut& o) {
    const int lane = threadIdx.x & 31;
    const int n = fc->n_occ, ns = fc->n_sup;
    if (n > RA_MAX_OCC) return false;
    const bool act = (need >> lane) & 1u;
    float3 lo = act ? p : make3(3e38f, 3e38f, 3e38f), hi = act ? p : make3(-3e38f, -3e38f, -3e38f);
#pragma unroll
    for (int m = 16; m; m >>= 1) {
        lo = make3(fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, m)), fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, m)), fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, m)));
        hi = make3(fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, m)), fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, m)), fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, m)));
    }
    const float3 c = make3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
    const float ex = hi.x - c.x, ey = hi.y - c.y, ez = hi.z - c.z;
    const float rho = sqrtf(ex * ex + ey * ey + ez * ez) * 1.0001f + 1e-6f;      // every far point is within rho of c
    if (rho > g_pkt_rho) { if (lane == 0) KNN_STAT(7, 1); return false; }
    if (lane == 0) { KNN_STAT(6, 1); KNN_STAT(8, __popc(need)); }
    const float B = act ? o.d2[2] : 0.f;          // seeds of the near phase are real vertices: upper bound of the true 3rd distance (or 3e38)
    KnnOut w;
    w.d2[0] = w.d2[1] = w.d2[2] = 3.0e38f; w.id[0] = w.id[1] = w.id[2] = -1;
    __shared__ float4 pkt_buf[8][32];              // per-warp staging of a cell's vertices (blocks of up to 8 warps)
    float4* buf = pkt_buf[(threadIdx.x >> 5) & 7];
    // one coarse cell: its box and vertex range come in the same two loads; each far lane tests the box against ITS OWN point and
    // bound (as tight as a single-query search), the cell is scanned if any lane needs it, by the lanes that need it
    auto scan = [&](int cc) {
        const float4 clo = __ldg(&sv.occ_lo[cc]), chi = __ldg(&sv.occ_hi[cc]);
        const bool mine = act && bbox_dist2(p, clo, chi) * 0.9999f <= fminf(B, w.d2[2]);
        if (!__any_sync(0xffffffffu, mine)) return;
        const int s = __float_as_int(clo.w), e = __float_as_int(chi.w);
        if (lane == 0) { KNN_STAT(9, 1); KNN_STAT(10, e - s); }
        // the cell's vertices come in with ONE coalesced load per 32 and are handed to the lanes through shared memory (broadcast
        // reads: fixed ~25-cycle latency) -- warp-uniform global loads of one vertex at a time left the loop waiting on L1 / L2
        // (46 % of its stall samples: profiles/r02_ncu_k_trace_shadow_floor.txt)
        for (int base = s; base < e; base += 32) {
            const int nv = min(32, e - base);
            __syncwarp();
            if (lane < nv) buf[lane] = __ldg(&sv.pos2[base + lane]);
            __syncwarp();
            if (mine) {
                int k = 0;
                for (; k + 4 <= nv; k += 4) {
                    const float4 q0 = buf[k], q1 = buf[k + 1], q2 = buf[k + 2], q3 = buf[k + 3];
                    const float e0 = dist2_ref(p, q0), e1 = dist2_ref(p, q1), e2 = dist2_ref(p, q2), e3 = dist2_ref(p, q3);
                    if (fminf(fminf(e0, e1), fminf(e2, e3)) < w.d2[2]) {
                        knn_insert(w, e0, __float_as_int(q0.w)); knn_insert(w, e1, __float_as_int(q1.w));
                        knn_insert(w, e2, __float_as_int(q2.w)); knn_insert(w, e3, __float_as_int(q3.w));
                    }
                }
                for (; k < nv; k++) {
                    const float4 q = buf[k];
                    knn_insert(w, dist2_ref(p, q), __float_as_int(q.w));
                }
            }
        }
    };
    auto radius2 = [&]() {                         // squared radius of the packet sphere from the lanes' current bounds
        const float bm = warp_max_nonneg(act ? fminf(B, w.d2[2]) : 0.f);
        if (bm >= 3.0e38f) return 3.0e38f;
        const float R = sqrtf(bm) * 1.0001f + rho;
        return R * R * 1.0001f;
    };
    // bootstrap: the cell nearest to the centre
    int bc;
    {
        float best = 3.0e38f; int bs = 0;
        for (int s = lane; s < ns; s += 32) {
            const float lb = bbox_dist2(c, __ldg(&sv.sup_lo[s]), __ldg(&sv.sup_hi[s]));
            if (lb < best) { best = lb; bs = s; }
        }
        bs = warp_argmin(best, bs);
        const int c0 = __float_as_int(__ldg(&sv.sup_lo[bs]).w), c1 = __float_as_int(__ldg(&sv.sup_hi[bs]).w);
        best = 3.0e38f; bc = c0;
        for (int k = c0 + lane; k < c1; k += 32) {
            const float lb = bbox_dist2(c, __ldg(&sv.occ_lo[k]), __ldg(&sv.occ_hi[k]));
            if (lb < best) { best = lb; bc = k; }
        }
        bc = warp_argmin(best, bc);
        // (a compact loop of its own: the unrolled scan above is instantiated once, for the main walk)
        const int s = __float_as_int(__ldg(&sv.occ_lo[bc]).w), e = __float_as_int(__ldg(&sv.occ_hi[bc]).w);
        if (lane == 0) { KNN_STAT(9, 1); KNN_STAT(10, e - s); }
        for (int base = s; base < e; base += 32) {
            const int nv = min(32, e - base);
            __syncwarp();
            if (lane < nv) buf[lane] = __ldg(&sv.pos2[base + lane]);
            __syncwarp();
            if (act)
#pragma unroll 1
                for (int k = 0; k < nv; k++) {
                    const float4 q = buf[k];
                    knn_insert(w, dist2_ref(p, q), __float_as_int(q.w));
                }
        }
    }
    float R2 = radius2();
    for (int s0 = 0; s0 < ns; s0 += 32) {               // super cells: one per lane
        const int si = s0 + lane;
        bool squal = false;
        if (si < ns) squal = bbox_dist2(c, __ldg(&sv.sup_lo[si]), __ldg(&sv.sup_hi[si])) * 0.9999f <= R2;
        unsigned sm = __ballot_sync(0xffffffffu, squal);
        while (sm) {
            const int ss = s0 + __ffs(sm) - 1;
            sm &= sm - 1;
            const float4 slo = __ldg(&sv.sup_lo[ss]), shi = __ldg(&sv.sup_hi[ss]);
            if (bbox_dist2(c, slo, shi) * 0.9999f > R2) continue;      // the bound has shrunk since the lane-parallel test
            const int c0 = __float_as_int(slo.w), c1 = __float_as_int(shi.w);
            for (int cb = c0; cb < c1; cb += 32) {      // its coarse cells: one per lane
                const int k = cb + lane;
                bool qual = false;
                if (k < c1 && k != bc) qual = bbox_dist2(c, __ldg(&sv.occ_lo[k]), __ldg(&sv.occ_hi[k])) * 0.9999f <= R2;
                unsigned qm = __ballot_sync(0xffffffffu, qual);
                while (qm) {
                    scan(cb + __ffs(qm) - 1);
                    qm &= qm - 1;
                }
            }
            R2 = radius2();
        }
    }
    if (act) o = w;
    return true;
}

// Warp-cooperative exact 3-NN.  MUST be called by all 32 lanes (warp-uniform call site).  Every lane first runs the
// fine-grid search for its own point; the points it cannot finish (farther than 4 fine cells from every vertex) are then
// processed one at a time by the WHOLE warp over the two-level box hierarchy (super cells -> occupied coarse cells ->
// vertices): lanes split the boxes of a level for the distance test and stride over the vertices of the qualifying cells,
// followed by one shuffle merge of the per-lane triples.  When more than `coop_max` lanes are far (surface rays entering the
// box: coherent lanes) every lane walks the hierarchy itself instead -- measured on the 512^2 frame: surface stage 2.4 ms
// per-lane vs 3.4 ms cooperative, shadow stage 18.2 vs 17.9 ms.
__device__ void knn3_warp(const FrameConst* __restrict__ fc, const SortedVerts& sv, int nverts, float3 p, bool active, KnnOut& o, int coop_max = 12,
                          bool packet = false) {
    const int lane = threadIdx.x & 31;
    bool done = true;
    if (active) done = knn3_near(fc, sv, p, o);
    if (active) { if (done) KNN_STAT(0, 1); else KNN_STAT(1, 1); }
    unsigned need = __ballot_sync(0xffffffffu, active && !done);
    if (!need) return;
    if (packet && __popc(need) >= g_pkt_min && knn3_packet(fc, sv, p, need, o)) return;
    if (__popc(need) > coop_max) {
        if (active && !done) knn3_far(fc, sv, nverts, p, o);
        return;
    }
    const int n = fc->n_occ, ns = fc->n_sup;
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const float3 q = make3(__shfl_sync(0xffffffffu, p.x, src), __shfl_sync(0xffffffffu, p.y, src), __shfl_sync(0xffffffffu, p.z, src));
        // seeds of the near phase are real vertices: their 3rd distance (if finite) bounds the true one.  Every vertex within that
        // bound lies in a cell whose box qualifies, so the scan below finds the complete answer on its own.
        float B = __shfl_sync(0xffffffffu, o.d2[2], src);
        KnnOut w;
        w.d2[0] = w.d2[1] = w.d2[2] = 3.0e38f; w.id[0] = w.id[1] = w.id[2] = -1;
        if (n > RA_MAX_OCC) {                                   // cell list overflow: cooperative brute force
            for (int v = lane; v < nverts; v += 32) knn_insert(w, dist2_ref(q, __ldg(&sv.pos[v])), v);
        } else {
            int skip = -1;
            if (B >= 3.0e38f) {
                // bootstrap: nearest cell of the nearest super cell, scanned cooperatively; the 3rd smallest of the lanes' best
                // distances is an upper bound of the true 3rd distance (three warp minima, no full merge)
                float best = 3.0e38f; int bs = 0;
                for (int s = lane; s < ns; s += 32) {
                    float lb = bbox_dist2(q, __ldg(&sv.sup_lo[s]), __ldg(&sv.sup_hi[s]));
                    if (lb < best) { best = lb; bs = s; }
                }
                bs = warp_argmin(best, bs);
                const int c0 = __float_as_int(__ldg(&sv.sup_lo[bs]).w), c1 = __float_as_int(__ldg(&sv.sup_hi[bs]).w);
                best = 3.0e38f; int bc = c0;
                for (int c = c0 + lane; c < c1; c += 32) {
                    float lb = bbox_dist2(q, __ldg(&sv.occ_lo[c]), __ldg(&sv.occ_hi[c]));
                    if (lb < best) { best = lb; bc = c; }
                }
                bc = warp_argmin(best, bc);
                skip = bc;
                const int s = __float_as_int(__ldg(&sv.occ_lo[bc]).w), e = __float_as_int(__ldg(&sv.occ_hi[bc]).w);
                for (int v = s + lane; v < e; v += 32) { float4 t = __ldg(&sv.pos2[v]); knn_insert(w, dist2_ref(q, t), __float_as_int(t.w)); }
                float mine = w.d2[0];
#pragma unroll
                for (int k = 0; k < 3; k++) {                   // k-th smallest lane minimum (one winner removed per round)
                    B = warp_min_f(mine);
                    const unsigned win = __ballot_sync(0xffffffffu, mine == B);
                    if (lane == __ffs(win) - 1) mine = 3.0e38f;
                }
            }
            for (int s0 = 0; s0 < ns; s0 += 32) {               // super cells: one per lane
                const int si = s0 + lane;
                bool squal = false;
                if (si < ns) squal = bbox_dist2(q, __ldg(&sv.sup_lo[si]), __ldg(&sv.sup_hi[si])) * 0.9999f <= B;      // 1e-4 slack: boxes round independently of dist2_ref
                unsigned sm = __ballot_sync(0xffffffffu, squal);
                while (sm) {
                    const int ss = s0 + __ffs(sm) - 1;
                    sm &= sm - 1;
                    const int c0 = __float_as_int(__ldg(&sv.sup_lo[ss]).w), c1 = __float_as_int(__ldg(&sv.sup_hi[ss]).w);
                    for (int cb = c0; cb < c1; cb += 32) {      // its coarse cells: one per lane
                        const int c = cb + lane;
                        bool qual = false;
                        if (c < c1 && c != skip) qual = bbox_dist2(q, __ldg(&sv.occ_lo[c]), __ldg(&sv.occ_hi[c])) * 0.9999f <= B;
                        unsigned qm = __ballot_sync(0xffffffffu, qual);
                        while (qm) {
                            const int cc = cb + __ffs(qm) - 1;
                            qm &= qm - 1;
                            const int s = __float_as_int(__ldg(&sv.occ_lo[cc]).w), e = __float_as_int(__ldg(&sv.occ_hi[cc]).w);
                            if (lane == 0) { KNN_STAT(2, 1); KNN_STAT(3, e - s); }      // (one count per query, not per lane)
                            for (int v = s + lane; v < e; v += 32) {
                                float4 t = __ldg(&sv.pos2[v]);
                                knn_insert(w, dist2_ref(q, t), __float_as_int(t.w));
                            }
                        }
                    }
                }
            }
        }
        // global top 3 of the lanes' sorted triples (disjoint vertex sets): three rounds of "smallest head wins, the winner pops" --
        // ~90 instructions instead of the ~255 of five shuffle-merge rounds with three insertions each (A/B on one box: no measurable
        // difference in the frame -- the body's far queries are latency-bound, not instruction-bound)
        KnnOut g;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int win = warp_argmin(w.d2[0], lane);
            g.d2[r] = __shfl_sync(0xffffffffu, w.d2[0], win); g.id[r] = __shfl_sync(0xffffffffu, w.id[0], win);
            if (lane == win) { w.d2[0] = w.d2[1]; w.id[0] = w.id[1]; w.d2[1] = w.d2[2]; w.id[1] = w.id[2]; w.d2[2] = 3.0e38f; w.id[2] = -1; }
        }
        if (lane == src) o = g;
    }
}

// ------------------------------------------------------------------------------------------ front-end
struct HdqFront {
    float smpl;       // mean signed vertex distance after the |.| rule (base_network.py:374-375)
    bool in_shell;
    float3 bpts;      // big-pose point (valid if in_shell)
    float Ar[9];      // blended A[:3,:3]       (valid if in_shell and want_mats)
    float Rinv[9];    // inverse_3x3 of it
    float bigAr[9];   // blended big_A[:3,:3]
    float bigRinv[9];
};

__device__ __forceinline__ void inverse3x3_ref(const float* R, float* M) {
    // adjugate / (det + 1e-8)   blend_utils.py:125-165
    float r00 = R[0], r01 = R[1], r02 = R[2], r10 = R[3], r11 = R[4], r12 = R[5], r20 = R[6], r21 = R[7], r22 = R[8];
    float m00 = r11 * r22 - r21 * r12, m10 = -r10 * r22 + r20 * r12, m20 = r10 * r21 - r20 * r11;
    float m01 = -r01 * r22 + r21 * r02, m11 = r00 * r22 - r20 * r02, m21 = -r00 * r21 + r20 * r01;
    float m02 = r01 * r12 - r11 * r02, m12 = -r00 * r12 + r10 * r02, m22 = r00 * r11 - r10 * r01;
    float D = r00 * m00 + r01 * m10 + r02 * m20;
    float inv = D + 1e-8f;
    M[0] = m00 / inv; M[1] = m01 / inv; M[2] = m02 / inv;
    M[3] = m10 / inv; M[4] = m11 / inv; M[5] = m12 / inv;
    M[6] = m20 / inv; M[7] = m21 / inv; M[8] = m22 / inv;
}

// MUST be called by all 32 lanes of a warp (warp-cooperative 3-NN inside); `active` = this lane has a point.
template <bool WANT_MATS>
__device__ void hdq_front(const FrameConst* __restrict__ fc, const SortedVerts& sv, int nverts, float3 x, bool active, float th,
                          float blend_radius, HdqFront& out, int coop_max = 12, bool packet = false) {
    // world -> pose: (x - Th) @ R      blend_utils.py:252-261
    float3 q = make3(x.x - fc->Th[0], x.y - fc->Th[1], x.z - fc->Th[2]);
    float3 p = make3(q.x * fc->R[0] + q.y * fc->R[3] + q.z * fc->R[6],
                     q.x * fc->R[1] + q.y * fc->R[4] + q.z * fc->R[7],
                     q.x * fc->R[2] + q.y * fc->R[5] + q.z * fc->R[8]);
    KnnOut nn;
    knn3_warp(fc, sv, nverts, p, active, nn, coop_max, packet);
    out.in_shell = false;
    out.smpl = 0.f;
    if (!active) return;
    float th2 = th * th;
    float sdfk[3];
    float4 tv0 = __ldg(&sv.tv[nn.id[0]]);
    float d2f[3];
    int idf[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float4 v = __ldg(&sv.pos[nn.id[k]]);
        float4 n = __ldg(&sv.nrm[nn.id[k]]);
        float dt = (p.x - v.x) * n.x + (p.y - v.y) * n.y + (p.z - v.z) * n.z;
        sdfk[k] = sqrtf(nn.d2[k]) * signf(dt);
        d2f[k] = nn.d2[k];
        idf[k] = nn.id[k];
    }
    // geodesic filter: neighbours whose canonical vertex is >= th from the closest one's -> closest  (:148-160)
#pragma unroll
    for (int k = 1; k < 3; k++) {
        float4 t = __ldg(&sv.tv[nn.id[k]]);
        float ex = t.x - tv0.x, ey = t.y - tv0.y, ez = t.z - tv0.z;
        float g2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
        if (!(g2 < th2)) { sdfk[k] = sdfk[0]; d2f[k] = d2f[0]; idf[k] = idf[0]; }
    }
    float smpl = (sdfk[0] + sdfk[1] + sdfk[2]) / 3.0f;
    out.smpl = (smpl < -th) ? smpl : fabsf(smpl);
    out.in_shell = nn.d2[0] < th2;
    if (!out.in_shell) return;
    // gaussian blend of the 3 neighbours' skinning transforms   base_network.py:287-296
    float inv2r2 = 1.0f / (2.0f * blend_radius * blend_radius);
    float w[3], ws = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) { w[k] = expf(-d2f[k] * inv2r2); ws += w[k]; }
    ws += 1.1920929e-07f;   // torch.finfo(float32).eps
    float T[24];
#pragma unroll
    for (int i = 0; i < 24; i++) T[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float wk = w[k] / ws;
        const float4* Tp = reinterpret_cast<const float4*>(sv.T + (size_t)idf[k] * 24);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            float4 t = __ldg(&Tp[i]);
            T[i * 4 + 0] += wk * t.x; T[i * 4 + 1] += wk * t.y; T[i * 4 + 2] += wk * t.z; T[i * 4 + 3] += wk * t.w;
        }
    }
    float Ar[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    float Rinv[9];
    inverse3x3_ref(Ar, Rinv);
    float3 dlt = make3(p.x - T[3], p.y - T[7], p.z - T[11]);
    float3 tp = make3(Rinv[0] * dlt.x + Rinv[1] * dlt.y + Rinv[2] * dlt.z,
                      Rinv[3] * dlt.x + Rinv[4] * dlt.y + Rinv[5] * dlt.z,
                      Rinv[6] * dlt.x + Rinv[7] * dlt.y + Rinv[8] * dlt.z);
    const float* B = T + 12;
    out.bpts = make3(B[0] * tp.x + B[1] * tp.y + B[2] * tp.z + B[3],
                     B[4] * tp.x + B[5] * tp.y + B[6] * tp.z + B[7],
                     B[8] * tp.x + B[9] * tp.y + B[10] * tp.z + B[11]);
    if (WANT_MATS) {
        float Br[9] = {B[0], B[1], B[2], B[4], B[5], B[6], B[8], B[9], B[10]};
#pragma unroll
        for (int i = 0; i < 9; i++) { out.Ar[i] = Ar[i]; out.Rinv[i] = Rinv[i]; out.bigAr[i] = Br[i]; }
        inverse3x3_ref(Br, out.bigRinv);
    }
}

// smooth transition of the network distance with the SMPL distance  (base_network.py:377-381)
__device__ __forceinline__ float hdq_blend(float net, float smpl, float th, bool smooth) {
    if (!smooth) return net;
    float r = clampf(fabsf(net) / th, 0.f, 1.f);
    return smpl * r + net * (1.f - r);
}

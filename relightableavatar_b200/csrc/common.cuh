// Shared declarations for the ra_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>

#define RA_MAX_GRID_DIM 64
#define RA_MAX_CELLS (RA_MAX_GRID_DIM * RA_MAX_GRID_DIM * RA_MAX_GRID_DIM)
#define RA_NLIGHT_MAX 512
#ifndef RA_KNN_RMAX
#define RA_KNN_RMAX 2
#endif
#ifndef RA_TRACE_MINBLOCKS
#define RA_TRACE_MINBLOCKS 3      // resident 256-thread blocks per SM the surface tracing kernel is compiled for (register cap: 80)
#endif
#ifndef RA_SHADOW_MINBLOCKS
#define RA_SHADOW_MINBLOCKS 4     // ... the shadow tracing kernel: 64 registers (32 B of spills) -> 32 instead of 24 resident warps; measured on the
#endif                            // 512^2 frame: visibility stage 19.1 / 18.2 / 17.4 ms at 2 / 3 / 4 blocks (the kernel is L2-latency bound: more warps win)
#ifndef RA_PKT_MIN
#define RA_PKT_MIN 2         // far lanes a warp needs before its far-field 3-NN runs as ONE packet search (hdq.cuh: knn3_packet; measured flat from 2 to 6)
#endif
#ifndef RA_PKT_RHO
#define RA_PKT_RHO 1.0e9f    // largest packet radius (metres) served that way: no limit -- with per-lane box tests a wide packet only scans the union
#endif                       // of its lanes' candidate cells (floor pass at 512^2: 162 / 148 / 134 / 124 / 108 / 106 ms at 0.1 / 0.15 / 0.25 / 0.4 / 1 m / no limit)
#define RA_GRID2_RATIO 3.0f
#define RA_MAX_OCC 8192
#define RA_MAX_SUP 512       // super cells (blocks of coarse cells) of the far-field 3-NN hierarchy
#ifndef RA_NB_LEVELS
#define RA_NB_LEVELS 4       // neighbourhood-list levels: certified search radius up to 4 fine cells (14 cm at 3.5 cm cells >= the 12.5 cm shell)
#endif

// Per-frame constants living in device memory (written by k_frame_prep, read by every kernel).
struct FrameConst {
    float R[9];            // pose->world rotation (batch.R), row-major
    float Th[3];
    float g_org[3];        // uniform grid over the posed vertices (pose space)
    float g_h, g_inv_h;
    int g_dim[3];
    int g_cells;
    float g2_org[3];       // coarse second-level grid for query points far from the body
    float g2_h, g2_inv_h;
    int g2_dim[3];
    int g2_cells;
    int n_occ;             // occupied coarse cells (list in SortedVerts::occ_lo/occ_hi, grouped by super cell)
    int n_sup;             // occupied super cells (SortedVerts::sup_lo/sup_hi)
    float wb[6];           // wbounds (2,3) as given (unpadded)
    float resd_b0[256];    // layer-0 bias + W0[:,63:219] . poses      (cond folded, base_network.py:34-40)
    float resd_b4[256];    // layer-4 bias + W4[:,256+63:475] . poses
    float rend_b3[256];    // render l3 bias + W3[:,256:412] . mat_cond (base_network.py:166-167,501-504)
};

// Per-vertex data sorted by grid cell.
struct SortedVerts {
    float4* pos;    // xyz (pose space), w = original vertex index (int bits)
    float4* nrm;    // pnorm xyz
    float4* tv;     // big-pose vertex xyz
    float* T;       // [N][24]: A_v (3x4 row-major) then bigA_v (3x4): sum_j weights[v][j] * A_j
    int* cell_start;  // [cells+1]
    float4* pos2;     // vertices in coarse-cell order: xyz, w = index into pos/nrm/tv/T (int bits)
    int* cell_start2; // [cells2+1]
    float4* occ_lo;   // occupied coarse cells: tight bbox min xyz, w = first vertex (int bits) in pos2
    float4* occ_hi;   //                        tight bbox max xyz, w = end vertex (int bits)
    float4* sup_lo;   // occupied super cells (blocks of coarse cells): bbox min xyz, w = first entry (int bits) of occ_lo/occ_hi
    float4* sup_hi;   //                                                  bbox max xyz, w = end entry
    float4* occ_tmp;  // [2 * RA_MAX_OCC] scratch of the per-frame build
    int* occ_sup;     // [RA_MAX_OCC] scratch: super cell of every occupied coarse cell
    // per-cell neighbourhood lists (rebuilt per frame): level 0 = the 3x3x3 block around the cell, level 1 / 2 = the cube
    // shells of radius 2 / 3.  nb_pos entries: xyz, w = index into pos/nrm/tv/T (int bits).  A query scans ONE contiguous
    // list per level instead of walking grid rows -- same candidates, no per-row control flow (warp divergence).
    unsigned char* nb_mask;          // [cells] bit lv set: the cell's level-lv list holds vertices (a far query skips the empty levels: 0 = straight to the far phase)
    int* nb_start[RA_NB_LEVELS];     // [cells+1] each
    float4* nb_pos[RA_NB_LEVELS];
};

struct KnnOut {
    float d2[3];
    int id[3];      // sorted-vertex ids
};

__device__ __forceinline__ float3 make3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return make3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float signf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
// reference normalize(): x / (||x|| + 1e-8)   net_utils.py:1626-1628
__device__ __forceinline__ float3 normalize_ref(float3 v) {
    float n = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z) + 1e-8f;
    return make3(v.x / n, v.y / n, v.z / n);
}
// F.normalize(x, eps=1e-7): x / max(||x||, 1e-7)   relight_utils.py:534-536
__device__ __forceinline__ float3 normalize_f(float3 v) {
    float n = fmaxf(sqrtf(v.x * v.x + v.y * v.y + v.z * v.z), 1e-7f);
    return make3(v.x / n, v.y / n, v.z / n);
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// torch.nn.Softplus(beta=100, threshold=20)
__device__ __forceinline__ float softplus100(float z) {
    float t = 100.f * z;
    return (t > 20.f) ? z : log1pf(expf(t)) * 0.01f;
}
// d softplus / dz expressed from the activation value a = softplus100(z): sigmoid(100 z) = 1 - exp(-100 a)
__device__ __forceinline__ float dsoftplus100_from_act(float a) { return 1.f - expf(-100.f * a); }

// warp-aggregated append: returns the slot for lanes with pred, -1 otherwise
__device__ __forceinline__ int warp_append(int* counter, bool pred) {
    unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return -1;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}

// fp32 CUDA-core MLP path: per-layer tiled SGEMM with fused bias/activation (forward) and fused
// activation-derivative (input-gradient backward), plus the small element-wise kernels around them.
// This is the reference-precision path (RA_PRECISION_FP32) and the surface-attribute path (forward +
// analytic d sdf / d bpts, base_network.py:456-475 via autograd in the reference).
#pragma once
#include "common.cuh"

enum { EPI_NONE = 0, EPI_RELU = 1, EPI_SOFTPLUS = 2, EPI_MUL_DRELU = 3, EPI_MUL_DSOFTPLUS = 4 };

struct GemmArgs {
    const float* X; int ldx;      // (M, K) row-major, K % 8 == 0, rows readable up to M
    const float* W; int ldw;      // (N, K) row-major (torch Linear layout), zero padded to K
    const float* bias;            // (N) or null
    float* Y; int ldy;            // (M, N) written at column offset 0 of Y
    const float* aux; int ldaux;  // activation values for the derivative epilogues
    const int* count;             // device row count (total); rows [row0, row0 + M) processed
    int row0, rows_cap;
    int N, K;
};

#define GBM 128
#define GBN 128
#define GBK 8

template <int EPI>
__global__ void __launch_bounds__(256) k_gemm(GemmArgs a) {
    int total = *a.count;
    int M = min(total - a.row0, a.rows_cap);
    int m0 = blockIdx.x * GBM;
    if (m0 >= M) return;
    int n0 = blockIdx.y * GBN;
    __shared__ __align__(16) float As[2][GBK][GBM + 4];
    __shared__ __align__(16) float Bs[2][GBK][GBN + 4];
    int tid = threadIdx.x;
    int tx = tid & 15, ty = tid >> 4;
    int lrow = tid >> 1, lk = (tid & 1) * 4;
    const float* Xp = a.X + (size_t)(m0 + lrow) * a.ldx + lk;
    const float* Wp = a.W + (size_t)(n0 + lrow) * a.ldw + lk;
    bool xok = (m0 + lrow) < M, wok = (n0 + lrow) < a.N;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
    float4 xa = xok ? *reinterpret_cast<const float4*>(Xp) : make_float4(0, 0, 0, 0);
    float4 wa = wok ? __ldg(reinterpret_cast<const float4*>(Wp)) : make_float4(0, 0, 0, 0);
    As[0][lk + 0][lrow] = xa.x; As[0][lk + 1][lrow] = xa.y; As[0][lk + 2][lrow] = xa.z; As[0][lk + 3][lrow] = xa.w;
    Bs[0][lk + 0][lrow] = wa.x; Bs[0][lk + 1][lrow] = wa.y; Bs[0][lk + 2][lrow] = wa.z; Bs[0][lk + 3][lrow] = wa.w;
    __syncthreads();
    int nk = a.K / GBK;
    for (int kt = 0; kt < nk; kt++) {
        int cur = kt & 1;
        if (kt + 1 < nk) {
            xa = xok ? *reinterpret_cast<const float4*>(Xp + (kt + 1) * GBK) : make_float4(0, 0, 0, 0);
            wa = wok ? __ldg(reinterpret_cast<const float4*>(Wp + (kt + 1) * GBK)) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < GBK; k++) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            int nx = cur ^ 1;
            As[nx][lk + 0][lrow] = xa.x; As[nx][lk + 1][lrow] = xa.y; As[nx][lk + 2][lrow] = xa.z; As[nx][lk + 3][lrow] = xa.w;
            Bs[nx][lk + 0][lrow] = wa.x; Bs[nx][lk + 1][lrow] = wa.y; Bs[nx][lk + 2][lrow] = wa.z; Bs[nx][lk + 3][lrow] = wa.w;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
        size_t grow = (size_t)m;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= a.N) continue;
            float v = acc[i][j];
            if (a.bias) v += __ldg(&a.bias[n]);
            if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
            if (EPI == EPI_SOFTPLUS) v = softplus100(v);
            if (EPI == EPI_MUL_DRELU) v = (a.aux[grow * a.ldaux + n] > 0.f) ? v : 0.f;
            if (EPI == EPI_MUL_DSOFTPLUS) v *= dsoftplus100_from_act(a.aux[grow * a.ldaux + n]);
            a.Y[grow * a.ldy + n] = v;
        }
    }
}
// All pointers are CHUNK-LOCAL: the host offsets per-point arrays by row0; row0 only bounds M.

// ---- same GEMM on the (legacy mma.sync) tensor-core path with error-compensated 3xTF32 ---------------------------------
// a = a_hi + a_lo with a_hi = tf32(a), a_lo = tf32(a - a_hi) (same for b);  a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32
// accumulate: fp32-grade results (rel. error ~1e-6) at ~3x the CUDA-core SGEMM rate.  Used for the surface-attribute pass,
// whose outputs (normals, albedo, roughness) are the pixel and therefore stay at fp32 accuracy.
__device__ __forceinline__ uint32_t f2tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#define GTP 136      // padded tile width: 136 % 32 == 8 -> conflict-free fragment loads

template <int EPI>
__global__ void __launch_bounds__(256) k_gemm_tf32x3(GemmArgs a) {
    int total = *a.count;
    int M = min(total - a.row0, a.rows_cap);
    int m0 = blockIdx.x * GBM;
    if (m0 >= M) return;
    int n0 = blockIdx.y * GBN;
    __shared__ uint32_t Ah[2][GBK][GTP], Al[2][GBK][GTP], Bh[2][GBK][GTP], Bl[2][GBK][GTP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int lrow = tid >> 1, lk = (tid & 1) * 4;
    const float* Xp = a.X + (size_t)(m0 + lrow) * a.ldx + lk;
    const float* Wp = a.W + (size_t)(n0 + lrow) * a.ldw + lk;
    const bool xok = (m0 + lrow) < M, wok = (n0 + lrow) < a.N;
    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[i][j][r] = 0.f;
    auto stage = [&](int buf, float4 xa, float4 wa) {
        const float xv[4] = {xa.x, xa.y, xa.z, xa.w}, wv[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t h = f2tf32(xv[q]);
            Ah[buf][lk + q][lrow] = h; Al[buf][lk + q][lrow] = f2tf32(xv[q] - __uint_as_float(h));
            uint32_t hb = f2tf32(wv[q]);
            Bh[buf][lk + q][lrow] = hb; Bl[buf][lk + q][lrow] = f2tf32(wv[q] - __uint_as_float(hb));
        }
    };
    float4 xa = xok ? *reinterpret_cast<const float4*>(Xp) : make_float4(0, 0, 0, 0);
    float4 wa = wok ? __ldg(reinterpret_cast<const float4*>(Wp)) : make_float4(0, 0, 0, 0);
    stage(0, xa, wa);
    __syncthreads();
    const int nk = a.K / GBK;
    for (int kt = 0; kt < nk; kt++) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            xa = xok ? *reinterpret_cast<const float4*>(Xp + (kt + 1) * GBK) : make_float4(0, 0, 0, 0);
            wa = wok ? __ldg(reinterpret_cast<const float4*>(Wp + (kt + 1) * GBK)) : make_float4(0, 0, 0, 0);
        }
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = wn + j * 8 + g;
            bh[j][0] = Bh[cur][t][n]; bh[j][1] = Bh[cur][t + 4][n];
            bl[j][0] = Bl[cur][t][n]; bl[j][1] = Bl[cur][t + 4][n];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int m = wm + i * 16 + g;
            uint32_t ah[4] = {Ah[cur][t][m], Ah[cur][t][m + 8], Ah[cur][t + 4][m], Ah[cur][t + 4][m + 8]};
            uint32_t al[4] = {Al[cur][t][m], Al[cur][t][m + 8], Al[cur][t + 4][m], Al[cur][t + 4][m + 8]};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                mma_tf32(acc[i][j], al, bh[j]);
                mma_tf32(acc[i][j], ah, bl[j]);
                mma_tf32(acc[i][j], ah, bh[j]);
            }
        }
        if (kt + 1 < nk) stage(cur ^ 1, xa, wa);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int m = m0 + wm + i * 16 + g + ((r & 2) ? 8 : 0);
                const int n = n0 + wn + j * 8 + 2 * t + (r & 1);
                if (m >= M || n >= a.N) continue;
                float v = acc[i][j][r];
                if (a.bias) v += __ldg(&a.bias[n]);
                if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
                if (EPI == EPI_SOFTPLUS) v = softplus100(v);
                if (EPI == EPI_MUL_DRELU) v = (a.aux[(size_t)m * a.ldaux + n] > 0.f) ? v : 0.f;
                if (EPI == EPI_MUL_DSOFTPLUS) v *= dsoftplus100_from_act(a.aux[(size_t)m * a.ldaux + n]);
                a.Y[(size_t)m * a.ldy + n] = v;
            }
}


// ---- pipelined variant: cp.async multi-stage ring of raw fp32 tiles, hi/lo split at fragment-load time ------------------
// The kernel above stages 8 K-columns per __syncthreads and converts on the store side; with two CTAs per SM it is bound by
// that load -> convert -> store -> barrier -> load chain (~107 TFLOP/s of tf32 MMAs).  Here 16 K-columns per stage travel
// global -> shared with cp.async (zero-filled beyond M / N), 4 stages deep, one barrier per stage; fragments are split into
// tf32 hi / lo in registers.  Same products in the same order (lo*hi, hi*lo, hi*hi per k8 step, k ascending): bit-identical
// results.  Row stride 20 floats: 16 B-aligned rows and conflict-free fragment reads (g * 20 mod 32 = {0,20,8,28,16,4,24,12}).
#define G2K 16
#define G2LD 20
#define G2STAGES 4
#define G2_SMEM_BYTES (G2STAGES * 2 * GBM * G2LD * 4)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(256, 2) k_gemm_tf32x3_p(GemmArgs a) {
    int total = *a.count;
    int M = min(total - a.row0, a.rows_cap);
    int m0 = blockIdx.x * GBM;
    if (m0 >= M) return;
    int n0 = blockIdx.y * GBN;
    extern __shared__ __align__(16) float g2s[];                 // [stage][X | W][128][G2LD]
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(g2s);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int lrow = tid >> 1, lc = (tid & 1) * 8;              // this thread copies floats [lc, lc + 8) of row lrow of both tiles
    const bool xok = (m0 + lrow) < M, wok = (n0 + lrow) < a.N;
    const float* Xp = a.X + (size_t)(xok ? m0 + lrow : 0) * a.ldx + lc;
    const float* Wp = a.W + (size_t)(wok ? n0 + lrow : 0) * a.ldw + lc;
    const uint32_t dX = (uint32_t)((lrow * G2LD + lc) * 4), dW = dX + GBM * G2LD * 4;
    const int nk = a.K / G2K;
    auto issue = [&](int kt) {
        const uint32_t sb = s_base + (uint32_t)(kt % G2STAGES) * (2 * GBM * G2LD * 4);
        cp_async16(sb + dX, Xp + kt * G2K, xok ? 16 : 0);
        cp_async16(sb + dX + 16, Xp + kt * G2K + 4, xok ? 16 : 0);
        cp_async16(sb + dW, Wp + kt * G2K, wok ? 16 : 0);
        cp_async16(sb + dW + 16, Wp + kt * G2K + 4, wok ? 16 : 0);
    };
    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[i][j][r] = 0.f;
#pragma unroll
    for (int s = 0; s < G2STAGES - 1; s++) {
        if (s < nk) issue(s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < nk; kt++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(G2STAGES - 2) : "memory");
        __syncthreads();                                        // stage kt landed for everyone; stage kt-1 is free again
        if (kt + G2STAGES - 1 < nk) issue(kt + G2STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float* Xs = g2s + (size_t)(kt % G2STAGES) * (2 * GBM * G2LD);
        const float* Ws = Xs + GBM * G2LD;
#pragma unroll
        for (int ks = 0; ks < G2K / 8; ks++) {
            const int kb = ks * 8 + t;
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float* wr = Ws + (wn + j * 8 + g) * G2LD + kb;
                const float b0 = wr[0], b1 = wr[4];
                bh[j][0] = f2tf32(b0); bl[j][0] = f2tf32(b0 - __uint_as_float(bh[j][0]));
                bh[j][1] = f2tf32(b1); bl[j][1] = f2tf32(b1 - __uint_as_float(bh[j][1]));
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float* xr = Xs + (wm + i * 16 + g) * G2LD + kb;
                const float av[4] = {xr[0], xr[8 * G2LD], xr[4], xr[8 * G2LD + 4]};
                uint32_t ah[4], al[4];
#pragma unroll
                for (int q = 0; q < 4; q++) { ah[q] = f2tf32(av[q]); al[q] = f2tf32(av[q] - __uint_as_float(ah[q])); }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    mma_tf32(acc[i][j], al, bh[j]);
                    mma_tf32(acc[i][j], ah, bl[j]);
                    mma_tf32(acc[i][j], ah, bh[j]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int m = m0 + wm + i * 16 + g + ((r & 2) ? 8 : 0);
                const int n = n0 + wn + j * 8 + 2 * t + (r & 1);
                if (m >= M || n >= a.N) continue;
                float v = acc[i][j][r];
                if (a.bias) v += __ldg(&a.bias[n]);
                if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
                if (EPI == EPI_SOFTPLUS) v = softplus100(v);
                if (EPI == EPI_MUL_DRELU) v = (a.aux[(size_t)m * a.ldaux + n] > 0.f) ? v : 0.f;
                if (EPI == EPI_MUL_DSOFTPLUS) v *= dsoftplus100_from_act(a.aux[(size_t)m * a.ldaux + n]);
                a.Y[(size_t)m * a.ldy + n] = v;
            }
}


// Y[m, n] for n < N <= 4: one warp per row.
__global__ void k_skinny(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                         const float* __restrict__ bias, float* Y, int ldy, const int* count, int row0, int rows_cap,
                         int N, int K) {
    int total = *count;
    int M = min(total - row0, rows_cap);
    if (M <= 0) return;
    int warps_per_block = blockDim.x >> 5;
    int lane = threadIdx.x & 31;
    for (int m = blockIdx.x * warps_per_block + (threadIdx.x >> 5); m < M; m += gridDim.x * warps_per_block) {
        const float* x = X + (size_t)m * ldx;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane; k < K; k += 32) {
            float xv = x[k];
#pragma unroll
            for (int n = 0; n < 4; n++)
                if (n < N) s[n] = fmaf(xv, __ldg(&W[(size_t)n * ldw + k]), s[n]);
        }
#pragma unroll
        for (int n = 0; n < 4; n++)
            for (int o = 16; o; o >>= 1) s[n] += __shfl_xor_sync(0xffffffffu, s[n], o);
        if (lane == 0)
            for (int n = 0; n < N; n++) Y[(size_t)m * ldy + n] = s[n] + (bias ? bias[n] : 0.f);
    }
}

// D[m, k] = (sum_{n<N} U[m, n] * W[n, k]) * dact(aux[m, k]);  U == null means N == 1 and U = 1 (start of the
// SDF backward: d a_7 = W8[0, :]).  N <= 4.
template <int EPI>
__global__ void k_outer_small(const float* __restrict__ U, int ldu, const float* __restrict__ W, int ldw, int N, int K,
                              const float* __restrict__ aux, int ldaux, float* D, int ldd, const int* count, int row0,
                              int rows_cap) {
    int total = *count;
    int M = min(total - row0, rows_cap);
    if (M <= 0) return;
    const unsigned n_el = (unsigned)M * (unsigned)K;     // M <= 262 144 rows per chunk, K <= 512: 32-bit index arithmetic
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
        int m = (int)(i / (unsigned)K), k = (int)(i % (unsigned)K);
        size_t g = (size_t)m;
        float v = 0.f;
        if (U == nullptr) v = __ldg(&W[k]);
        else
            for (int n = 0; n < N; n++) v = fmaf(U[g * ldu + n], __ldg(&W[(size_t)n * ldw + k]), v);
        float a = aux[g * ldaux + k];
        if (EPI == EPI_MUL_DRELU) v = (a > 0.f) ? v : 0.f;
        if (EPI == EPI_MUL_DSOFTPLUS) v *= dsoftplus100_from_act(a);
        D[g * ldd + k] = v;
    }
}

// positional encoding of x (3) into dst[0 .. 3+6L), zero padding up to `width`   embedder.py:26-37
__device__ __forceinline__ void pe_write(float3 x, int L, float* dst, int width) {
    dst[0] = x.x; dst[1] = x.y; dst[2] = x.z;
    float f = 1.f;
    for (int l = 0; l < L; l++) {
        float* d = dst + 3 + l * 6;
        d[0] = sinf(x.x * f); d[1] = sinf(x.y * f); d[2] = sinf(x.z * f);
        d[3] = cosf(x.x * f); d[4] = cosf(x.y * f); d[5] = cosf(x.z * f);
        f *= 2.f;
    }
    for (int k = 3 + 6 * L; k < width; k++) dst[k] = 0.f;
}

// X0[m, 0:64] = PE_L(p[m]);  optionally the same into a second buffer at a column offset (skip input)
// One thread per (row, feature): the 64-wide rows leave as coalesced segments (a thread per row wrote 63 scattered words
// through a local-memory buffer: 49 us for the frame's 31 k rows).  Feature values are the same sinf / cosf calls as pe_write.
__device__ __forceinline__ float pe_feature_ref(const float* __restrict__ x3, int L, int k) {
    if (k < 3) return x3[k];
    if (k >= 3 + 6 * L) return 0.f;
    const int j = k - 3, l = j / 6, r = j % 6;
    const float a = x3[r % 3] * (float)(1 << l);
    return (r < 3) ? sinf(a) : cosf(a);
}
__global__ void k_encode(const float* __restrict__ pts, int L, float* X0, int ld0, int w0, float* X1, int ld1, int off1,
                         int w1, const int* count, int row0, int rows_cap) {
    int total = *count;
    int M = min(total - row0, rows_cap);
    if (M <= 0) return;
    const unsigned n_el = (unsigned)M * 64u;              // M <= 262 144 rows per chunk
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
        const unsigned m = i >> 6, k = i & 63u;
        const float v = pe_feature_ref(pts + (size_t)m * 3, L, (int)k);
        if ((int)k < w0) X0[(size_t)m * ld0 + k] = v;
        if (X1 && (int)k < w1) X1[(size_t)m * ld1 + off1 + k] = v;
    }
}

// resd = resd_limit * tanh(z8);  cpts = bpts + resd   (base_network.py:41,466)
__global__ void k_resd_finish(const float* __restrict__ z8, int ldz, const float* __restrict__ bpts, float limit,
                              float* resd, float* cpts, const int* count, int row0, int rows_cap) {
    int total = *count;
    int M = min(total - row0, rows_cap);
    if (M <= 0) return;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        size_t g = (size_t)m;
        for (int c = 0; c < 3; c++) {
            float r = tanhf(z8[g * ldz + c]) * limit;
            resd[g * 3 + c] = r;
            cpts[g * 3 + c] = bpts[g * 3 + c] + r;
        }
    }
}

// Gradient of a PE input: g_x = d[0:3] + sum_l 2^l (cos(2^l x) d_sin[l] - sin(2^l x) d_cos[l]),
// where d = dA (cols offA..) + dB (cols offB..)
__device__ __forceinline__ float3 pe_backward(float3 x, int L, const float* dA, const float* dB) {
    float g[3];
    float xs[3] = {x.x, x.y, x.z};
    for (int c = 0; c < 3; c++) g[c] = dA[c] + (dB ? dB[c] : 0.f);
    float f = 1.f;
    for (int l = 0; l < L; l++) {
        for (int c = 0; c < 3; c++) {
            float ds = dA[3 + l * 6 + c] + (dB ? dB[3 + l * 6 + c] : 0.f);
            float dc = dA[3 + l * 6 + 3 + c] + (dB ? dB[3 + l * 6 + 3 + c] : 0.f);
            float s, co;
            sincosf(xs[c] * f, &s, &co);
            g[c] += f * (co * ds - s * dc);
        }
        f *= 2.f;
    }
    return make3(g[0], g[1], g[2]);
}

// g_cp = PE8-backward(cp; d_pe(layer0) + d_pe(skip));  u = g_cp * limit * (1 - tanh(z8)^2)  (upstream of the resd MLP)
__global__ void k_sdf_grad_to_cp(const float* __restrict__ cpts, const float* __restrict__ dA, int lda,
                                 const float* __restrict__ dB, int ldb, int offB, const float* __restrict__ z8, int ldz,
                                 float limit, float* gcp, float* u, int ldu, const int* count, int row0, int rows_cap) {
    int total = *count;
    int M = min(total - row0, rows_cap);
    if (M <= 0) return;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        size_t g = (size_t)m;
        float3 x = make3(cpts[g * 3], cpts[g * 3 + 1], cpts[g * 3 + 2]);
        float3 gr = pe_backward(x, 8, dA + g * lda, dB + g * ldb + offB);
        gcp[g * 3] = gr.x; gcp[g * 3 + 1] = gr.y; gcp[g * 3 + 2] = gr.z;
        float gv[3] = {gr.x, gr.y, gr.z};
        for (int c = 0; c < 3; c++) {
            float t = tanhf(z8[g * ldz + c]);
            u[g * ldu + c] = gv[c] * limit * (1.f - t * t);
        }
    }
}

// g_bp = g_cp + PE10-backward(bp; d_pe(layer0) + d_pe(skip))
__global__ void k_resd_grad_to_bp(const float* __restrict__ bpts, const float* __restrict__ dA, int lda,
                                  const float* __restrict__ dB, int ldb, int offB, const float* __restrict__ gcp,
                                  float* gbp, const int* count, int row0, int rows_cap) {
    int total = *count;
    int M = min(total - row0, rows_cap);
    if (M <= 0) return;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        size_t g = (size_t)m;
        float3 x = make3(bpts[g * 3], bpts[g * 3 + 1], bpts[g * 3 + 2]);
        float3 gr = pe_backward(x, 10, dA + g * lda, dB + g * ldb + offB);
        gbp[g * 3] = gcp[g * 3] + gr.x;
        gbp[g * 3 + 1] = gcp[g * 3 + 1] + gr.y;
        gbp[g * 3 + 2] = gcp[g * 3 + 2] + gr.z;
    }
}

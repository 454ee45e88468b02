// Fused distance-query MLP on the 5th-gen tensor cores (tcgen05 / TMEM / TMA bulk copies), sm_100a.
//
// One launch evaluates, for every in-shell query point of the work list, the residual-deformation MLP
// (base_network.py:34-42) and the SDF MLP (net_utils.py:1337-1352, output 0 only):
//     bpts -> PE10 -> 9 linears (ReLU, skip@4) -> 0.05*tanh -> cpts -> PE8 -> 9 linears (Softplus100, skip@4) -> sdf
// 18 dependent GEMMs per 128-point tile with the activations never leaving the SM:
//   * operands fp16, accumulation fp32 in TMEM (tcgen05.mma kind::f16, M=128, N<=256, K=16 per instruction);
//   * A (activations): shared memory, canonical K-major no-swizzle layout [K/8][128 rows][8 halves];
//     written by the epilogue warps of the previous layer straight from TMEM (tcgen05.ld);
//   * B (weights): streamed from L2 by the TMA engine (cp.async.bulk + mbarrier complete_tx) in 32-wide K chunks
//     through a 2-stage ring, pre-packed on the host into the exact shared-memory image;
//   * the 156-d pose condition is folded into the biases of layers 0 and 4 once per frame;
//     the SDF skip input cat([h3, PE8])/sqrt(2) is realised by a column permutation of W4.
// Two CTAs per SM (114 KB smem, 256 TMEM columns each): while one CTA's epilogue warps run bias/activation
// out of TMEM, the other CTA's MMAs own the tensor pipe.  Warp roles per CTA: 0 = TMA producer, 1 = MMA
// issuer (+ TMEM allocator), 2..9 = epilogue (warp%4 = TMEM lane quarter, two column halves).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <string>
#include <vector>
#include <cmath>
#include "common.cuh"
#include "../../include/ra_b200.h"

// per-layer clock64 timeline of CTA 0 (tools/tc_timeline.py): compile with -DRA_TC_TIMELINE
#ifdef RA_TC_TIMELINE
#define TC_TL(x) x
#else
#define TC_TL(x)
#endif

#define TC_LAYERS 18
#define TC_TILE_M 128
#define TC_KCHUNK 32
#define TC_STAGES 2
#define TC_THREADS 320
#define TC_ACT_BYTES (128 * 256 * 2)
#define TC_PE_BYTES (128 * 64 * 2)
#define TC_STAGE_BYTES (TC_KCHUNK * 256 * 2)
#define TC_SMEM_BYTES (TC_ACT_BYTES + TC_PE_BYTES + TC_STAGES * TC_STAGE_BYTES + 512)

enum { TC_EPI_RELU = 0, TC_EPI_RESD_FINAL = 1, TC_EPI_SOFTPLUS = 2, TC_EPI_S3 = 3, TC_EPI_SDF_FINAL = 4 };

struct TcLayer {
    int N;            // accumulator columns (multiple of 16)
    int nchunks;      // K / 32
    int pe_from;      // chunks >= pe_from read A from the PE buffer (chunk - pe_from), earlier ones from ACT
    int epi;
    unsigned goff;    // byte offset of the first chunk image in the weight blob
    unsigned boff;    // byte offset of the bias chunk: [2][N][8] halves, k=0: fp16(b), k=1: fp16(b - fp16(b)), rest 0
};

struct TcParams {
    TcLayer layer[TC_LAYERS];
    const float* bias[TC_LAYERS];
    const unsigned char* blob;
    const float* bpts;
    float* out;
    const int* count;
    float resd_limit;
    unsigned long long* dbg;   // optional timeline of CTA 0, tile iteration 1 (debug builds of tools/tc_timeline.py)
};

struct TcWeights {
    unsigned char* blob = nullptr;
    float* bias = nullptr;        // [18][256]
    TcParams p{};
    bool ready = false;
    unsigned long long* dbg = nullptr;
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x4000;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B; SBO = stride between 8-row groups, LBO = stride between the
// two K core matrices of one K=16 step (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_f16(int N) {
    // c_format F32 (bit 4), a/b F16 (0), K-major both, n_dim = N>>3 @17, m_dim = 128>>4 @24
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
// ---- half2 epilogue math -------------------------------------------------------------------------------
// softplus_100(z) = max(z, 0) + g(|z|),  g(a) = log1p(exp(-100 a)) / 100  in (0, 0.00693], evaluated in half2.
// Two evaluations of g (RA_SP_MODE selects; pairs = adjacent accumulator columns):
//   exp  : e = 2^(-144.27 a) by the native ex2.approx.f16x2 (two MUFU.EX2.F16 + a PRMT), g ~ e (c1 + c2 e + c3 e^2).
//          MUFU-bound: 2 MUFU per pair at 16 / clk / SM = 2048 clk per 128 x 256 tile, more than the tile's MMAs (2176).
//   poly : w = max(1 - a / 0.075, 0), g ~ w (c1 + w (c2 + w (c3 + w (c4 + w c5)))): 6 HFMA2, no MUFU.
//   mixed: even pairs exp, odd pairs poly -- balances the xu and fma pipes.
// Emulated end to end on the fitted SDF net (fp16 operands / activations, fp32 accumulate, 40 k points near the surface):
// mean |sdf error| 3.7e-5 with an exact softplus, 8.8e-5 exp, 5.3e-5 poly, 6.8e-5 mixed -- all at the fp16 operand floor.
#ifndef RA_SP_MODE
#define RA_SP_MODE 1      // measured (k_mlp_tc6 softplus-layer epilogue, clk per tile): exp 2600, mixed 2250, poly 2030
#endif
__device__ __forceinline__ __half2 h2_sp_exp(const __half2 z) {
    const __half2 t = __hmul2(__habs2(z), __float2half2_rn(-144.269504f));
    uint32_t eu;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(eu) : "r"(*reinterpret_cast<const uint32_t*>(&t)));
    const __half2 e = *reinterpret_cast<const __half2*>(&eu);
    __half2 p = __hfma2(e, __float2half2_rn(0.0011465454f), __float2half2_rn(-0.0040842847f));
    p = __hfma2(p, e, __float2half2_rn(0.0098745818f));
    return __hfma2(p, e, __hmax2(z, __float2half2_rn(0.f)));
}
__device__ __forceinline__ __half2 h2_sp_poly(const __half2 z) {
    const __half2 w = __hfma2_relu(__habs2(z), __float2half2_rn(-1.f / 0.075f), __float2half2_rn(1.f));      // max(fma, 0) in one HFMA2.RELU
    __half2 p = __hfma2(w, __float2half2_rn(0.01282501220703125f), __float2half2_rn(-0.006336212158203125f));
    p = __hfma2(p, w, __float2half2_rn(-0.0011892318725585938f));
    p = __hfma2(p, w, __float2half2_rn(0.001796722412109375f));
    p = __hfma2(p, w, __float2half2_rn(-0.00015354156494140625f));
    return __hfma2(p, w, __hmax2(z, __float2half2_rn(0.f)));
}
// `odd`: parity of the pair index (column / 2) -- compile-time at every call site
__device__ __forceinline__ uint32_t h2_softplus100(float a, float b, bool odd) {
    const __half2 z = __floats2half2_rn(a, b);
    const __half2 r = (RA_SP_MODE == 1 || (RA_SP_MODE == 2 && odd)) ? h2_sp_poly(z) : h2_sp_exp(z);
    return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t h2_relu(float a, float b) {
    uint32_t r;                                   // round-to-nearest conversion and ReLU in one instruction (same values as cvt + max)
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// write 8 consecutive K values (cols k0..k0+7, k0 % 8 == 0) of this thread's row into a K-major buffer
__device__ __forceinline__ void put8(uint32_t buf, int row, int k0, const float* v) {
    st_shared_v4(buf + (uint32_t)(k0 >> 3) * 2048u + (uint32_t)row * 16u, pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]),
                 pack_h2(v[6], v[7]));
}

// Positional encoding [x, sin(2^l x), cos(2^l x)]_l (embedder.py:26-37) of this thread's point into chunks [ch0, ch1)
// of a 64-wide K-major buffer.  Every feature is evaluated directly: the argument 2^l x (exact in fp32) is reduced to
// [-pi, pi] with a two-constant Cody-Waite step (|error| < 3e-7 for |2^l x| < 2^10) and goes through the MUFU sine / cosine
// (|error| < 1e-6 on that interval) -- ~7 instructions per (axis, octave) and far below the fp16 rounding (2.4e-4) the value
// receives as a tensor-core operand.  (The first version anchored a double-angle recurrence on accurate sincosf calls: ~2000 clk
// per 128 x 64 encode, every column-group warp recomputing all octaves -- the largest non-MMA item of the kernel's timeline.)
__device__ __forceinline__ float pe_reduce(float t) {
    const float k = rintf(t * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, t);
    return fmaf(k, 1.7484556000744883e-7f, r);
}
// feature `idx` (compile-time after unrolling) of the L-octave encoding of x; idx >= 3 + 6 L: zero padding
template <int L>
__device__ __forceinline__ float pe_feat(const float3& x, int idx) {
    if (idx < 3) return idx == 0 ? x.x : (idx == 1 ? x.y : x.z);
    if (idx >= 3 + 6 * L) return 0.f;
    const int j = idx - 3, l = j / 6, r = j % 6, c = r % 3;
    const float t = pe_reduce((c == 0 ? x.x : (c == 1 ? x.y : x.z)) * (float)(1 << l));
    return (r < 3) ? __sinf(t) : __cosf(t);
}
template <int L, int CH>
__device__ __forceinline__ void write_pe_chunk(uint32_t buf, int row, const float3& x) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = pe_feat<L>(x, CH * 8 + j);
    put8(buf, row, CH * 8, v);
}
// chunks [ch0, ch1) with ch0 = 2 cg (four column-group warps, two chunks each) or ch0 = 4 half (two warps, four chunks each)
template <int L>
__device__ __forceinline__ void write_pe(uint32_t buf, int row, float3 x, int ch0, int ch1) {
    switch (ch0) {
    case 0: write_pe_chunk<L, 0>(buf, row, x); write_pe_chunk<L, 1>(buf, row, x); if (ch1 > 2) { write_pe_chunk<L, 2>(buf, row, x); write_pe_chunk<L, 3>(buf, row, x); } break;
    case 2: write_pe_chunk<L, 2>(buf, row, x); write_pe_chunk<L, 3>(buf, row, x); break;
    case 4: write_pe_chunk<L, 4>(buf, row, x); write_pe_chunk<L, 5>(buf, row, x); if (ch1 > 6) { write_pe_chunk<L, 6>(buf, row, x); write_pe_chunk<L, 7>(buf, row, x); } break;
    default: write_pe_chunk<L, 6>(buf, row, x); write_pe_chunk<L, 7>(buf, row, x); break;
    }
}
// one PE8 feature by index (the three features that land in a mixed chunk of the S4 skip input)
__device__ __forceinline__ float pe_feature(float3 x, int idx) { return pe_feat<8>(x, idx); }

// hidden-layer epilogue: 128 accumulator columns of this warp's rows -> activation -> fp16 A operand of the next layer.
// Biases are already inside the accumulator (added by a K=16 rank-2 MMA step, see the MMA issuer).
template <bool SOFTPLUS>
__device__ __forceinline__ void epi_hidden(uint32_t t_lane, uint32_t s_act, int row, int half) {
    uint32_t ra[32], rb[32];
    const int cbase = half * 128;
    tmem_ld32(t_lane + (uint32_t)cbase, ra);
    tmem_ld_wait();
#pragma unroll
    for (int cb = 0; cb < 4; cb++) {
        uint32_t* cur = (cb & 1) ? rb : ra;
        uint32_t* nxt = (cb & 1) ? ra : rb;
        if (cb < 3) tmem_ld32(t_lane + (uint32_t)(cbase + (cb + 1) * 32), nxt);     // in flight while we work on `cur`
        const int c0 = cbase + cb * 32;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t h[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float a = __uint_as_float(cur[g * 8 + 2 * j]), b = __uint_as_float(cur[g * 8 + 2 * j + 1]);
                h[j] = SOFTPLUS ? h2_softplus100(a, b, j & 1) : h2_relu(a, b);
            }
            st_shared_v4(s_act + (uint32_t)((c0 >> 3) + g) * 2048u + (uint32_t)row * 16u, h[0], h[1], h[2], h[3]);
        }
        if (cb < 3) tmem_ld_wait();
    }
}

// ------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(TC_THREADS, 2) k_mlp_tc(const __grid_constant__ TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_act = s_base;
    const uint32_t s_pe = s_base + TC_ACT_BYTES;
    const uint32_t s_w = s_pe + TC_PE_BYTES;
    const uint32_t s_bar = s_w + TC_STAGES * TC_STAGE_BYTES;   // barriers: full[2], empty[2], act_ready, acc_ready ; tmem ptr
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 16, bar_act = s_bar + 32, bar_acc = s_bar + 40;
    const uint32_t s_ones = s_bar + 128;    // 256 B: core matrix of rows [1,1,0,0,0,0,0,0], then a zero core matrix
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (s_bar - s_base) + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int count = *P.count;
    const int n_tiles = (count + TC_TILE_M - 1) / TC_TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_act, 8);
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 64) {     // "ones" A operand of the bias step: every row reads (1, 1, 0, ..., 0) over K = 16
        uint32_t v = (threadIdx.x < 32 && (threadIdx.x & 3) == 0) ? 0x3C003C00u : 0u;     // half2(1, 1) in the first word of each 16 B row
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_ones + threadIdx.x * 4), "r"(v) : "memory");
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_bar + 64), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: stream every layer's weight chunks (+ its bias chunk), once per tile =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < TC_LAYERS; l++) {
                    const uint32_t bytes = (uint32_t)P.layer[l].N * TC_KCHUNK * 2;
                    const unsigned char* src = P.blob + P.layer[l].goff;
                    const int nch = P.layer[l].nchunks;
                    for (int c = 0; c <= nch; c++, it++) {
                        uint32_t s = it & 1, ph = (it >> 1) & 1;
                        const uint32_t nb = (c < nch) ? bytes : bytes / 2;
                        const unsigned char* g = (c < nch) ? src + (size_t)c * bytes : P.blob + P.layer[l].boff;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, nb);
                        tma_bulk_g2s(s_w + s * TC_STAGE_BYTES, g, nb, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, lc = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < TC_LAYERS; l++, lc++) {
                    const int N = P.layer[l].N;
                    const uint32_t idesc = make_idesc_f16(N);
                    const uint32_t lbo_b = (uint32_t)N * 16u;
                    const int nch = P.layer[l].nchunks;
                    TC_TL(const bool rec = P.dbg && blockIdx.x == 0 && tile == (int)gridDim.x;)
                    TC_TL(if (rec) P.dbg[l * 8 + 0] = clock64();)
                    mbar_wait(bar_act, lc & 1);
                    tc_fence_after();
                    TC_TL(if (rec) P.dbg[l * 8 + 1] = clock64();)
                    for (int c = 0; c <= nch; c++, it++) {
                        uint32_t s = it & 1, ph = (it >> 1) & 1;
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        uint32_t b_base = s_w + s * TC_STAGE_BYTES;
                        if (c < nch) {
                            uint32_t a_base = (c >= P.layer[l].pe_from) ? (s_pe + (uint32_t)(c - P.layer[l].pe_from) * 4u * 2048u)
                                                                         : (s_act + (uint32_t)c * 4u * 2048u);
#pragma unroll
                            for (int kk = 0; kk < TC_KCHUNK / 16; kk++) {
                                uint64_t ad = make_sdesc(a_base + (uint32_t)kk * 2u * 2048u, 2048u, 128u);
                                uint64_t bd = make_sdesc(b_base + (uint32_t)kk * 2u * lbo_b, lbo_b, 128u);
                                umma_f16(tmem, ad, bd, idesc, (c | kk) ? 1u : 0u);
                            }
                        } else {
                            // bias step: D += ones(128 x 16) * [b_hi; b_lo; 0...]  -- A rows all alias one core matrix (SBO = 0)
                            uint64_t ad = make_sdesc(s_ones, 128u, 0u);
                            uint64_t bd = make_sdesc(b_base, lbo_b, 128u);
                            umma_f16(tmem, ad, bd, idesc, 1u);
                        }
                        umma_commit(bar_empty + 8 * s);      // frees the weight stage when these MMAs retire
                    }
                    umma_commit(bar_acc);                     // accumulator of this layer complete
                    TC_TL(if (rec) P.dbg[l * 8 + 2] = clock64();)
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int e = warp - 2;
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = e >> 2;                // column half: 0 -> cols [0,128), 1 -> [128,256)
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t lc = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int gidx = tile * TC_TILE_M + row;
            float3 bp = make3(0.f, 0.f, 0.f);
            if (gidx < count) bp = make3(P.bpts[(size_t)gidx * 3], P.bpts[(size_t)gidx * 3 + 1], P.bpts[(size_t)gidx * 3 + 2]);
            float3 cp = bp;
            // prologue: PE10(bp) -> PE buffer (63 features, padded to 64); the two column-half warps split the chunks
            write_pe<10>(s_pe, row, bp, half * 4, half * 4 + 4);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_act);
#pragma unroll 1
            for (int l = 0; l < TC_LAYERS; l++, lc++) {
                const int epi = P.layer[l].epi;
                TC_TL(const bool rec = P.dbg && blockIdx.x == 0 && tile == (int)gridDim.x && warp == 2 && lane == 0;)
                TC_TL(if (rec) P.dbg[l * 8 + 3] = clock64();)
                mbar_wait(bar_acc, lc & 1);
                tc_fence_after();
                TC_TL(if (rec) P.dbg[l * 8 + 4] = clock64();)
                if (epi == TC_EPI_RELU) {
                    epi_hidden<false>(t_lane, s_act, row, half);
                } else if (epi == TC_EPI_SOFTPLUS) {
                    epi_hidden<true>(t_lane, s_act, row, half);
                } else if (epi == TC_EPI_S3) {
                    // S3: 205 outputs -> ACT cols [48, 253); PE8(cp) features 0..47 -> cols [0,48), 48..50 -> cols 253..255
                    // half 0: accumulator cols [0,104); half 1: [104,208) + the tail
                    const int a0 = half ? 104 : 0;
#pragma unroll 1
                    for (int cb = 0; cb < 13; cb++) {            // 13 groups of 8 accumulator columns
                        const int c0 = a0 + cb * 8;
                        uint32_t r[16];
                        tmem_ld16(t_lane + (uint32_t)(c0 & ~15), r);   // 16-col aligned load, pick the 8 we need
                        tmem_ld_wait();
                        const int o = c0 & 15;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = __uint_as_float(o ? r[8 + j] : r[j]);
                        uint32_t h[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) h[j] = h2_softplus100(v[2 * j], v[2 * j + 1], j & 1);
                        if (c0 == 200) {      // cols 248..255: outputs 200..204 then PE8 features 48,49,50
                            float p48 = pe_feature(cp, 48), p49 = pe_feature(cp, 49), p50 = pe_feature(cp, 50);
                            h[2] = (h[2] & 0x0000FFFFu) | (pack_h2(0.f, p48) & 0xFFFF0000u);
                            h[3] = pack_h2(p49, p50);
                        }
                        st_shared_v4(s_act + (uint32_t)((48 + c0) >> 3) * 2048u + (uint32_t)row * 16u, h[0], h[1], h[2], h[3]);
                    }
                    // PE8 features 0..47 (6 chunks) copied from the PE buffer: 3 chunks per half
                    for (int ch = half * 3; ch < half * 3 + 3; ch++) {
                        uint32_t a, b, c, d;
                        ld_shared_v4(s_pe + (uint32_t)ch * 2048u + (uint32_t)row * 16u, a, b, c, d);
                        st_shared_v4(s_act + (uint32_t)ch * 2048u + (uint32_t)row * 16u, a, b, c, d);
                    }
                } else if (epi == TC_EPI_RESD_FINAL) {
                    uint32_t r[16];
                    tmem_ld16(t_lane, r);
                    tmem_ld_wait();
                    float rx = tanhf(__uint_as_float(r[0])) * P.resd_limit;
                    float ry = tanhf(__uint_as_float(r[1])) * P.resd_limit;
                    float rz = tanhf(__uint_as_float(r[2])) * P.resd_limit;
                    cp = make3(bp.x + rx, bp.y + ry, bp.z + rz);
                    // PE8(cp): 51 features padded to 64 -> PE buffer (input of S0, later copied into the S4 skip columns)
                    write_pe<8>(s_pe, row, cp, half * 4, half * 4 + 4);
                } else {   // TC_EPI_SDF_FINAL
                    if (half == 0) {
                        uint32_t r[16];
                        tmem_ld16(t_lane, r);
                        tmem_ld_wait();
                        if (gidx < count) P.out[gidx] = __uint_as_float(r[0]);
                    }
                }
                TC_TL(if (rec) P.dbg[l * 8 + 5] = clock64();)
                if (l + 1 < TC_LAYERS) {
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_act);
                }
                TC_TL(if (rec) P.dbg[l * 8 + 6] = clock64();)
            }
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// per frame: the pose-folded biases of residual layers 0 and 4 -> their fp16 hi/lo bias chunks in the weight blob
__global__ void k_tc_pack_bias(const float* __restrict__ b0, const float* __restrict__ b4, __half* dst0, __half* dst4) {
    int n = threadIdx.x;     // 256 threads
    for (int w = 0; w < 2; w++) {
        const float b = (w ? b4 : b0)[n];
        __half* d = (w ? dst4 : dst0);
        __half hi = __float2half_rn(b);
        __half lo = __float2half_rn(b - __half2float(hi));
        d[n * 8 + 0] = hi; d[n * 8 + 1] = lo;
        for (int j = 2; j < 8; j++) d[n * 8 + j] = __float2half_rn(0.f);
        for (int j = 0; j < 8; j++) d[(256 + n) * 8 + j] = __float2half_rn(0.f);
    }
}

// ------------------------------------------------------------------------------------------ host side
static int tc_init(TcWeights& t, std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_mlp_tc): ") + cudaGetErrorString(e); return 1; }
    return 0;
}

static void tc_free(TcWeights& t) {
    if (t.blob) cudaFree(t.blob);
    if (t.bias) cudaFree(t.bias);
    t.blob = nullptr; t.bias = nullptr;
}

// Pack one layer: src (N_src, K_src) fp32 row-major; colmap[k] = source column for packed column k (-1 -> 0); rows >= N_src -> 0.
static void tc_pack_layer(std::vector<__half>& blob, const std::vector<float>& w, int N_src, int K_src, const std::vector<int>& colmap,
                          int N_pad, float scale, int row0 = 0) {
    int K = (int)colmap.size();
    int nch = K / TC_KCHUNK;
    size_t base = blob.size();
    blob.resize(base + (size_t)nch * TC_KCHUNK * N_pad, __float2half(0.f));
    for (int c = 0; c < nch; c++)
        for (int kq = 0; kq < TC_KCHUNK / 8; kq++)
            for (int n = 0; n < N_pad; n++)
                for (int j = 0; j < 8; j++) {
                    int k = c * TC_KCHUNK + kq * 8 + j;
                    float v = 0.f;
                    int sn = row0 + n;
                    if (n < N_src && colmap[k] >= 0) v = w[(size_t)sn * K_src + colmap[k]] * scale;
                    blob[base + (size_t)c * TC_KCHUNK * N_pad + ((size_t)kq * N_pad + n) * 8 + j] = __float2half_rn(v);
                }
}

// bias chunk image [2][N_pad][8] halves: k = 0 holds fp16(b), k = 1 holds fp16(b - fp16(b)); multiplied by the all-ones A step
static void tc_pack_bias(std::vector<__half>& blob, const std::vector<float>& b, int N_src, int N_pad) {
    size_t base = blob.size();
    blob.resize(base + (size_t)2 * N_pad * 8, __float2half(0.f));
    for (int n = 0; n < N_src; n++) {
        __half hi = __float2half_rn(b[n]);
        blob[base + (size_t)n * 8 + 0] = hi;
        blob[base + (size_t)n * 8 + 1] = __float2half_rn(b[n] - __half2float(hi));
    }
}

static int tc_upload(TcWeights& t, const ra_weights* w, int cond, std::string& err, cudaStream_t st) {
    auto fetch = [&](const float* src, size_t n, std::vector<float>& dst) -> bool {
        dst.resize(n);
        return cudaMemcpyAsync(dst.data(), src, n * sizeof(float), cudaMemcpyDefault, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    };
    const int rK[9] = {63 + cond, 256, 256, 256, 256 + 63 + cond, 256, 256, 256, 256};      // cond = pose condition width (3 * n_bones)
    static const int sN[9] = {256, 256, 256, 205, 256, 256, 256, 256, 257};
    static const int sK[9] = {51, 256, 256, 256, 256, 256, 256, 256, 256};
    std::vector<__half> blob;
    std::vector<float> bias((size_t)TC_LAYERS * 256, 0.f);
    TcParams& P = t.p;
    auto ident = [](int K_used, int K_pad) { std::vector<int> m(K_pad, -1); for (int k = 0; k < K_used; k++) m[k] = k; return m; };
    for (int l = 0; l < 9; l++) {           // residual MLP
        std::vector<float> hw, hb;
        int N = (l == 8) ? 3 : 256;
        if (!fetch(w->resd_w[l], (size_t)N * rK[l], hw) || !fetch(w->resd_b[l], N, hb)) { err = "tc_upload: copy failed"; return 1; }
        std::vector<int> cm = (l == 0) ? ident(63, 64) : (l == 4 ? ident(319, 320) : ident(256, 256));
        int Np = (l == 8) ? 16 : 256;
        P.layer[l].N = Np; P.layer[l].nchunks = (int)cm.size() / TC_KCHUNK; P.layer[l].goff = (unsigned)(blob.size() * 2);
        P.layer[l].pe_from = (l == 0) ? 0 : (l == 4 ? 8 : 1 << 20);
        P.layer[l].epi = (l == 8) ? TC_EPI_RESD_FINAL : TC_EPI_RELU;
        tc_pack_layer(blob, hw, N, rK[l], cm, Np, 1.f);
        for (int n = 0; n < N; n++) bias[(size_t)l * 256 + n] = hb[n];
        P.layer[l].boff = (unsigned)(blob.size() * 2);
        tc_pack_bias(blob, hb, N, Np);
    }
    const float rs2 = (float)(1.0 / std::sqrt(2.0));
    for (int l = 0; l < 9; l++) {           // SDF MLP
        std::vector<float> hw, hb;
        if (!fetch(w->sdf_w[l], (size_t)sN[l] * sK[l], hw) || !fetch(w->sdf_b[l], sN[l], hb)) { err = "tc_upload: copy failed"; return 1; }
        int L = 9 + l;
        std::vector<int> cm;
        int Nsrc = sN[l], Np = 256;
        float scale = 1.f;
        if (l == 0) cm = ident(51, 64);
        else if (l == 4) {                  // packed col j: [PE 0..47 | h3 0..204 | PE 48..50]; source cols: h3 at 0..204, PE at 205..255
            cm.assign(256, -1);
            for (int j = 0; j < 48; j++) cm[j] = 205 + j;
            for (int j = 0; j < 205; j++) cm[48 + j] = j;
            for (int j = 0; j < 3; j++) cm[253 + j] = 205 + 48 + j;
            scale = rs2;
        } else cm = ident(256, 256);
        if (l == 3) Np = 208;
        if (l == 8) { Nsrc = 1; Np = 16; }  // only row 0 (the sdf) is needed for a distance query
        P.layer[L].N = Np; P.layer[L].nchunks = (int)cm.size() / TC_KCHUNK; P.layer[L].goff = (unsigned)(blob.size() * 2);
        P.layer[L].pe_from = (l == 0) ? 0 : 1 << 20;
        P.layer[L].epi = (l == 8) ? TC_EPI_SDF_FINAL : (l == 3 ? TC_EPI_S3 : TC_EPI_SOFTPLUS);
        tc_pack_layer(blob, hw, Nsrc, sK[l], cm, Np, scale);
        for (int n = 0; n < std::min(Nsrc, 256); n++) bias[(size_t)L * 256 + n] = hb[n];
        P.layer[L].boff = (unsigned)(blob.size() * 2);
        tc_pack_bias(blob, hb, std::min(Nsrc, 256), Np);
    }
    tc_free(t);
    if (cudaMalloc((void**)&t.blob, blob.size() * 2) != cudaSuccess || cudaMalloc((void**)&t.bias, bias.size() * 4) != cudaSuccess) {
        err = "tc_upload: cudaMalloc failed"; return 1;
    }
    cudaMemcpy(t.blob, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(t.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    for (int l = 0; l < TC_LAYERS; l++) P.bias[l] = t.bias + (size_t)l * 256;
    P.blob = t.blob;
    t.ready = true;
    return 0;
}

// per frame: layers 0 and 4 of the residual MLP take their (pose-folded) biases from the frame constants
static void tc_set_frame(TcWeights& t, const FrameConst* fc, cudaStream_t st, int64_t& launches) {
    t.p.bias[0] = &fc->resd_b0[0];
    t.p.bias[4] = &fc->resd_b4[0];
    k_tc_pack_bias<<<1, 256, 0, st>>>(&fc->resd_b0[0], &fc->resd_b4[0], reinterpret_cast<__half*>(t.blob + t.p.layer[0].boff),
                                      reinterpret_cast<__half*>(t.blob + t.p.layer[4].boff));
    launches++;
}

static void tc_distance(TcWeights& t, const float* bpts, float* out, const int* count, float resd_limit, int sms, cudaStream_t st,
                        int64_t& launches) {
    TcParams p = t.p;
    p.bpts = bpts; p.out = out; p.count = count; p.resd_limit = resd_limit; p.dbg = t.dbg;
    k_mlp_tc<<<2 * sms, TC_THREADS, TC_SMEM_BYTES, st>>>(p);
    launches++;
}

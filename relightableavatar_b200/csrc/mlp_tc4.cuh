// Single-CTA-per-SM variant of the fused MLP kernel: TWO 128-point tiles ping-pong on the tensor pipe (tile A's MMAs
// run while tile B's epilogue warps convert its accumulator and vice versa) and share ONE 64 KB weight ring.
// Rationale (timeline of k_mlp_tc, tools/tc_timeline.py): with two independent CTAs per SM each MMA phase lasts
// ~5100 cycles instead of 2176 because a 32 KB ring cannot cover the ~1200-cycle L2->smem latency at the 62 B/clk a
// 256-wide layer needs (32 KB / 1200 clk = 27 B/clk).  Here the tile that is in its MMA phase owns the whole 64 KB.
// smem: 2 x (64 KB activations + 16 KB PE) + 4 x 16 KB ring = 224 KB; TMEM: 2 x 256 columns; 18 warps.
#pragma once
#include "mlp_tc.cuh"

#define TC4_STAGES 4
#define TC4_THREADS 576
#define TC4_SMEM_BYTES (2 * TC_ACT_BYTES + 2 * TC_PE_BYTES + TC4_STAGES * TC_STAGE_BYTES + 512)

// ------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(TC4_THREADS, 1) k_mlp_tc4(const __grid_constant__ TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    // two tiles (t = 0, 1) ping-pong on one tensor pipe and share ONE weight ring of 4 x 16 KB
    const uint32_t s_act0 = s_base;                                   // ACT[t] = s_act0 + t * TC_ACT_BYTES
    const uint32_t s_pe0 = s_base + 2 * TC_ACT_BYTES;                 // PE[t]  = s_pe0 + t * TC_PE_BYTES
    const uint32_t s_w = s_pe0 + 2 * TC_PE_BYTES;
    const uint32_t s_bar = s_w + TC4_STAGES * TC_STAGE_BYTES;         // barriers: full[4], empty[4], act_ready[2], acc_ready[2]; tmem ptr
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_act0 = s_bar + 64, bar_acc0 = s_bar + 80;
    const uint32_t s_ones = s_bar + 128;    // 256 B: core matrix of rows [1,1,0,0,0,0,0,0], then a zero core matrix
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (s_bar - s_base) + 96);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int count = *P.count;
    const int n_tiles = (count + TC_TILE_M - 1) / TC_TILE_M;
    const int n_pairs = (n_tiles + 1) / 2;     // this CTA works on tile pairs (2j, 2j+1)

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC4_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int t = 0; t < 2; t++) { mbar_init(bar_act0 + 8 * t, 8); mbar_init(bar_acc0 + 8 * t, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 64) {     // "ones" A operand of the bias step: every row reads (1, 1, 0, ..., 0) over K = 16
        uint32_t v = (threadIdx.x < 32 && (threadIdx.x & 3) == 0) ? 0x3C003C00u : 0u;     // half2(1, 1) in the first word of each 16 B row
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_ones + threadIdx.x * 4), "r"(v) : "memory");
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_bar + 96), "r"(512) : "memory");
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
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int l = 0; l < TC_LAYERS; l++) {
                    const uint32_t bytes = (uint32_t)P.layer[l].N * TC_KCHUNK * 2;
                    const unsigned char* src = P.blob + P.layer[l].goff;
                    const int nch = P.layer[l].nchunks;
                    for (int t = 0; t < 2; t++)                              // consumed by tile 0's MMAs, then tile 1's
                        for (int c = 0; c <= nch; c++, it++) {
                            uint32_t s = it & (TC4_STAGES - 1), ph = (it / TC4_STAGES) & 1;
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
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int l = 0; l < TC_LAYERS; l++, lc++) {
                    const int N = P.layer[l].N;
                    const uint32_t idesc = make_idesc_f16(N);
                    const uint32_t lbo_b = (uint32_t)N * 16u;
                    const int nch = P.layer[l].nchunks;
                    for (int t = 0; t < 2; t++) {
                        const uint32_t s_act = s_act0 + t * TC_ACT_BYTES, s_pe = s_pe0 + t * TC_PE_BYTES;
                        const uint32_t tmem_d = tmem + (uint32_t)t * 256u;
                        mbar_wait(bar_act0 + 8 * t, lc & 1);
                        tc_fence_after();
                        for (int c = 0; c <= nch; c++, it++) {
                            uint32_t s = it & (TC4_STAGES - 1), ph = (it / TC4_STAGES) & 1;
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
                                    umma_f16(tmem_d, ad, bd, idesc, (c | kk) ? 1u : 0u);
                                }
                            } else {
                                uint64_t ad = make_sdesc(s_ones, 128u, 0u);
                                uint64_t bd = make_sdesc(b_base, lbo_b, 128u);
                                umma_f16(tmem_d, ad, bd, idesc, 1u);
                            }
                            umma_commit(bar_empty + 8 * s);
                        }
                        umma_commit(bar_acc0 + 8 * t);
                    }
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int e = (warp - 2) & 7;
        const int tt = (warp - 2) >> 3;         // which tile of the pair this warp serves
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = e >> 2;                // column half: 0 -> cols [0,128), 1 -> [128,256)
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)tt * 256u;
        const uint32_t s_act = s_act0 + tt * TC_ACT_BYTES, s_pe = s_pe0 + tt * TC_PE_BYTES;
        const uint32_t bar_act = bar_act0 + 8 * tt, bar_acc = bar_acc0 + 8 * tt;
        uint32_t lc = 0;
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const int tile = pair * 2 + tt;
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
                mbar_wait(bar_acc, lc & 1);
                tc_fence_after();
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
                if (l + 1 < TC_LAYERS) {
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_act);
                }
            }
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}


static int tc4_init(std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc4, cudaFuncAttributeMaxDynamicSharedMemorySize, TC4_SMEM_BYTES);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_mlp_tc4): ") + cudaGetErrorString(e); return 1; }
    return 0;
}
static void tc4_distance(TcWeights& t, const float* bpts, float* out, const int* count, float resd_limit, int sms, cudaStream_t st,
                         int64_t& launches) {
    TcParams p = t.p;
    p.bpts = bpts; p.out = out; p.count = count; p.resd_limit = resd_limit; p.dbg = nullptr;
    k_mlp_tc4<<<sms, TC4_THREADS, TC4_SMEM_BYTES, st>>>(p);
    launches++;
}

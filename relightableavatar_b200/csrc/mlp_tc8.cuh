// k_mlp_tc8: the CTA-pair kernel of mlp_tc6.cuh with the A operand of every hidden layer in TENSOR MEMORY.
//
// Why (profiles/r02_k_mlp_tc6_timeline.txt, r02_attempts_second_half.txt): in k_mlp_tc6 the MMAs' A + B operand reads (64 B/clk), the
// epilogue's 64 KB of A stores per layer and slot (30 B/clk) and the TMA fill of the weight ring (30 B/clk) add up to ~124 of the
// 128 B/clk of shared memory; with the epilogue's stores switched off (wrong results, timing only) the kernel runs 7 % faster.
// Here the activations never touch shared memory: the epilogue writes them as packed fp16 pairs with tcgen05.st into 128 tensor-memory
// columns of the slot and the MMAs take A from there (`tcgen05.mma [d], [a], b-desc`).  Shared memory carries the weight ring and the
// 64-wide positional-encoding inputs only.
//
// Tensor memory has 512 columns and a 256-wide layer needs 256 fp32 accumulator columns + 128 columns of fp16 A per tile, so the layer's
// N is processed as TWO HALVES with a 128-column accumulator:  slot p = columns [256 p, 256 p + 128) D, [256 p + 128, 256 p + 256) A.
//   MMA issue order per layer : (slot 0, half 0) (slot 1, half 0) (slot 0, half 1) (slot 1, half 1)
//   epilogue (16 warps)       : the same order.  Half 0's outputs are converted and PARKED in 16 registers per thread (A is still being
//                               read by half 1's MMAs); after half 1 both halves are stored to A and the next layer may start.
//   barriers                  : full[8] / empty[8] (ring of 8 KB stages: 64 K-columns of one half), acc_ready[2] (per slot, every half),
//                               d_free[2] (half 0's accumulator has been read: half 1 may overwrite it), act_ready[2] (A / PE of the next
//                               layer complete)
// Weight image: per (layer, half, rank) the rows [base_h + rank * n_h / 2, + n_h / 2) as K-major 32-column chunk images, then the bias
// chunk -- accumulator column j of half h is output column base_h + j, so the next layer's K index is simply base_h + j.
// Same fp16 operands, same K order, same fp32 accumulation as k_mlp_tc / k_mlp_tc6: distances are bit-identical
// (tests/test_gpu_parity.py::test_two_cta_kernel_variant_matches_single_cta).
#pragma once
#include "mlp_tc7.cuh"

#define TC8_STAGES 8
#define TC8_STAGE_BYTES (64 * 64 * 2)        // 64 K-columns x 64 rows (n_h / 2) of fp16
#define TC8_SMEM_BYTES (2 * TC_PE_BYTES + TC8_STAGES * TC8_STAGE_BYTES + 1024)

struct Tc8Layer {
    int nh;                 // halves (2; 1 for the N = 16 output layers)
    int n_h[2];             // accumulator columns of each half (multiples of 16)
    int base[2];            // first output column of each half
    int nch, n_act;         // 64-wide K chunks in total / of them from tensor memory (the rest: positional encoding from shared memory)
    int epi;
    unsigned goff[2][2];    // [half][rank] byte offset of the first chunk image
    unsigned boff[2][2];    // [half][rank] bias chunk [2][n_h / 2][8]
};
struct Tc8Params {
    Tc8Layer layer[TC_LAYERS];
    const unsigned char* blob;
    const float* bpts;
    float* out;
    const int* count;
    float resd_limit;
    unsigned long long* dbg;   // optional timeline (-DRA_TC_TIMELINE builds, tools/tc8_timeline.py)
};
struct Tc8Weights {
    unsigned char* blob = nullptr;
    Tc8Params p{};
    bool ready = false;
    unsigned long long* dbg = nullptr;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// 16 accumulator columns -> activation -> 8 packed fp16 pairs
template <bool SOFTPLUS>
__device__ __forceinline__ void tc8_act16(const uint32_t* r, uint32_t* h) {
#pragma unroll
    for (int j = 0; j < 8; j++)
        h[j] = SOFTPLUS ? h2_softplus100(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), j & 1) : h2_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
}
// this warp's 32 accumulator columns at `t_src`, 16 at a time (register budget: two slots park 16 registers each) -> 16 packed pairs
template <bool SOFTPLUS>
__device__ __forceinline__ void tc8_load_act32(uint32_t t_src, uint32_t* h) {
    uint32_t r[16];
    tmem_ld16(t_src, r);
    tmem_ld_wait();
    tc8_act16<SOFTPLUS>(r, h);
    tmem_ld16(t_src + 16u, r);
    tmem_ld_wait();
    tc8_act16<SOFTPLUS>(r, h + 8);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC6_THREADS, 1) k_mlp_tc8(const __grid_constant__ Tc8Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_pe0 = s_base;                                    // PE[p] = s_pe0 + p * TC_PE_BYTES
    const uint32_t s_w = s_pe0 + 2 * TC_PE_BYTES;
    const uint32_t s_bar = s_w + TC8_STAGES * TC8_STAGE_BYTES;
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 64, bar_act0 = s_bar + 128, bar_acc0 = s_bar + 144, bar_dfree0 = s_bar + 160;
    const uint32_t s_tmem = s_bar + 192;
    const uint32_t s_ones = s_bar + 256;    // 256 B: core matrix of rows [1,1,0,0,0,0,0,0], then a zero core matrix
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - s_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int count = *P.count;
    if (count == 0) return;
    const int n_tiles = (count + TC_TILE_M - 1) / TC_TILE_M;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int np = (n_tiles <= 2 * n_clusters) ? 1 : 2;           // short work lists: slot 0 only (see mlp_tc6.cuh)
    const int n_quads = (np == 2) ? (n_tiles + 3) / 4 : (n_tiles + 1) / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC8_STAGES; s++) { mbar_init(bar_full + 8 * s, rank == 0 ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int p = 0; p < 2; p++) { mbar_init(bar_act0 + 8 * p, 2 * TC6_EPI_WARPS); mbar_init(bar_acc0 + 8 * p, 1); mbar_init(bar_dfree0 + 8 * p, 2 * TC6_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 64) {     // "ones" A operand of the bias step: every row reads (1, 1, 0, ..., 0) over K = 16
        uint32_t v = (threadIdx.x < 32 && (threadIdx.x & 3) == 0) ? 0x3C003C00u : 0u;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_ones + threadIdx.x * 4), "r"(v) : "memory");
        fence_async_smem();
    }
    if (warp == TC6_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == TC6_WARP_TMA) {
        // ===================== TMA producer: this CTA's rows of every (layer, half) chunk, once per slot =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int quad = cluster_id; quad < n_quads; quad += n_clusters)
                for (int l = 0; l < TC_LAYERS; l++) {
                    const Tc8Layer& Ly = P.layer[l];
                    for (int h = 0; h < Ly.nh; h++) {
                        const uint32_t bytes = (uint32_t)(Ly.n_h[h] / 2) * 64u * 2u;      // one 64-wide chunk of this half
                        const unsigned char* src = P.blob + Ly.goff[h][rank];
                        const unsigned char* bsrc = P.blob + Ly.boff[h][rank];
                        for (int p = 0; p < np; p++)
                            for (int c = 0; c <= Ly.nch; c++, it++) {
                                const uint32_t s = it & (TC8_STAGES - 1), ph = (it / TC8_STAGES) & 1;
                                const uint32_t nb = (c < Ly.nch) ? bytes : bytes / 4;           // bias chunk: 16 K-columns
                                const unsigned char* g = (c < Ly.nch) ? src + (size_t)c * bytes : bsrc;
                                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                                mbar_expect_tx(bar_full + 8 * s, nb);
                                tma_bulk_g2s(s_w + s * TC8_STAGE_BYTES, g, nb, bar_full + 8 * s);
                            }
                    }
                }
        }
    } else if (warp == TC6_WARP_MMA) {
        if (lane == 0 && rank == 1) {
            // ===================== peer: tell the leader when my rows of each chunk have landed =====================
            uint32_t it = 0;
            for (int quad = cluster_id; quad < n_quads; quad += n_clusters)
                for (int l = 0; l < TC_LAYERS; l++) {
                    const int n = P.layer[l].nh * np * (P.layer[l].nch + 1);
                    for (int c = 0; c < n; c++, it++) {
                        const uint32_t s = it & (TC8_STAGES - 1), ph = (it / TC8_STAGES) & 1;
                        mbar_wait(bar_full + 8 * s, ph);
                        mbar_arrive_remote(bar_full + 8 * s, 0);
                    }
                }
        } else if (rank == 0) {
            // ===================== leader: MMA issuer (warp-uniform, one elected lane) =====================
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            uint32_t it = 0, lc = 0, dfc = 0;
            for (int quad = cluster_id; quad < n_quads; quad += n_clusters) {
                for (int l = 0; l < TC_LAYERS; l++, lc++) {
                    const Tc8Layer& Ly = P.layer[l];
                    for (int h = 0; h < Ly.nh; h++) {
                        const int n_h = Ly.n_h[h];
                        const uint32_t idesc = make_idesc2_f16(n_h);
                        const uint32_t lbo_b = (uint32_t)(n_h / 2) * 16u;
                        for (int p = 0; p < np; p++) {
                            const uint32_t s_pe = s_pe0 + p * TC_PE_BYTES;
                            const uint32_t tmem_d = tmem_u + (uint32_t)p * 256u, tmem_a = tmem_d + 128u;
                            TC_TL(const bool rec = P.dbg && blockIdx.x == 0 && quad == cluster_id + n_clusters && lane == 0;)
                            TC_TL(const int ti = (l * 2 + h) * 2 + p;)
                            TC_TL(if (rec) P.dbg[256 + ti * 3 + 0] = clock64();)
                            if (h == 0) mbar_wait(bar_act0 + 8 * p, lc & 1);             // A / PE of this layer complete (and D read)
                            else mbar_wait(bar_dfree0 + 8 * p, dfc & 1);                 // half 0's accumulator has been read
                            tc_fence_after();
                            TC_TL(if (rec) P.dbg[256 + ti * 3 + 1] = clock64();)
                            for (int c = 0; c <= Ly.nch; c++, it++) {
                                const uint32_t s = it & (TC8_STAGES - 1), ph = (it / TC8_STAGES) & 1;
                                mbar_wait(bar_full + 8 * s, ph);
                                tc_fence_after();
                                const uint32_t b_base = s_w + s * TC8_STAGE_BYTES;
                                if (c < Ly.nch) {
                                    uint64_t bd[4];
#pragma unroll
                                    for (int kk = 0; kk < 4; kk++) bd[kk] = make_sdesc(b_base + (uint32_t)kk * 2u * lbo_b, lbo_b, 128u);
                                    if (c < Ly.n_act) {                                   // A: 32 columns of packed fp16 pairs per 64-wide chunk
                                        const uint32_t a0 = tmem_a + (uint32_t)c * 32u;
                                        if (elect_one()) {
                                            umma2_f16_ts(tmem_d, a0, bd[0], idesc, c ? 1u : 0u);
                                            umma2_f16_ts(tmem_d, a0 + 8u, bd[1], idesc, 1u);
                                            umma2_f16_ts(tmem_d, a0 + 16u, bd[2], idesc, 1u);
                                            umma2_f16_ts(tmem_d, a0 + 24u, bd[3], idesc, 1u);
                                            umma_commit2(bar_empty + 8 * s);
                                        }
                                    } else {                                              // positional encoding from shared memory
                                        const uint32_t a_base = s_pe + (uint32_t)(c - Ly.n_act) * 8u * 2048u;
                                        uint64_t ad[4];
#pragma unroll
                                        for (int kk = 0; kk < 4; kk++) ad[kk] = make_sdesc(a_base + (uint32_t)kk * 2u * 2048u, 2048u, 128u);
                                        if (elect_one()) {
                                            umma2_f16(tmem_d, ad[0], bd[0], idesc, c ? 1u : 0u);
                                            umma2_f16(tmem_d, ad[1], bd[1], idesc, 1u);
                                            umma2_f16(tmem_d, ad[2], bd[2], idesc, 1u);
                                            umma2_f16(tmem_d, ad[3], bd[3], idesc, 1u);
                                            umma_commit2(bar_empty + 8 * s);
                                        }
                                    }
                                } else {
                                    const uint64_t ad = make_sdesc(s_ones, 128u, 0u);
                                    const uint64_t bd = make_sdesc(b_base, lbo_b, 128u);
                                    if (elect_one()) {
                                        umma2_f16(tmem_d, ad, bd, idesc, 1u);
                                        umma_commit2(bar_empty + 8 * s);
                                        umma_commit2(bar_acc0 + 8 * p);       // accumulator of (slot p, layer l, half h) complete, both CTAs
                                    }
                                }
                                __syncwarp();
                            }
                            TC_TL(if (rec) P.dbg[256 + ti * 3 + 2] = clock64();)
                        }
                        if (h == 1) dfc++;
                    }
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int cg = warp >> 2;               // column group: accumulator columns [32 cg, 32 cg + 32) of a half
        const int row = q * 32 + lane;
        const uint32_t t_lane0 = tmem + ((uint32_t)(q * 32) << 16);
        float3 cp[2];                           // deformed point of the slot's row (PE8 features of the S3 skip assembly)
        int gidx[2];
        uint32_t acc_c[2] = {0, 0};
        uint32_t hold0[16], hold1[16];          // half 0's activations of slot 0 / slot 1, parked until half 1 is done
        TC_TL(bool tl_quad = false;)

        auto arrive = [&](uint32_t bar) {        // one arrival per warp on the leader's barrier
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar, 0);
        };
        auto prologue = [&](int quad, int p) {      // PE10(bp) of the slot's new tile -> PE[p]; input of layer R0 ready
            const int tile = (np == 2) ? quad * 4 + p * 2 + (int)rank : quad * 2 + (int)rank;
            gidx[p] = tile * TC_TILE_M + row;
            cp[p] = make3(0.f, 0.f, 0.f);
            if (gidx[p] < count) cp[p] = make3(P.bpts[(size_t)gidx[p] * 3], P.bpts[(size_t)gidx[p] * 3 + 1], P.bpts[(size_t)gidx[p] * 3 + 2]);
            write_pe<10>(s_pe0 + p * TC_PE_BYTES, row, cp[p], cg * 2, cg * 2 + 2);
            fence_async_smem();
            arrive(bar_act0 + 8 * p);
        };
        // one (layer, half, slot) item; `hold`: the slot's parking registers
        auto item = [&](int l, int h, int p, uint32_t* hold, int next_quad) {
            const Tc8Layer& Ly = P.layer[l];
            const int epi = Ly.epi;
            const uint32_t s_pe = s_pe0 + p * TC_PE_BYTES;
            const uint32_t t_d = t_lane0 + (uint32_t)p * 256u, t_a = t_d + 128u;
            TC_TL(const bool rec = P.dbg && blockIdx.x == 0 && tl_quad && warp == 0 && lane == 0;)
            TC_TL(const int ti = (l * 2 + h) * 2 + p;)
            TC_TL(if (rec) P.dbg[ti * 3 + 0] = clock64();)
            mbar_wait(bar_acc0 + 8 * p, acc_c[p] & 1);
            acc_c[p]++;
            tc_fence_after();
            TC_TL(if (rec) P.dbg[ti * 3 + 1] = clock64();)
            const bool last_half = (h == Ly.nh - 1);
            if (epi == TC_EPI_RELU || epi == TC_EPI_SOFTPLUS) {
                if (!last_half) {
                    if (epi == TC_EPI_RELU) tc8_load_act32<false>(t_d + (uint32_t)(cg * 32), hold); else tc8_load_act32<true>(t_d + (uint32_t)(cg * 32), hold);
                    tc_fence_before();
                    arrive(bar_dfree0 + 8 * p);                  // the accumulator is in registers: half 1's MMAs may overwrite it
                    TC_TL(if (rec) P.dbg[ti * 3 + 2] = clock64();)
                    return;
                }
                // all MMAs of the layer have retired (acc_ready of the last half): A may be replaced by the next layer's input
                tmem_st16(t_a + (uint32_t)(16 * cg), hold);                     // outputs [32 cg, +32) of half 0 -> K pairs [16 cg, +16)
                uint32_t cur[16];
                if (epi == TC_EPI_RELU) tc8_load_act32<false>(t_d + (uint32_t)(cg * 32), cur); else tc8_load_act32<true>(t_d + (uint32_t)(cg * 32), cur);
                tmem_st16(t_a + (uint32_t)(64 + 16 * cg), cur);                 // outputs 128 + [32 cg, +32) of half 1
                tmem_st_wait();
            } else if (epi == TC_EPI_S3) {
                // 205 outputs -> K indices [48, 253) of S4's input (pairs 24 ..); PE8 features 0..47 -> K [0, 48), 48..50 -> K 253..255.
                // half 0: outputs [0, 112) (cg 3: 16 columns); half 1: outputs 112 + [0, 96) (cg 0..2; cg 3 copies the PE8 features)
                if (!last_half) {
                    if (cg < 3) tc8_load_act32<true>(t_d + (uint32_t)(cg * 32), hold);
                    else { uint32_t r[16]; tmem_ld16(t_d + 96u, r); tmem_ld_wait(); tc8_act16<true>(r, hold); }
                    tc_fence_before();
                    arrive(bar_dfree0 + 8 * p);
                    return;
                }
                uint32_t cur[16];
                if (cg < 3) {
                    tmem_st16(t_a + (uint32_t)(24 + 16 * cg), hold);            // half 0: outputs [32 cg, +32) -> pairs 24 + [16 cg, +16)
                    tc8_load_act32<true>(t_d + (uint32_t)(cg * 32), cur);
                    if (cg == 2) {            // outputs 176..207: 204 is the last real one; K 253..255 carry PE8 features 48, 49, 50
                        const float p48 = pe_feature(cp[p], 48), p49 = pe_feature(cp[p], 49), p50 = pe_feature(cp[p], 50);
                        cur[14] = (cur[14] & 0x0000FFFFu) | (pack_h2(0.f, p48) & 0xFFFF0000u);
                        cur[15] = pack_h2(p49, p50);
                    }
                    tmem_st16(t_a + (uint32_t)(80 + 16 * cg), cur);             // half 1: outputs 112 + [32 cg, +32) -> pairs 80 + [16 cg, +16)
                } else {
                    tmem_st8(t_a + 72u, hold);                                  // half 0: outputs [96, 112) -> pairs [72, 80)
#pragma unroll
                    for (int ch = 0; ch < 4; ch++) ld_shared_v4(s_pe + (uint32_t)ch * 2048u + (uint32_t)row * 16u, cur[4 * ch], cur[4 * ch + 1], cur[4 * ch + 2], cur[4 * ch + 3]);
                    tmem_st16(t_a, cur);                                        // PE8 features 0..31 -> pairs [0, 16)
#pragma unroll
                    for (int ch = 0; ch < 2; ch++) ld_shared_v4(s_pe + (uint32_t)(4 + ch) * 2048u + (uint32_t)row * 16u, cur[4 * ch], cur[4 * ch + 1], cur[4 * ch + 2], cur[4 * ch + 3]);
                    tmem_st8(t_a + 16u, cur);                                   // PE8 features 32..47 -> pairs [16, 24)
                }
                tmem_st_wait();
            } else if (epi == TC_EPI_RESD_FINAL) {
                uint32_t r[16];
                tmem_ld16(t_d, r);
                tmem_ld_wait();
                float rx = tanhf(__uint_as_float(r[0])) * P.resd_limit;
                float ry = tanhf(__uint_as_float(r[1])) * P.resd_limit;
                float rz = tanhf(__uint_as_float(r[2])) * P.resd_limit;
                cp[p] = make3(cp[p].x + rx, cp[p].y + ry, cp[p].z + rz);          // (cp held the big-pose point until here)
                write_pe<8>(s_pe, row, cp[p], cg * 2, cg * 2 + 2);            // PE8(cp): input of S0, later copied into S4's skip columns
            } else {   // TC_EPI_SDF_FINAL
                if (cg == 0) {
                    uint32_t r[16];
                    tmem_ld16(t_d, r);
                    tmem_ld_wait();
                    if (gidx[p] < count) P.out[gidx[p]] = __uint_as_float(r[0]);
                }
            }
            if (l + 1 < TC_LAYERS) {
                tc_fence_before();
                fence_async_smem();
                arrive(bar_act0 + 8 * p);
                TC_TL(if (rec) P.dbg[ti * 3 + 2] = clock64();)
            } else if (next_quad < n_quads) {
                tc_fence_before();
                prologue(next_quad, p);
            }
        };

        int quad = cluster_id;
        if (quad < n_quads) { prologue(quad, 0); if (np == 2) prologue(quad, 1); }
        for (; quad < n_quads; quad += n_clusters) {
            const int next = quad + n_clusters;
            TC_TL(tl_quad = (quad == cluster_id + n_clusters);)
#pragma unroll 1
            for (int l = 0; l < TC_LAYERS; l++) {
                const int nh = P.layer[l].nh;
#pragma unroll 1
                for (int h = 0; h < nh; h++) {
                    item(l, h, 0, hold0, next);
                    if (np == 2) item(l, h, 1, hold1, next);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == TC6_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// per frame: pose-folded biases of residual layers 0 / 4 -> the fp16 hi/lo bias chunks of (half, rank): [2][64][8] each
__global__ void k_tc8_pack_bias(const float* __restrict__ b0, const float* __restrict__ b4, unsigned char* blob,
                                unsigned o000, unsigned o001, unsigned o010, unsigned o011, unsigned o400, unsigned o401, unsigned o410, unsigned o411) {
    const int n = threadIdx.x;                 // output column 0..255: half n / 128, rank (n % 128) / 64, row n % 64
    const int h = n >> 7, r = (n >> 6) & 1, row = n & 63;
    for (int w = 0; w < 2; w++) {
        const float b = (w ? b4 : b0)[n];
        const unsigned off = w ? (h ? (r ? o411 : o410) : (r ? o401 : o400)) : (h ? (r ? o011 : o010) : (r ? o001 : o000));
        __half* d = reinterpret_cast<__half*>(blob + off);
        const __half hi = __float2half_rn(b);
        d[row * 8 + 0] = hi; d[row * 8 + 1] = __float2half_rn(b - __half2float(hi));
    }
}

static int tc8_init(std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc8, cudaFuncAttributeMaxDynamicSharedMemorySize, TC8_SMEM_BYTES);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_mlp_tc8): ") + cudaGetErrorString(e); return 1; }
    return 0;
}
static void tc8_free(Tc8Weights& t) { if (t.blob) cudaFree(t.blob); t.blob = nullptr; }

// the layer stack of tc2_upload, packed per (layer, half, rank)
static int tc8_upload(Tc8Weights& t, const ra_weights* w, int cond, std::string& err, cudaStream_t st) {
    auto fetch = [&](const float* src, size_t n, std::vector<float>& dst) -> bool {
        dst.resize(n);
        return cudaMemcpyAsync(dst.data(), src, n * sizeof(float), cudaMemcpyDefault, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    };
    const int rK[9] = {63 + cond, 256, 256, 256, 256 + 63 + cond, 256, 256, 256, 256};
    static const int sN[9] = {256, 256, 256, 205, 256, 256, 256, 256, 257};
    static const int sK[9] = {51, 256, 256, 256, 256, 256, 256, 256, 256};
    std::vector<__half> blob;
    Tc8Params& P = t.p;
    auto ident = [](int K_used, int K_pad) { std::vector<int> m(K_pad, -1); for (int k = 0; k < K_used; k++) m[k] = k; return m; };
    // Np: padded output width; K layout `cm` (multiple of 64); n_act64: 64-wide chunks taken from tensor memory
    auto pack = [&](int L, const std::vector<float>& hw, const std::vector<float>& hb, int N_src, int K_src, const std::vector<int>& cm,
                    int Np, float scale, int n_act64, int epi) {
        Tc8Layer& Ly = P.layer[L];
        Ly.nch = (int)cm.size() / 64; Ly.n_act = n_act64; Ly.epi = epi;
        if (Np == 16) { Ly.nh = 1; Ly.n_h[0] = 16; Ly.n_h[1] = 0; }
        else if (Np == 208) { Ly.nh = 2; Ly.n_h[0] = 112; Ly.n_h[1] = 96; }
        else { Ly.nh = 2; Ly.n_h[0] = Np / 2; Ly.n_h[1] = Np / 2; }
        Ly.base[0] = 0; Ly.base[1] = Ly.n_h[0];
        for (int h = 0; h < Ly.nh; h++)
            for (int r = 0; r < 2; r++) {
                const int nhr = Ly.n_h[h] / 2, row0 = Ly.base[h] + r * nhr;
                Ly.goff[h][r] = (unsigned)(blob.size() * 2);
                tc2_pack_rows(blob, hw, N_src, K_src, cm, row0, nhr, scale);
                Ly.boff[h][r] = (unsigned)(blob.size() * 2);
                tc2_pack_bias(blob, hb, N_src, row0, nhr);
            }
    };
    for (int l = 0; l < 9; l++) {
        std::vector<float> hw, hb;
        int N = (l == 8) ? 3 : 256;
        if (!fetch(w->resd_w[l], (size_t)N * rK[l], hw) || !fetch(w->resd_b[l], N, hb)) { err = "tc8_upload: copy failed"; return 1; }
        std::vector<int> cm = (l == 0) ? ident(63, 64) : (l == 4 ? ident(319, 320) : ident(256, 256));
        pack(l, hw, hb, N, rK[l], cm, l == 8 ? 16 : 256, 1.f, (l == 0) ? 0 : 4, (l == 8) ? TC_EPI_RESD_FINAL : TC_EPI_RELU);
    }
    const float rs2 = (float)(1.0 / std::sqrt(2.0));
    for (int l = 0; l < 9; l++) {
        std::vector<float> hw, hb;
        if (!fetch(w->sdf_w[l], (size_t)sN[l] * sK[l], hw) || !fetch(w->sdf_b[l], sN[l], hb)) { err = "tc8_upload: copy failed"; return 1; }
        std::vector<int> cm;
        float scale = 1.f;
        if (l == 0) cm = ident(51, 64);
        else if (l == 4) {
            cm.assign(256, -1);
            for (int j = 0; j < 48; j++) cm[j] = 205 + j;
            for (int j = 0; j < 205; j++) cm[48 + j] = j;
            for (int j = 0; j < 3; j++) cm[253 + j] = 205 + 48 + j;
            scale = rs2;
        } else cm = ident(256, 256);
        int Np = 256, Nsrc = sN[l];
        if (l == 3) Np = 208;
        if (l == 8) { Np = 16; Nsrc = 1; }
        pack(9 + l, hw, hb, Nsrc, sK[l], cm, Np, scale, (l == 0) ? 0 : 4, (l == 8) ? TC_EPI_SDF_FINAL : (l == 3 ? TC_EPI_S3 : TC_EPI_SOFTPLUS));
    }
    tc8_free(t);
    if (cudaMalloc((void**)&t.blob, blob.size() * 2) != cudaSuccess) { err = "tc8_upload: cudaMalloc failed"; return 1; }
    cudaMemcpy(t.blob, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice);
    P.blob = t.blob;
    t.ready = true;
    return 0;
}

static void tc8_set_frame(Tc8Weights& t, const FrameConst* fc, cudaStream_t st, int64_t& launches) {
    const Tc8Layer& a = t.p.layer[0]; const Tc8Layer& b = t.p.layer[4];
    k_tc8_pack_bias<<<1, 256, 0, st>>>(&fc->resd_b0[0], &fc->resd_b4[0], t.blob, a.boff[0][0], a.boff[0][1], a.boff[1][0], a.boff[1][1],
                                       b.boff[0][0], b.boff[0][1], b.boff[1][0], b.boff[1][1]);
    launches++;
}

static void tc8_distance(Tc8Weights& t, const float* bpts, float* out, const int* count, float resd_limit, int sms, cudaStream_t st,
                         int64_t& launches) {
    Tc8Params p = t.p;
    p.bpts = bpts; p.out = out; p.count = count; p.resd_limit = resd_limit; p.dbg = t.dbg;
    k_mlp_tc8<<<(sms / 2) * 2, TC6_THREADS, TC8_SMEM_BYTES, st>>>(p);
    launches++;
}

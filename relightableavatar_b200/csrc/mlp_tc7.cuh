// k_mlp_tc7: the schedule of k_mlp_tc6 (CTA pairs, two pair-tiles in flight, see mlp_tc6.cuh) with the A operand of the two
// OUTPUT layers held in tensor memory.
//
// Why (profiles/r01_k_mlp_tc6_timeline.txt): with both operands in shared memory a tcgen05.mma of this kernel reads its SM's
// 128 x 16 A tile (4 KB) at ~32 B/clk -- 128 clk per K=16 step whatever N is.  At N=256 that equals the FLOP floor, but the
// residual net's last layer (256 -> 3, padded N=16) and the SDF net's last layer (256 -> 1 for a distance query) cost the same
// ~2200 clk per tile as a full 256 x 256 layer for 1-3 useful output columns: 2 of 18 layer slots, ~9 % of the kernel.
// Here the epilogue warps of the layer BEFORE (R7 / S7) write their fp16 activations with tcgen05.st IN PLACE over the lower 128
// columns of the slot's (now dead) fp32 accumulator -- two halves per 32-bit column, exactly the K-major A layout of the
// `tcgen05.mma [d], [a], b-desc` form -- and the 17 MMAs of the output layer read A from there and accumulate into columns
// [128, 144) of the same slot: no shared-memory A pass, no 64 KB A store.  Same fp16 operands, same fp32 accumulation: the
// distances stay bit-identical to k_mlp_tc / k_mlp_tc2 / k_mlp_tc6 (tests/test_gpu_parity.py::test_two_cta_kernel_variant_matches_single_cta).
// (Tried first and measured slower: the output layers as fp32 dot products on the CUDA cores inside the R7 / S7 epilogues --
// 64-192 weights per thread with no reuse; whether fetched by warp-uniform __ldg or as kernel parameters through the uniform
// constant path, the dependent weight fetches made those epilogues 4000-7000 clk: profiles/r02_k_mlp_tc7_attempts.txt.)
//   smem / warp roles / barriers: as k_mlp_tc6.
#pragma once
#include "mlp_tc6.cuh"

#define TC7_TS_LAYER(l) ((l) == 8 || (l) == 17)        // N = 16 output layers: A from tensor memory
#define TC7_FEEDS_TS(l) ((l) == 7 || (l) == 16)        // their inputs are written to tensor memory instead of shared memory

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// A (M x 16, fp16 pairs in 8 consecutive 32-bit columns) from tensor memory, B from shared memory
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// hidden-layer epilogue whose output feeds a TS layer: this warp's 64 accumulator columns -> activation -> 32 packed fp16 pairs,
// written over columns [32 cg, 32 cg + 32) of the same slot once all four column-group warps of the lane quarter have read theirs
template <bool SOFTPLUS>
__device__ __forceinline__ void epi_hidden64_tmem(uint32_t t_lane, int cg, int q) {
    uint32_t ra[32], rb[32], h[32];
    const int cbase = cg * 64;
    tmem_ld32(t_lane + (uint32_t)cbase, ra);
    tmem_ld32(t_lane + (uint32_t)(cbase + 32), rb);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j++) {
        h[j] = SOFTPLUS ? h2_softplus100(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]), j & 1)
                        : h2_relu(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
        h[16 + j] = SOFTPLUS ? h2_softplus100(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]), j & 1)
                             : h2_relu(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]));
    }
    tc_fence_before();
    named_bar_sync(1 + q, 128);          // the quarter's accumulator columns are all in registers: they may be overwritten
    tc_fence_after();
    tmem_st32(t_lane + (uint32_t)(cg * 32), h);
    tmem_st_wait();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC6_THREADS, 1) k_mlp_tc7(const __grid_constant__ Tc2Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_act0 = s_base;                                   // ACT[p] = s_act0 + p * TC_ACT_BYTES
    const uint32_t s_pe0 = s_base + 2 * TC_ACT_BYTES;                 // PE[p]  = s_pe0 + p * TC_PE_BYTES
    const uint32_t s_w = s_pe0 + 2 * TC_PE_BYTES;
    const uint32_t s_bar = s_w + TC6_STAGES * TC6_STAGE_BYTES;
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 64, bar_act0 = s_bar + 128, bar_acc0 = s_bar + 144;
    const uint32_t s_tmem = s_bar + 160;
    const uint32_t s_ones = s_bar + 256;    // 256 B: core matrix of rows [1,1,0,0,0,0,0,0], then a zero core matrix
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - s_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int count = *P.count;
    const int n_tiles = (count + TC_TILE_M - 1) / TC_TILE_M;
    const int n_quads = (n_tiles + 3) / 4;       // a cluster works on 4 tiles at a time: slot p, rank r -> tile 4 g + 2 p + r
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC6_STAGES; s++) { mbar_init(bar_full + 8 * s, rank == 0 ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int p = 0; p < 2; p++) { mbar_init(bar_act0 + 8 * p, 2 * TC6_EPI_WARPS); mbar_init(bar_acc0 + 8 * p, 1); }
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

    // ---- the schedule: one flat sequence of (slot, tile, layer) items that all three roles walk in the same order.
    // Step g serves slot 0's item g and slot 1's item g - skew (item = 18 * tile + layer): slot 1 runs `skew` layers behind
    // slot 0, so that the two thin stretches of the stack (R8 -> S0 and S8 -> next tile's encode -> R0: 5-MMA layers chained through
    // tanh / sin-cos epilogues, ~6000 clk of dependent latency with ~700 clk of tensor work) of one slot fall into the other
    // slot's 256-wide layers instead of coinciding with its own thin stretch (profiles/r02_k_mlp_tc7_timeline.txt).
    const int nq = (cluster_id < n_quads) ? (n_quads - cluster_id + n_clusters - 1) / n_clusters : 0;     // quads of this cluster
    const int total = nq * TC_LAYERS;
    const int skew = P.skew;

    if (warp == TC6_WARP_TMA) {
        // ===================== TMA producer: this CTA's half (N/2 rows) of every chunk, once per item =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int g = 0; g < total + skew; g++)
                for (int p = 0; p < 2; p++) {
                    const int gi = g - p * skew;
                    if (gi < 0 || gi >= total) continue;
                    const int l = gi % TC_LAYERS;
                    const uint32_t bytes = (uint32_t)(P.layer[l].N / 2) * TC_KCHUNK * 2;
                    const unsigned char* src = P.blob + P.layer[l].goff[rank];
                    const unsigned char* bsrc = P.blob + P.layer[l].boff[rank];
                    const int nch = P.layer[l].nchunks / 2;      // 64-wide K chunks (all layer widths are multiples of 64)
                    for (int c = 0; c <= nch; c++, it++) {
                        const uint32_t s = it & (TC6_STAGES - 1), ph = (it / TC6_STAGES) & 1;
                        const uint32_t nb = (c < nch) ? 2 * bytes : bytes / 2;
                        const unsigned char* gsrc = (c < nch) ? src + (size_t)c * 2 * bytes : bsrc;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, nb);
                        tma_bulk_g2s(s_w + s * TC6_STAGE_BYTES, gsrc, nb, bar_full + 8 * s);
                    }
                }
        }
    } else if (warp == TC6_WARP_MMA) {
        if (lane == 0 && rank == 1) {
            // ===================== peer: tell the leader when my half of each chunk has landed =====================
            uint32_t it = 0;
            for (int g = 0; g < total + skew; g++)
                for (int p = 0; p < 2; p++) {
                    const int gi = g - p * skew;
                    if (gi < 0 || gi >= total) continue;
                    const int n = P.layer[gi % TC_LAYERS].nchunks / 2 + 1;
                    for (int c = 0; c < n; c++, it++) {
                        const uint32_t s = it & (TC6_STAGES - 1), ph = (it / TC6_STAGES) & 1;
                        mbar_wait(bar_full + 8 * s, ph);
                        mbar_arrive_remote(bar_full + 8 * s, 0);
                    }
                }
        } else if (rank == 0) {
            // ===================== leader: MMA issuer for both slots of the pair =====================
            // The whole warp runs this loop with warp-uniform values (descriptors live in uniform registers); one elected
            // lane issues the tcgen05 instructions.  Issuing from inside `if (lane == 0)` makes the compiler wrap every
            // UTCHMMA in an ELECT / R2UR.BROADCAST waterfall loop (~100 clk per MMA on the single issuing thread).
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            uint32_t it = 0, lc[2] = {0, 0};
            for (int g = 0; g < total + skew; g++)
                for (int p = 0; p < 2; p++) {
                    const int gi = g - p * skew;
                    if (gi < 0 || gi >= total) continue;
                    const int l = gi % TC_LAYERS;
                    const int N = P.layer[l].N;
                    const uint32_t idesc = make_idesc2_f16(N);
                    const uint32_t lbo_b = (uint32_t)(N / 2) * 16u;
                    const int nch = P.layer[l].nchunks / 2;
                    const int pe_from = P.layer[l].pe_from / 2;      // in 64-wide chunks (0, 4 or "never")
                    const uint32_t s_act = s_act0 + p * TC_ACT_BYTES, s_pe = s_pe0 + p * TC_PE_BYTES;
                    const bool ts = TC7_TS_LAYER(l);
                    // output layers: A = columns [0, 128) of the slot (written by the previous epilogue), D = columns [128, 144)
                    const uint32_t tmem_a = tmem_u + (uint32_t)p * 256u;
                    const uint32_t tmem_d = tmem_a + (ts ? 128u : 0u);
                    TC_TL(const bool rec = P.dbg && blockIdx.x == 0 && gi / TC_LAYERS == 1 && lane == 0;)
                    TC_TL(unsigned long long wsum = 0;)
                    TC_TL(if (rec) P.dbg[(l * 2 + p) * 8 + 0] = clock64();)
                    mbar_wait(bar_act0 + 8 * p, lc[p] & 1);
                    lc[p]++;
                    tc_fence_after();
                    TC_TL(if (rec) P.dbg[(l * 2 + p) * 8 + 1] = clock64();)
                    for (int c = 0; c <= nch; c++, it++) {
                        const uint32_t s = it & (TC6_STAGES - 1), ph = (it / TC6_STAGES) & 1;
                        TC_TL(unsigned long long w0 = clock64();)
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        TC_TL(unsigned long long w1 = clock64(); wsum += w1 - w0;)
                        const uint32_t b_base = s_w + s * TC6_STAGE_BYTES;
                        if (c < nch) {
                            const uint32_t a_base = (c >= pe_from) ? (s_pe + (uint32_t)(c - pe_from) * 8u * 2048u) : (s_act + (uint32_t)c * 8u * 2048u);
                            uint64_t ad[4], bd[4];
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {             // 4 K=16 steps: A advances 2 core-matrix columns, B likewise
                                ad[kk] = make_sdesc(a_base + (uint32_t)kk * 2u * 2048u, 2048u, 128u);
                                bd[kk] = make_sdesc(b_base + (uint32_t)kk * 2u * lbo_b, lbo_b, 128u);
                            }
                            if (elect_one()) {
                                if (ts) {                             // K step kk of chunk c: 8 columns of packed fp16 pairs
                                    const uint32_t a0 = tmem_a + (uint32_t)c * 32u;
                                    umma2_f16_ts(tmem_d, a0, bd[0], idesc, c ? 1u : 0u);
                                    umma2_f16_ts(tmem_d, a0 + 8u, bd[1], idesc, 1u);
                                    umma2_f16_ts(tmem_d, a0 + 16u, bd[2], idesc, 1u);
                                    umma2_f16_ts(tmem_d, a0 + 24u, bd[3], idesc, 1u);
                                } else {
                                    umma2_f16(tmem_d, ad[0], bd[0], idesc, c ? 1u : 0u);
                                    umma2_f16(tmem_d, ad[1], bd[1], idesc, 1u);
                                    umma2_f16(tmem_d, ad[2], bd[2], idesc, 1u);
                                    umma2_f16(tmem_d, ad[3], bd[3], idesc, 1u);
                                }
                                umma_commit2(bar_empty + 8 * s);      // frees this stage in both CTAs when the MMAs retire
                            }
                        } else {
                            const uint64_t ad = make_sdesc(s_ones, 128u, 0u);
                            const uint64_t bd = make_sdesc(b_base, lbo_b, 128u);
                            if (elect_one()) {
                                umma2_f16(tmem_d, ad, bd, idesc, 1u);
                                umma_commit2(bar_empty + 8 * s);
                                umma_commit2(bar_acc0 + 8 * p);       // accumulator of (slot p, layer l) complete, both CTAs
                            }
                        }
                        __syncwarp();
                    }
                    TC_TL(if (rec) { P.dbg[(l * 2 + p) * 8 + 2] = clock64(); P.dbg[(l * 2 + p) * 8 + 7] = wsum; })
                }
        }
    } else {
        // ===================== epilogue warps: all 16 serve the items in schedule order =====================
        const int e = warp;
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int cg = e >> 2;                  // column group: cols [64 cg, 64 cg + 64)
        const int row = q * 32 + lane;
        const uint32_t t_lane0 = tmem + ((uint32_t)(q * 32) << 16);
        float3 bp[2], cp[2];
        int gidx[2];
        uint32_t lc[2] = {0, 0};

        auto prologue = [&](int quad, int p) {      // PE10(bp) of the slot's new tile -> PE[p]; input of layer R0 ready
            const int tile = quad * 4 + p * 2 + (int)rank;
            gidx[p] = tile * TC_TILE_M + row;
            bp[p] = make3(0.f, 0.f, 0.f);
            if (gidx[p] < count) bp[p] = make3(P.bpts[(size_t)gidx[p] * 3], P.bpts[(size_t)gidx[p] * 3 + 1], P.bpts[(size_t)gidx[p] * 3 + 2]);
            cp[p] = bp[p];
            write_pe<10>(s_pe0 + p * TC_PE_BYTES, row, bp[p], cg * 2, cg * 2 + 2);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar_act0 + 8 * p, 0);
        };

        if (nq > 0) { prologue(cluster_id, 0); prologue(cluster_id, 1); }
#pragma unroll 1
        for (int g = 0; g < total + skew; g++) {
#pragma unroll 1
            for (int p = 0; p < 2; p++) {
                const int gi = g - p * skew;
                if (gi < 0 || gi >= total) continue;
                const int l = gi % TC_LAYERS, ti = gi / TC_LAYERS;
                const int epi = P.layer[l].epi;
                const uint32_t s_act = s_act0 + p * TC_ACT_BYTES, s_pe = s_pe0 + p * TC_PE_BYTES;
                const uint32_t t_lane = t_lane0 + (uint32_t)p * 256u;
                TC_TL(const bool rec = P.dbg && blockIdx.x == 0 && ti == 1 && warp == 0 && lane == 0;)
                TC_TL(if (rec) P.dbg[(l * 2 + p) * 8 + 3] = clock64();)
                mbar_wait(bar_acc0 + 8 * p, lc[p] & 1);
                lc[p]++;
                tc_fence_after();
                TC_TL(if (rec) P.dbg[(l * 2 + p) * 8 + 4] = clock64();)
                if (TC7_FEEDS_TS(l)) {
                    if (epi == TC_EPI_RELU) epi_hidden64_tmem<false>(t_lane, cg, q);
                    else epi_hidden64_tmem<true>(t_lane, cg, q);
                } else if (epi == TC_EPI_RELU) {
                    epi_hidden64<false>(t_lane, s_act, row, cg);
                } else if (epi == TC_EPI_SOFTPLUS) {
                    epi_hidden64<true>(t_lane, s_act, row, cg);
                } else if (epi == TC_EPI_S3) {
                    // S3: 205 outputs -> ACT cols [48, 253); PE8(cp) features 0..47 -> cols [0,48), 48..50 -> cols 253..255.
                    // 26 groups of 8 accumulator columns over the 4 column-group warps: 7 / 7 / 6 / 6
                    const int g0 = (cg < 2) ? cg * 7 : 14 + (cg - 2) * 6;
                    const int g1 = g0 + ((cg < 2) ? 7 : 6);
#pragma unroll 1
                    for (int gg = g0; gg < g1; gg++) {
                        const int c0 = gg * 8;
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
                            float p48 = pe_feature(cp[p], 48), p49 = pe_feature(cp[p], 49), p50 = pe_feature(cp[p], 50);
                            h[2] = (h[2] & 0x0000FFFFu) | (pack_h2(0.f, p48) & 0xFFFF0000u);
                            h[3] = pack_h2(p49, p50);
                        }
                        st_shared_v4(s_act + (uint32_t)((48 + c0) >> 3) * 2048u + (uint32_t)row * 16u, h[0], h[1], h[2], h[3]);
                    }
                    if (cg >= 2) {            // PE8 features 0..47 (6 chunks) copied from the PE buffer
                        for (int ch = (cg - 2) * 3; ch < (cg - 2) * 3 + 3; ch++) {
                            uint32_t a, b, c, d;
                            ld_shared_v4(s_pe + (uint32_t)ch * 2048u + (uint32_t)row * 16u, a, b, c, d);
                            st_shared_v4(s_act + (uint32_t)ch * 2048u + (uint32_t)row * 16u, a, b, c, d);
                        }
                    }
                } else if (epi == TC_EPI_RESD_FINAL) {
                    uint32_t r[16];
                    tmem_ld16(t_lane + 128u, r);
                    tmem_ld_wait();
                    float rx = tanhf(__uint_as_float(r[0])) * P.resd_limit;
                    float ry = tanhf(__uint_as_float(r[1])) * P.resd_limit;
                    float rz = tanhf(__uint_as_float(r[2])) * P.resd_limit;
                    cp[p] = make3(bp[p].x + rx, bp[p].y + ry, bp[p].z + rz);
                    // PE8(cp): 51 features padded to 64 -> PE buffer (input of S0, later copied into the S4 skip columns)
                    write_pe<8>(s_pe, row, cp[p], cg * 2, cg * 2 + 2);
                } else {   // TC_EPI_SDF_FINAL
                    if (cg == 0) {
                        uint32_t r[16];
                        tmem_ld16(t_lane + 128u, r);
                        tmem_ld_wait();
                        if (gidx[p] < count) P.out[gidx[p]] = __uint_as_float(r[0]);
                    }
                }
                TC_TL(if (rec) P.dbg[(l * 2 + p) * 8 + 5] = clock64();)
                if (l + 1 < TC_LAYERS) {
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_act0 + 8 * p, 0);
                    TC_TL(if (rec) P.dbg[(l * 2 + p) * 8 + 6] = clock64();)
                } else if (ti + 1 < nq) {
                    tc_fence_before();       // the slot's accumulator has been read; its next tile may start
                    prologue(cluster_id + (ti + 1) * n_clusters, p);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == TC6_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static int tc7_init(std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc7, cudaFuncAttributeMaxDynamicSharedMemorySize, TC6_SMEM_BYTES);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_mlp_tc7): ") + cudaGetErrorString(e); return 1; }
    return 0;
}

static void tc7_distance(Tc2Weights& t, const float* bpts, float* out, const int* count, float resd_limit, int sms, cudaStream_t st,
                         int64_t& launches) {
    Tc2Params p = t.p;
    p.bpts = bpts; p.out = out; p.count = count; p.resd_limit = resd_limit; p.dbg = t.dbg;
    p.skew = t.skew;
    k_mlp_tc7<<<(sms / 2) * 2, TC6_THREADS, TC6_SMEM_BYTES, st>>>(p);      // one cluster of 2 CTAs per TPC (compile-time __cluster_dims__)
    launches++;
}

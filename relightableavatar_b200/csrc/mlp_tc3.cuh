// Cluster-multicast variant of the single-CTA fused MLP kernel (mlp_tc.cuh): two CTAs on the two SMs of a TPC run
// independent 128-point tile pipelines (own MMA issuer, own TMEM, cta_group::1) but fetch the weight stream TOGETHER:
// CTA r loads bytes [r*n/2, (r+1)*n/2) of every chunk image and multicasts them into both CTAs' rings
// (cp.async.bulk ... .multicast::cluster).  Each SM then pulls only half of the 1.95 MB weight stream per tile pair
// from L2 -- the per-SM L2 read port (~64 B/clk) is what paces the single-CTA kernel's MMA phases -- and receives the
// other half over the SM-to-SM network.  A stage is refilled once BOTH consumers have retired it (empty barrier
// count 2, tcgen05.commit multicast).  Same weight blob, same arithmetic, bit-identical results.
#pragma once
#include "mlp_tc.cuh"
#include "mlp_tc2.cuh"   // cluster helpers (cluster_ctarank, cluster_sync_all, mbar_wait_cluster)

#ifndef TC3_CS
#define TC3_CS 2          // cluster size (4 and 8 measured 2x slower: 42 ms/frame): CTA r loads slice r of TC3_CS of every chunk and multicasts it to all
#endif
#define TC3_MASK ((uint16_t)((1u << TC3_CS) - 1u))

__device__ __forceinline__ void tma_bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "h"(TC3_MASK) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(TC3_MASK)
                 : "memory");
}

// ------------------------------------------------------------------------------------------ the kernel
__global__ void __cluster_dims__(TC3_CS, 1, 1) __launch_bounds__(TC_THREADS, 2) k_mlp_tc3(const __grid_constant__ TcParams P) {
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
    const uint32_t rank = cluster_ctarank();
    const int n_pairs = (n_tiles + TC3_CS - 1) / TC3_CS;      // groups of TC3_CS tiles, one per CTA of the cluster
    const int cluster_id = blockIdx.x / TC3_CS, n_clusters = gridDim.x / TC3_CS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, TC3_CS); }   // empty: both CTAs' consumers must retire a stage
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
    cluster_sync_all();          // the peer's barriers must exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: stream every layer's weight chunks (+ its bias chunk), once per tile =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
                for (int l = 0; l < TC_LAYERS; l++) {
                    const uint32_t bytes = (uint32_t)P.layer[l].N * TC_KCHUNK * 2;
                    const unsigned char* src = P.blob + P.layer[l].goff;
                    const int nch = P.layer[l].nchunks;
                    for (int c = 0; c <= nch; c++, it++) {
                        uint32_t s = it & 1, ph = (it >> 1) & 1;
                        const uint32_t nb = (c < nch) ? bytes : bytes / 2;
                        const unsigned char* g = (c < nch) ? src + (size_t)c * bytes : P.blob + P.layer[l].boff;
                        mbar_wait_cluster(bar_empty + 8 * s, ph ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, nb);          // the whole chunk: my half + the peer's half
                        tma_bulk_g2s_mc(s_w + s * TC_STAGE_BYTES + rank * (nb / TC3_CS), g + rank * (nb / TC3_CS), nb / TC3_CS, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, lc = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
                for (int l = 0; l < TC_LAYERS; l++, lc++) {
                    const int N = P.layer[l].N;
                    const uint32_t idesc = make_idesc_f16(N);
                    const uint32_t lbo_b = (uint32_t)N * 16u;
                    const int nch = P.layer[l].nchunks;
                    mbar_wait(bar_act, lc & 1);
                    tc_fence_after();
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
                        umma_commit_mc(bar_empty + 8 * s);   // frees this stage in BOTH CTAs' producers when these MMAs retire
                    }
                    umma_commit(bar_acc);                     // accumulator of this layer complete
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
        for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
            const int tile = pair * TC3_CS + (int)rank;
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
    // teardown: nobody leaves while the peer may still multicast into this CTA's shared memory
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}


static int tc3_init(std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_mlp_tc3): ") + cudaGetErrorString(e); return 1; }
    return 0;
}
static void tc3_distance(TcWeights& t, const float* bpts, float* out, const int* count, float resd_limit, int sms, cudaStream_t st,
                         int64_t& launches) {
    TcParams p = t.p;
    p.bpts = bpts; p.out = out; p.count = count; p.resd_limit = resd_limit;
    k_mlp_tc3<<<2 * sms, TC_THREADS, TC_SMEM_BYTES, st>>>(p);
    launches++;
}

// 2-CTA (cta_group::2) variant of the fused distance-query MLP kernel (see mlp_tc.cuh for the single-CTA design).
//
// A cluster of two CTAs on the two SMs of a TPC evaluates two 128-point tiles at once: every tcgen05.mma is
// M=256 x N<=256 x K=16 across the pair.  Each CTA keeps ITS 128 activation rows (A) and accumulators, but only HALF
// of every weight chunk (B: N/2 rows) -- the tensor cores of both SMs read both halves.  Per SM this halves the
// shared-memory bandwidth spent on B operand reads, the TMA fill traffic and the L2 reads of the weight stream, which
// is what bounds the single-CTA kernel (A 32 + B 64 + fill 64 + epilogue 32 B/clk against 128 B/clk of smem).
//   rank 0 (leader): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue
//   rank 1 (peer)  : warp 0 TMA producer, warp 1 forwards "my half of the chunk landed" to the leader, warps 2-9 epilogue
// Barriers (same smem offsets in both CTAs): full[4] (leader: own TMA + peer forward), empty[4] / acc_ready
// (tcgen05.commit multicast to both CTAs), act_ready (leader only; 8 local + 8 remote epilogue-warp arrivals).
#pragma once
#include "mlp_tc.cuh"

#define TC2_STAGES 4
#define TC2_STAGE_BYTES (TC_KCHUNK * 128 * 2)
#define TC2_SMEM_BYTES (TC_ACT_BYTES + TC_PE_BYTES + TC2_STAGES * TC2_STAGE_BYTES + 512)

struct Tc2Layer {
    int N;              // accumulator columns of the pair-wide MMA (multiple of 16); each CTA holds N/2 weight rows
    int nchunks, pe_from, epi;
    unsigned goff[2];   // per rank: byte offset of the first chunk image [4][N/2][8 halves]
    unsigned boff[2];   // per rank: bias chunk [2][N/2][8]
};
struct Tc2Params {
    Tc2Layer layer[TC_LAYERS];
    const unsigned char* blob;
    const float* bpts;
    float* out;
    const int* count;
    float resd_limit;
    unsigned long long* dbg;   // optional timeline (debug builds, tools/tc6_timeline.py)
    int skew;                  // k_mlp_tc7: layers slot 1 runs behind slot 0 (env RA_TC_SKEW, default 4)
};
struct Tc2Weights {
    unsigned char* blob = nullptr;
    int skew = 4;                      // env RA_TC_SKEW (read in ra_create)
    Tc2Params p{};
    bool ready = false;
    unsigned long long* dbg = nullptr;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"      // default .release.cta (a cluster-scope release costs ~1000+ clk per arrive)
        "}" ::"r"(local_bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc2_f16(int N) {   // M = 256 across the CTA pair
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 2) k_mlp_tc2(const __grid_constant__ Tc2Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_act = s_base;
    const uint32_t s_pe = s_base + TC_ACT_BYTES;
    const uint32_t s_w = s_pe + TC_PE_BYTES;
    const uint32_t s_bar = s_w + TC2_STAGES * TC2_STAGE_BYTES;
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_act = s_bar + 64, bar_acc = s_bar + 72;
    const uint32_t s_ones = s_bar + 128;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (s_bar - s_base) + 96);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int count = *P.count;
    const int n_tiles = (count + TC_TILE_M - 1) / TC_TILE_M;
    const int n_pairs = (n_tiles + 1) / 2;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC2_STAGES; s++) { mbar_init(bar_full + 8 * s, rank == 0 ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_act, 16);
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 64) {
        uint32_t v = (threadIdx.x < 32 && (threadIdx.x & 3) == 0) ? 0x3C003C00u : 0u;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_ones + threadIdx.x * 4), "r"(v) : "memory");
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_bar + 96), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: this CTA's half (N/2 rows) of every weight chunk =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
                for (int l = 0; l < TC_LAYERS; l++) {
                    const uint32_t bytes = (uint32_t)(P.layer[l].N / 2) * TC_KCHUNK * 2;
                    const unsigned char* src = P.blob + P.layer[l].goff[rank];
                    const int nch = P.layer[l].nchunks;
                    for (int c = 0; c <= nch; c++, it++) {
                        uint32_t s = it & (TC2_STAGES - 1), ph = (it / TC2_STAGES) & 1;
                        const uint32_t nb = (c < nch) ? bytes : bytes / 2;
                        const unsigned char* g = (c < nch) ? src + (size_t)c * bytes : P.blob + P.layer[l].boff[rank];
                        mbar_wait(bar_empty + 8 * s, ph ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, nb);
                        tma_bulk_g2s(s_w + s * TC2_STAGE_BYTES, g, nb, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 1) {
            // ===================== peer: tell the leader when my half of each chunk has landed =====================
            uint32_t it = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters)
                for (int l = 0; l < TC_LAYERS; l++)
                    for (int c = 0; c <= P.layer[l].nchunks; c++, it++) {
                        uint32_t s = it & (TC2_STAGES - 1), ph = (it / TC2_STAGES) & 1;
                        mbar_wait(bar_full + 8 * s, ph);
                        mbar_arrive_remote(bar_full + 8 * s, 0);
                    }
        } else if (lane == 0 && rank == 0) {
            // ===================== leader: MMA issuer for the pair =====================
            uint32_t it = 0, lc = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
                for (int l = 0; l < TC_LAYERS; l++, lc++) {
                    const int N = P.layer[l].N;
                    const uint32_t idesc = make_idesc2_f16(N);
                    const uint32_t lbo_b = (uint32_t)(N / 2) * 16u;
                    const int nch = P.layer[l].nchunks;
                    mbar_wait_cluster(bar_act, lc & 1);
                    tc_fence_after();
                    for (int c = 0; c <= nch; c++, it++) {
                        uint32_t s = it & (TC2_STAGES - 1), ph = (it / TC2_STAGES) & 1;
                        mbar_wait_cluster(bar_full + 8 * s, ph);
                        tc_fence_after();
                        uint32_t b_base = s_w + s * TC2_STAGE_BYTES;
                        if (c < nch) {
                            uint32_t a_base = (c >= P.layer[l].pe_from) ? (s_pe + (uint32_t)(c - P.layer[l].pe_from) * 4u * 2048u)
                                                                         : (s_act + (uint32_t)c * 4u * 2048u);
#pragma unroll
                            for (int kk = 0; kk < TC_KCHUNK / 16; kk++) {
                                uint64_t ad = make_sdesc(a_base + (uint32_t)kk * 2u * 2048u, 2048u, 128u);
                                uint64_t bd = make_sdesc(b_base + (uint32_t)kk * 2u * lbo_b, lbo_b, 128u);
                                umma2_f16(tmem, ad, bd, idesc, (c | kk) ? 1u : 0u);
                            }
                        } else {
                            uint64_t ad = make_sdesc(s_ones, 128u, 0u);
                            uint64_t bd = make_sdesc(b_base, lbo_b, 128u);
                            umma2_f16(tmem, ad, bd, idesc, 1u);
                        }
                        umma_commit2(bar_empty + 8 * s);
                    }
                    umma_commit2(bar_acc);
                }
            }
        }
    } else {
        // ===================== epilogue warps (identical in both CTAs; each CTA owns one tile of the pair) =====================
        const int e = warp - 2;
        const int q = warp & 3;
        const int half = e >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t lc = 0;
        for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
            const int tile = pair * 2 + (int)rank;
            const int gidx = tile * TC_TILE_M + row;
            float3 bp = make3(0.f, 0.f, 0.f);
            if (gidx < count) bp = make3(P.bpts[(size_t)gidx * 3], P.bpts[(size_t)gidx * 3 + 1], P.bpts[(size_t)gidx * 3 + 2]);
            float3 cp = bp;
            write_pe<10>(s_pe, row, bp, half * 4, half * 4 + 4);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar_act, 0);
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
                    const int a0 = half ? 104 : 0;
#pragma unroll 1
                    for (int cb = 0; cb < 13; cb++) {
                        const int c0 = a0 + cb * 8;
                        uint32_t r[16];
                        tmem_ld16(t_lane + (uint32_t)(c0 & ~15), r);
                        tmem_ld_wait();
                        const int o = c0 & 15;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = __uint_as_float(o ? r[8 + j] : r[j]);
                        uint32_t h[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) h[j] = h2_softplus100(v[2 * j], v[2 * j + 1], j & 1);
                        if (c0 == 200) {
                            float p48 = pe_feature(cp, 48), p49 = pe_feature(cp, 49), p50 = pe_feature(cp, 50);
                            h[2] = (h[2] & 0x0000FFFFu) | (pack_h2(0.f, p48) & 0xFFFF0000u);
                            h[3] = pack_h2(p49, p50);
                        }
                        st_shared_v4(s_act + (uint32_t)((48 + c0) >> 3) * 2048u + (uint32_t)row * 16u, h[0], h[1], h[2], h[3]);
                    }
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
                    write_pe<8>(s_pe, row, cp, half * 4, half * 4 + 4);
                } else {
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
                    if (lane == 0) mbar_arrive_remote(bar_act, 0);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// per frame: pose-folded biases of residual layers 0 / 4 -> fp16 hi/lo bias chunks ([2][128][8] per rank)
__global__ void k_tc2_pack_bias(const float* __restrict__ b0, const float* __restrict__ b4, __half* d00, __half* d01, __half* d40, __half* d41) {
    int n = threadIdx.x;
    for (int w = 0; w < 2; w++) {
        const float b = (w ? b4 : b0)[n];
        __half* d = w ? (n < 128 ? d40 : d41) : (n < 128 ? d00 : d01);
        int r = n & 127;
        __half hi = __float2half_rn(b);
        __half lo = __float2half_rn(b - __half2float(hi));
        d[r * 8 + 0] = hi; d[r * 8 + 1] = lo;
        for (int j = 2; j < 8; j++) d[r * 8 + j] = __float2half_rn(0.f);
        for (int j = 0; j < 8; j++) d[(128 + r) * 8 + j] = __float2half_rn(0.f);
    }
}

// ------------------------------------------------------------------------------------------ host side
static void tc2_pack_rows(std::vector<__half>& blob, const std::vector<float>& w, int N_src, int K_src, const std::vector<int>& colmap,
                          int row0, int Nh, float scale) {
    int K = (int)colmap.size(), nch = K / TC_KCHUNK;
    size_t base = blob.size();
    blob.resize(base + (size_t)nch * TC_KCHUNK * Nh, __float2half(0.f));
    for (int c = 0; c < nch; c++)
        for (int kq = 0; kq < TC_KCHUNK / 8; kq++)
            for (int n = 0; n < Nh; n++)
                for (int j = 0; j < 8; j++) {
                    int k = c * TC_KCHUNK + kq * 8 + j, sn = row0 + n;
                    float v = (sn < N_src && colmap[k] >= 0) ? w[(size_t)sn * K_src + colmap[k]] * scale : 0.f;
                    blob[base + (size_t)c * TC_KCHUNK * Nh + ((size_t)kq * Nh + n) * 8 + j] = __float2half_rn(v);
                }
}
static void tc2_pack_bias(std::vector<__half>& blob, const std::vector<float>& b, int N_src, int row0, int Nh) {
    size_t base = blob.size();
    blob.resize(base + (size_t)2 * Nh * 8, __float2half(0.f));
    for (int n = 0; n < Nh; n++) {
        if (row0 + n >= N_src) continue;
        __half hi = __float2half_rn(b[row0 + n]);
        blob[base + (size_t)n * 8 + 0] = hi;
        blob[base + (size_t)n * 8 + 1] = __float2half_rn(b[row0 + n] - __half2float(hi));
    }
}

static int tc2_init(Tc2Weights&, std::string& err) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(k_mlp_tc2): ") + cudaGetErrorString(e); return 1; }
    return 0;
}
static void tc2_free(Tc2Weights& t) { if (t.blob) cudaFree(t.blob); t.blob = nullptr; }

static int tc2_upload(Tc2Weights& t, const ra_weights* w, int cond, std::string& err, cudaStream_t st) {
    auto fetch = [&](const float* src, size_t n, std::vector<float>& dst) -> bool {
        dst.resize(n);
        return cudaMemcpyAsync(dst.data(), src, n * sizeof(float), cudaMemcpyDefault, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    };
    const int rK[9] = {63 + cond, 256, 256, 256, 256 + 63 + cond, 256, 256, 256, 256};      // cond = pose condition width (3 * n_bones)
    static const int sN[9] = {256, 256, 256, 205, 256, 256, 256, 256, 257};
    static const int sK[9] = {51, 256, 256, 256, 256, 256, 256, 256, 256};
    std::vector<__half> blob;
    Tc2Params& P = t.p;
    auto ident = [](int K_used, int K_pad) { std::vector<int> m(K_pad, -1); for (int k = 0; k < K_used; k++) m[k] = k; return m; };
    auto pack = [&](int L, const std::vector<float>& hw, const std::vector<float>& hb, int N_src, int K_src, const std::vector<int>& cm,
                    int Np, float scale, int pe_from, int epi) {
        P.layer[L].N = Np; P.layer[L].nchunks = (int)cm.size() / TC_KCHUNK; P.layer[L].pe_from = pe_from; P.layer[L].epi = epi;
        for (int r = 0; r < 2; r++) {
            P.layer[L].goff[r] = (unsigned)(blob.size() * 2);
            tc2_pack_rows(blob, hw, N_src, K_src, cm, r * (Np / 2), Np / 2, scale);
            P.layer[L].boff[r] = (unsigned)(blob.size() * 2);
            tc2_pack_bias(blob, hb, N_src, r * (Np / 2), Np / 2);
        }
    };
    for (int l = 0; l < 9; l++) {
        std::vector<float> hw, hb;
        int N = (l == 8) ? 3 : 256;
        if (!fetch(w->resd_w[l], (size_t)N * rK[l], hw) || !fetch(w->resd_b[l], N, hb)) { err = "tc2_upload: copy failed"; return 1; }
        std::vector<int> cm = (l == 0) ? ident(63, 64) : (l == 4 ? ident(319, 320) : ident(256, 256));
        pack(l, hw, hb, N, rK[l], cm, l == 8 ? 16 : 256, 1.f, (l == 0) ? 0 : (l == 4 ? 8 : 1 << 20), (l == 8) ? TC_EPI_RESD_FINAL : TC_EPI_RELU);
    }
    const float rs2 = (float)(1.0 / std::sqrt(2.0));
    for (int l = 0; l < 9; l++) {
        std::vector<float> hw, hb;
        if (!fetch(w->sdf_w[l], (size_t)sN[l] * sK[l], hw) || !fetch(w->sdf_b[l], sN[l], hb)) { err = "tc2_upload: copy failed"; return 1; }
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
        pack(9 + l, hw, hb, Nsrc, sK[l], cm, Np, scale, (l == 0) ? 0 : 1 << 20, (l == 8) ? TC_EPI_SDF_FINAL : (l == 3 ? TC_EPI_S3 : TC_EPI_SOFTPLUS));
    }
    tc2_free(t);
    if (cudaMalloc((void**)&t.blob, blob.size() * 2) != cudaSuccess) { err = "tc2_upload: cudaMalloc failed"; return 1; }
    cudaMemcpy(t.blob, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice);
    P.blob = t.blob;
    t.ready = true;
    return 0;
}

static void tc2_set_frame(Tc2Weights& t, const FrameConst* fc, cudaStream_t st, int64_t& launches) {
    k_tc2_pack_bias<<<1, 256, 0, st>>>(&fc->resd_b0[0], &fc->resd_b4[0], reinterpret_cast<__half*>(t.blob + t.p.layer[0].boff[0]),
                                       reinterpret_cast<__half*>(t.blob + t.p.layer[0].boff[1]),
                                       reinterpret_cast<__half*>(t.blob + t.p.layer[4].boff[0]),
                                       reinterpret_cast<__half*>(t.blob + t.p.layer[4].boff[1]));
    launches++;
}

static void tc2_distance(Tc2Weights& t, const float* bpts, float* out, const int* count, float resd_limit, int sms, cudaStream_t st,
                         int64_t& launches) {
    Tc2Params p = t.p;
    p.bpts = bpts; p.out = out; p.count = count; p.resd_limit = resd_limit; p.dbg = nullptr;
    k_mlp_tc2<<<2 * sms, TC_THREADS, TC2_SMEM_BYTES, st>>>(p);      // 148 clusters of 2 CTAs (compile-time __cluster_dims__)
    launches++;
}

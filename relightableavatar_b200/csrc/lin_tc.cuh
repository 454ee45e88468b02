// One linear layer Y = epi(X W^T + b) of the fp32 surface-attribute pass on the tcgen05 tensor cores with fp32-grade accuracy.
//
// The attribute pass (forward with saved activations + analytic input gradient + heads; base_network.py:456-494,
// relight_network.py:91-120) is ~40 plain GEMMs over the ~31 k surface samples of a frame.  On the legacy mma.sync path
// (3xTF32, mlp_simt.cuh) they run at ~120 TFLOP/s of tf32 MMAs -- the ceiling of that path on sm_100 -- 4.3 ms per frame.
// Here every fp32 operand is split into two fp16 values, x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (|error| <= 3e-8
// for |x| < 0.125, 2^-22 relative above), and the product is accumulated in fp32 TMEM as lo*hi + hi*lo + hi*hi
// (tcgen05.mma kind::f16, M=128, N<=256, K=16): the same compensation scheme as 3xTF32, on the fast tensor path.
//   smem  : one 4-stage ring, per 16-wide K chunk 16 KB of weights (pre-packed [2][Npad][8] hi image, then the lo image, by TMA)
//           and 8 KB of activations (hi / lo images [2][128][8] written by the worker warps): 96 KB, two CTAs per SM
//   warps : 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..9 = workers: load (whole 128 B lines, two groups of four chunks
//           in flight) + split X, then the fp32 epilogue (bias, ReLU / Softplus100 / derivative masks of the
//           backward pass) straight to global memory.
// Persistent: one CTA per SM walks the 128-row tiles (tile = blockIdx.x, + gridDim.x, ...); the row count is read on the device
// (no host synchronisation), CTAs without a tile exit at once.
#pragma once
#include "mlp_tc6.cuh"
#include "mlp_simt.cuh"
#include <map>
#include <tuple>

#define LT_STAGES 4
#define LT_KC 16
#define LT_W_BYTES (2 * 2 * 256 * 8 * 2)            // hi + lo weight images of a 16-wide chunk at Npad = 256: 16 KB
#define LT_A_BYTES (2 * 2 * 128 * 8 * 2)            // hi + lo activation images of a 16-wide chunk, 128 rows: 8 KB
#define LT_STAGE_BYTES (LT_W_BYTES + LT_A_BYTES)
#define LT_MAX_K 4096
#define LT_WORKERS 8
#define LT_THREADS (64 + 32 * LT_WORKERS)
#define LT_SMEM_BYTES (LT_STAGES * LT_STAGE_BYTES + 256)       // 96.25 KB: two CTAs per SM

struct LinTcArgs {
    const float* X; int ldx;
    const unsigned char* wblob;      // packed weights: K/16 chunks of (hi [2][Npad][8] halves, lo [2][Npad][8] halves)
    const float* bias;
    float* Y; int ldy;
    const float* aux; int ldaux;
    const int* count; int row0, rows_cap;
    int N, Npad, K;
};

__device__ __forceinline__ void lt_split8(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        float2 hf = __half22float2(h);
        __half2 l = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        hi[j] = *reinterpret_cast<uint32_t*>(&h);
        lo[j] = *reinterpret_cast<uint32_t*>(&l);
    }
}

// Both operands stream through one ring: per 16-wide K chunk the weights arrive by TMA, the activations are loaded (fp32,
// 32 B per row and K-group), split and stored by the worker warps while the MMAs of the previous chunks run.
template <int EPI>
__global__ void __launch_bounds__(LT_THREADS, 2) k_lin_tc(const __grid_constant__ LinTcArgs P) {
    const int total = *P.count;
    const int M = min(total - P.row0, P.rows_cap);
    if ((int)blockIdx.x * 128 >= M) return;                  // whole CTA, before any barrier / TMEM allocation
    const int n_tiles = (M + 127) / 128;                     // persistent: this CTA takes tiles blockIdx.x, + gridDim.x, ...
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base + LT_STAGES * LT_STAGE_BYTES;
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_acc = s_bar + 64, s_tmem = s_bar + 80;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - s_base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch = P.K / LT_KC;
    const uint32_t chunk_bytes = (uint32_t)P.Npad * 64u;     // hi + lo of one weight chunk

    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; s++) { mbar_init(bar_full + 8 * s, 1 + LT_WORKERS / 2); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: weight chunks =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int c = 0; c < nch; c++, it++) {
                    const uint32_t s = it & (LT_STAGES - 1), ph = (it / LT_STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    mbar_expect_tx(bar_full + 8 * s, chunk_bytes);
                    tma_bulk_g2s(s_base + s * LT_STAGE_BYTES, P.wblob + (size_t)c * chunk_bytes, chunk_bytes, bar_full + 8 * s);
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform, one elected lane) =====================
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t idesc = make_idesc_f16(P.Npad);
        const uint32_t lbo_b = (uint32_t)P.Npad * 16u;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int c = 0; c < nch; c++, it++) {
                const uint32_t s = it & (LT_STAGES - 1), ph = (it / LT_STAGES) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t b_hi = s_base + s * LT_STAGE_BYTES, b_lo = b_hi + chunk_bytes / 2;
                const uint32_t a_hi = b_hi + LT_W_BYTES, a_lo = a_hi + LT_A_BYTES / 2;
                const uint64_t ahi = make_sdesc(a_hi, 2048u, 128u), alo = make_sdesc(a_lo, 2048u, 128u);
                const uint64_t bhi = make_sdesc(b_hi, lbo_b, 128u), blo = make_sdesc(b_lo, lbo_b, 128u);
                if (elect_one()) {
                    umma_f16(tmem_u, alo, bhi, idesc, c ? 1u : 0u);
                    umma_f16(tmem_u, ahi, blo, idesc, 1u);
                    umma_f16(tmem_u, ahi, bhi, idesc, 1u);
                    umma_commit(bar_empty + 8 * s);
                    if (c == nch - 1) umma_commit(bar_acc);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== workers: activation chunks, then the epilogue =====================
        // Every global access is COALESCED: a load / store instruction of a warp covers 4 rows x 128 contiguous bytes (8 lanes per
        // row), i.e. 4 cache lines.  (The first version gave every thread its own row: each 16-byte access of a warp touched 32
        // different lines and the L1TEX tag stage -- 24.5 k line accesses per 128-row tile for X, the saved activations and Y --
        // bound the kernel at 64-78 % l1tex throughput, 2.2 TB/s of DRAM traffic whatever the row count, tensor pipe 11 %:
        // profiles/r02_ncu_k_lin_tc_before_coalescing.txt.)
        //  * X: of every group of four chunks (256 B of a row) the warps of half 0 own the first two chunks, the warps of half 1 the
        //    last two.  Request j of a warp: lane l reads the float4 `l & 7` of row 4 j + (l >> 3); the values are split into hi / lo
        //    halves in registers and go to the K-major operand images with 8-byte stores.  The next group's loads are in flight while
        //    the current group is converted.
        //  * Y: a 32 x 32 accumulator block is transposed through this warp's 4 KB of the (idle) activation half of the ring --
        //    XOR-swizzled 16-byte pieces, conflict-free both ways -- and leaves as row segments; bias / activation / derivative masks
        //    are applied on the way out with the saved activations read in the same coalesced pattern.
        const int e = warp - 2, q = warp & 3, half = e >> 2;          // half: which chunk pair of a group / which column blocks
        const int lr = lane >> 3, pc = lane & 7;                      // coalesced pattern: row-in-request, 16-byte piece of the row segment
        const uint32_t stg = s_base + (uint32_t)(e >> 1) * LT_STAGE_BYTES + LT_W_BYTES + (uint32_t)(e & 1) * 4096u;
        uint32_t it0 = 0, tc = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tc++, it0 += nch) {
            const int m0 = tile * 128 + q * 32;                        // first row of this warp's quarter
            float4 va[8], vb[8];                                       // two groups in flight: request j -> row 4 j + lr, piece pc
            auto load_group = [&](int c0, float4* dst) {
                const int c = c0 + 2 * half + (pc >> 2);               // the chunk this lane's piece belongs to
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int m = m0 + 4 * j + lr;
                    dst[j] = (m < M && c < nch) ? *reinterpret_cast<const float4*>(P.X + (size_t)m * P.ldx + (c0 + 2 * half) * LT_KC + pc * 4)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            auto store_group = [&](int c0, const float4* src) {
                const int cA = c0 + 2 * half;
                if (cA >= nch) return;
                const bool two = cA + 1 < nch;
                const uint32_t itA = it0 + (uint32_t)cA, itB = itA + 1;
                const uint32_t sA = itA & (LT_STAGES - 1), sB = itB & (LT_STAGES - 1);
                mbar_wait(bar_empty + 8 * sA, ((itA / LT_STAGES) & 1) ^ 1);
                if (two) mbar_wait(bar_empty + 8 * sB, ((itB / LT_STAGES) & 1) ^ 1);
                // piece pc: chunk (pc >> 2), K-group (pc >> 1) & 1, halves 4 (pc & 1) .. + 3 of the 16-byte operand row
                const uint32_t a_dst = s_base + ((pc >> 2) ? sB : sA) * LT_STAGE_BYTES + LT_W_BYTES + (uint32_t)((pc >> 1) & 1) * 2048u + (uint32_t)(pc & 1) * 8u;
                if (two || (pc >> 2) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 x = src[j];
                        const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
                        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                        const __half2 l0 = __floats2half2_rn(x.x - f0.x, x.y - f0.y), l1 = __floats2half2_rn(x.z - f1.x, x.w - f1.y);
                        const uint32_t a = a_dst + (uint32_t)(q * 32 + 4 * j + lr) * 16u;
                        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(*reinterpret_cast<const uint32_t*>(&h0)), "r"(*reinterpret_cast<const uint32_t*>(&h1)) : "memory");
                        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a + LT_A_BYTES / 2), "r"(*reinterpret_cast<const uint32_t*>(&l0)), "r"(*reinterpret_cast<const uint32_t*>(&l1)) : "memory");
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_full + 8 * sA); if (two) mbar_arrive(bar_full + 8 * sB); }
            };
            load_group(0, va);
            for (int c0 = 0; c0 < nch; c0 += 8) {
                if (c0 + 4 < nch) load_group(c0 + 4, vb);
                store_group(c0, va);
                if (c0 + 4 < nch) {
                    if (c0 + 8 < nch) load_group(c0 + 8, va);
                    store_group(c0 + 4, vb);
                }
            }
            // ---- epilogue, 32 accumulator columns at a time (this warp: blocks half, half + 2, ...)
            mbar_wait(bar_acc, tc & 1);
            tc_fence_after();
            const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
            const int nblk = P.Npad / 32;
            for (int cb = half; cb < nblk; cb += 2) {
                uint32_t r[32];
                tmem_ld32(t_lane + (uint32_t)(cb * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; j++)
                    st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                const int nb = cb * 32 + pc * 4;                       // this lane's 4 output columns
                const bool full4 = nb + 3 < P.N;
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (P.bias) {
#pragma unroll
                    for (int t = 0; t < 4; t++) if (nb + t < P.N) bv[t] = __ldg(&P.bias[nb + t]);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int R = 4 * j + lr, m = m0 + R;
                    uint32_t w0, w1, w2, w3;
                    ld_shared_v4(stg + (uint32_t)R * 128u + (uint32_t)((pc ^ (R & 7)) << 4), w0, w1, w2, w3);
                    if (m >= M || nb >= P.N) continue;
                    float o[4] = {__uint_as_float(w0), __uint_as_float(w1), __uint_as_float(w2), __uint_as_float(w3)};
                    float au[4] = {0.f, 0.f, 0.f, 0.f};
                    if (EPI == EPI_MUL_DRELU || EPI == EPI_MUL_DSOFTPLUS) {
                        const float* ar = P.aux + (size_t)m * P.ldaux + nb;
                        if (full4) { const float4 a4 = *reinterpret_cast<const float4*>(ar); au[0] = a4.x; au[1] = a4.y; au[2] = a4.z; au[3] = a4.w; }
                        else
                            for (int t = 0; t < 4; t++) if (nb + t < P.N) au[t] = ar[t];
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        float v = o[t] + bv[t];
                        if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
                        if (EPI == EPI_SOFTPLUS) v = softplus100(v);
                        if (EPI == EPI_MUL_DRELU) v = (au[t] > 0.f) ? v : 0.f;
                        if (EPI == EPI_MUL_DSOFTPLUS) v *= dsoftplus100_from_act(au[t]);
                        o[t] = v;
                    }
                    float* yr = P.Y + (size_t)m * P.ldy + nb;
                    if (full4) *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[2], o[3]);
                    else
                        for (int t = 0; t < 4; t++) if (nb + t < P.N) yr[t] = o[t];
                }
                __syncwarp();                    // the staging block is reused by the next column block
            }
            // all accumulator reads and staging traffic of this tile are done before ANY worker feeds the next tile (the activation
            // halves of the ring double as staging; the next tile's first MMA overwrites the accumulator)
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LT_WORKERS) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// ------------------------------------------------------------------------------------------ host side
struct LinTcWeights {
    std::map<std::tuple<const float*, int, int, int>, unsigned char*> cache;     // (W, ldw, N, K) -> packed device blob
};

static void lin_tc_clear(LinTcWeights& w) {
    for (auto& kv : w.cache) cudaFree(kv.second);
    w.cache.clear();
}

// pack on first use (the weight matrices are static between ra_upload_weights calls); returns null on failure
static const unsigned char* lin_tc_pack(LinTcWeights& w, const float* W, int ldw, int N, int K, int Npad, cudaStream_t st) {
    auto key = std::make_tuple(W, ldw, N, K);
    auto it = w.cache.find(key);
    if (it != w.cache.end()) return it->second;
    std::vector<float> hw((size_t)N * ldw);
    if (cudaStreamSynchronize(st) != cudaSuccess) return nullptr;
    if (cudaMemcpy(hw.data(), W, hw.size() * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return nullptr;
    const int nch = K / LT_KC;
    std::vector<__half> blob((size_t)nch * Npad * 32, __float2half(0.f));       // per chunk: hi [2][Npad][8], lo [2][Npad][8]
    for (int c = 0; c < nch; c++)
        for (int kq = 0; kq < 2; kq++)
            for (int n = 0; n < N; n++)
                for (int j = 0; j < 8; j++) {
                    const int k = c * LT_KC + kq * 8 + j;
                    const float v = hw[(size_t)n * ldw + k];
                    const __half hi = __float2half_rn(v);
                    const __half lo = __float2half_rn(v - __half2float(hi));
                    const size_t base = (size_t)c * Npad * 32;
                    blob[base + ((size_t)kq * Npad + n) * 8 + j] = hi;
                    blob[base + (size_t)Npad * 16 + ((size_t)kq * Npad + n) * 8 + j] = lo;
                }
    unsigned char* d = nullptr;
    if (cudaMalloc((void**)&d, blob.size() * 2) != cudaSuccess) return nullptr;
    cudaMemcpy(d, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice);
    cudaStreamSynchronize(cudaStreamLegacy);      // pageable H2D copies return once staged and run on the legacy stream: a non-blocking caller stream is not ordered behind them
    w.cache[key] = d;
    return d;
}

static inline bool lin_tc_ok(int N, int K, int ldx) { return N >= 1 && N <= 256 && K >= LT_KC && K <= LT_MAX_K && (K % LT_KC) == 0 && (ldx % 4) == 0; }

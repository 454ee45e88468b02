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
//   warps : 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..9 = workers: load + split X chunk by chunk (4 chunks of
//           global loads in flight per thread), then the fp32 epilogue (bias, ReLU / Softplus100 / derivative masks of the
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
#define LT_PRE 4                                     // chunks whose global loads a worker keeps in flight

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
        for (int s = 0; s < LT_STAGES; s++) { mbar_init(bar_full + 8 * s, 1 + LT_WORKERS); mbar_init(bar_empty + 8 * s, 1); }
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
        const int e = warp - 2, q = warp & 3, half = e >> 2;          // half: which K-group of a chunk / which column blocks
        const int row = q * 32 + lane;
        uint32_t it = 0, tc = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tc++) {
            const int m = tile * 128 + row;
            const bool rok = m < M;
            const float* xr = P.X + (size_t)(rok ? m : 0) * P.ldx + half * 8;
            for (int c0 = 0; c0 < nch; c0 += LT_PRE) {
                float v[LT_PRE][8];
#pragma unroll
                for (int u = 0; u < LT_PRE; u++) {              // LT_PRE chunks' loads in flight
                    if (rok && c0 + u < nch) {
                        const float4 a = *reinterpret_cast<const float4*>(xr + (c0 + u) * LT_KC), b = *reinterpret_cast<const float4*>(xr + (c0 + u) * LT_KC + 4);
                        v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++) v[u][j] = 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < LT_PRE; u++) {
                    if (c0 + u >= nch) break;
                    const uint32_t s = it & (LT_STAGES - 1), ph = (it / LT_STAGES) & 1;
                    it++;
                    uint32_t hi[4], lo[4];
                    lt_split8(v[u], hi, lo);
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t a_hi = s_base + s * LT_STAGE_BYTES + LT_W_BYTES + (uint32_t)half * 2048u + (uint32_t)row * 16u;
                    st_shared_v4(a_hi, hi[0], hi[1], hi[2], hi[3]);
                    st_shared_v4(a_hi + LT_A_BYTES / 2, lo[0], lo[1], lo[2], lo[3]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * s);
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
                if (!rok) continue;
                const int n0 = cb * 32;
                float* yr = P.Y + (size_t)m * P.ldy + n0;
                const float* ar = (EPI == EPI_MUL_DRELU || EPI == EPI_MUL_DSOFTPLUS) ? P.aux + (size_t)m * P.ldaux + n0 : nullptr;
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++) {
                    float o[4];
                    const int nb = n0 + j4 * 4;
                    float4 av = make_float4(0, 0, 0, 0);
                    const bool full4 = nb + 3 < P.N;
                    if (ar && full4) av = *reinterpret_cast<const float4*>(ar + j4 * 4);
                    const float avs[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int n = nb + j;
                        float v = __uint_as_float(r[j4 * 4 + j]);
                        if (n < P.N) {
                            if (P.bias) v += __ldg(&P.bias[n]);
                            float au = 0.f;
                            if (ar) au = full4 ? avs[j] : ar[j4 * 4 + j];
                            if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
                            if (EPI == EPI_SOFTPLUS) v = softplus100(v);
                            if (EPI == EPI_MUL_DRELU) v = (au > 0.f) ? v : 0.f;
                            if (EPI == EPI_MUL_DSOFTPLUS) v *= dsoftplus100_from_act(au);
                        }
                        o[j] = v;
                    }
                    if (full4) *reinterpret_cast<float4*>(yr + j4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
                    else
                        for (int j = 0; j < 4; j++)
                            if (nb + j < P.N) yr[j4 * 4 + j] = o[j];
                }
            }
            tc_fence_before();                   // accumulator reads done before this warp feeds the next tile's first chunk
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
    w.cache[key] = d;
    return d;
}

static inline bool lin_tc_ok(int N, int K, int ldx) { return N >= 1 && N <= 256 && K >= LT_KC && K <= LT_MAX_K && (K % LT_KC) == 0 && (ldx % 4) == 0; }

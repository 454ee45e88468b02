// C-ABI entry points (include/ra_b200.h) and host-side orchestration of the kernels.
// One handle per device.  The product path enqueues a fixed sequence of kernels on the caller's
// stream; every data-dependent size stays in device counters (no host sync) except in the
// RA_PRECISION_FP32 debug/reference-precision mode and the 128-sample volume renderer, which chunk
// their MLP work by a counter read back from the device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <unordered_set>
#include <dlfcn.h>
#include <cmath>

#include "../../include/ra_b200.h"
#include "common.cuh"
#include "hdq.cuh"
#include "mlp_simt.cuh"
#include "render.cuh"
#include "ground.cuh"
#include "prep.cuh"
#include "mlp_tc.cuh"
#include "mlp_tc2.cuh"
#include "mlp_tc6.cuh"
#include "mlp_tc7.cuh"
#include "mlp_tc8.cuh"
#include "lin_tc.cuh"
#include "visual.cuh"

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + std::to_string(__LINE__); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define LAUNCH(h, kern, grid, block, smem, stream, ...)          \
    do {                                                         \
        kern<<<grid, block, smem, stream>>>(__VA_ARGS__);        \
        (h)->launches++;                                         \
    } while (0)

static const int ATTR_CH = 262144;   // rows per fp32 MLP chunk
static const int64_t RA_GROUND_SLOTS_MAX = 96ll << 20;      // entries of the floor pass's ray list (x 50 B: 4.7 GB): a whole 512^2 image (67 M); larger images go in batches

struct Lin {            // one fp32 linear layer, K zero-padded to a multiple of 8
    float* w = nullptr; float* b = nullptr; int N = 0, K = 0;     // (N, K) row-major
    float* wt = nullptr; int Nt = 0, Kt = 0;                      // transposed (K_in rows, N padded to 8) for backward
};

struct ra_handle {
    ra_config cfg;
    std::unordered_set<void*> allocs;   // every device block the handle owns through dalloc() -- released by ra_destroy
    std::string err;
    int64_t launches = 0;
    int dev = 0, sms = 148;
    int tb_surf = 256, tb_shadow = 256;      // tracing-kernel block sizes (env RA_TB_SURF / RA_TB_SHADOW)
    float cell_h = 0.035f, grid2_ratio = 4.0f;  // tunables (env RA_CELL_H / RA_GRID2_RATIO); swept on B200 with the list-based 3-NN: .03 / .035 / .04 / .05 -> 23.8 / 23.7 / 24.3 / 24.7 ms per frame
    bool have_weights = false, have_frame = false;
    // ---- weights
    Lin resd[9], sdf[9], rend[5], alb[3], rgh[3];
    float *resd_w0_raw = nullptr, *resd_b0_raw = nullptr, *resd_w4_raw = nullptr, *resd_b4_raw = nullptr;
    float *rend_w3_raw = nullptr, *rend_b3_raw = nullptr;
    float* w4_skipT = nullptr;       // resd WT4 rows 256..319 live inside resd[4].wt; helper pointers below
    float beta = 0.1f;
    float *env_main = nullptr; int emh = 0, emw = 0;
    float* main_light = nullptr; int mlh = 0, mlw = 0;      // cfg.replace_light: env-map that lights the main pass instead of the learned one (ra_set_main_light)
    float *lxyz = nullptr, *larea = nullptr, *lsharp = nullptr, *ldir = nullptr;
    TcWeights tc;                    // fp16 UMMA images for the tcgen05 path
    Tc2Weights tc2;                  // ... and for its 2-CTA (cta_group::2) variant
    Tc8Weights tc8;                  // ... and for k_mlp_tc8 (A operand in tensor memory: per (layer, half, rank) images)
    int attr_tc = 3;                 // env RA_ATTR_TC: 3 = tcgen05 fp16-split GEMM (lin_tc.cuh), 2 = pipelined 3xTF32 mma.sync GEMM, 1 = first 3xTF32 GEMM, 0 = CUDA-core SGEMM
    LinTcWeights lin_tc;             // packed hi / lo weight images of the attribute-pass GEMMs (built on first use)
    int tc_variant = 6;              // env RA_TC_VARIANT: 6 = CTA-pair two-slot kernel k_mlp_tc6 (default); 7 = k_mlp_tc7 (tc6 + output layers with A in tensor memory + skewed slots: measured no faster, kept as the tcgen05.st / TS-MMA cross-check); 8 = k_mlp_tc8 (A of every hidden layer in tensor memory, N in two halves: bit-identical, 41 % slower -- half-width MMAs are A-operand-bound); 1 = single-CTA kernel k_mlp_tc; 2 = first CTA-pair kernel k_mlp_tc2 (all bit-identical)
    // ---- frame
    FrameConst* fc = nullptr;
    SortedVerts sv{};
    int *cell_count = nullptr, *cell_fill = nullptr, *vert_cell = nullptr, *vert_order = nullptr;
    int* nb_cnt = nullptr;           // neighbourhood-list counts, RA_NB_LEVELS x (RA_MAX_CELLS + 1)
    ra_frame frame{};
    // ---- per-render workspace
    int64_t P_cap = 0, q_cap = 0, attr_cap = 0, vol_rays = 8192;
    SurfState ss{};
    float *surf = nullptr, *acc = nullptr, *depth = nullptr;
    int* fg_ray = nullptr;
    FgMaps fm{};
    float *lvis = nullptr, *ldot = nullptr;
    ShadowRays sr{};
    QueryList q{}, q2{};             // q2: second work list for the overlapped half of the shadow rays
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int pkt_order = 1, pkt_search = 5;       // shadow rays generated as packets (same light, 32 neighbouring pixels): bit 0 floor pass, bit 1 human pass; far-field 3-NN per packet (bit 2: surface rays and volume samples, coherent as they are) (env RA_PKT_ORDER, RA_PKT_SEARCH)
    int final_skip = 1;              // shadow rays whose state is a fixed point are skipped by later tracer launches (env RA_TRACE_FINAL=0: off, for the bit-identity test)
    int overlap = 0;                 // env RA_OVERLAP=1: split the shadow stage over two streams (experiment: -0.3..0.5 ms with k_mlp_tc6, but the co-running
                                     // tracing warps slow the MLP epilogue and blur the per-kernel timing; capping MLP registers for more co-residency lost more than it gained)
    AttrList al{};
    float* raw = nullptr;
    Counters cnt{};
    int* counters_blk = nullptr;     // n_fg, n_shadow, n_attr, q.count, (pad), n_queries(ull), n_inshell(ull)
    float *pt_smpl = nullptr; int* pt_slot = nullptr;     // ra_query_sdf scratch
    float* bg_spec = nullptr;
    int* blk_cnt = nullptr; int64_t blk_cap = 0;      // image assembly scratch
    ShadowRays sr_g{}; QueryList q_g{}; int64_t g_slots = 0;      // floor pass: its own ray list / query list, sized for a whole image where that fits (ra_render_ground)
    int g_W = 0, g_H = 0;             // image size of the last ra_ground_begin (the floor's 8 x 4 pixel packets)
    int* pix2ray = nullptr; int64_t pix_cap = 0;      // ground pass: image pixel -> ray (ra_ground_begin)
    float *g_weight = nullptr, *g_light = nullptr;
    KthState* kth = nullptr;          // two order-statistic states (f3: Depth / Shading / Specular / Residual percentiles)
    BodyDev body; int* prep_mm = nullptr; float* prep_nacc = nullptr;      // f1: uploaded body, bounds scratch, normal accumulator    // ground pass: far-field blend weight (F), per-light radiance table (L,3)
    // ---- fp32 MLP chunk buffers
    float *Xr0, *ra_[8], *Xr4, *z8, *resd_o, *cpts_o, *Xs0, *sb_[8], *Xs4, *out257, *GA, *GB, *dpe0, *dpes, *gcp, *u4, *gbp, *nrm_o;
    float *hd1, *hd2, *head_a, *head_r, *Xrn, *rn1, *rn2;
    // profiling (bench.py): event pairs around the MLP kernel + stage marks
    bool prof = false;
    std::vector<cudaEvent_t> ev_mlp;      // pairs
    size_t ev_mlp_used = 0;
    std::vector<cudaEvent_t> ev_stage;    // 5 marks per render
    size_t ev_stage_used = 0;
    // last render
    int64_t last_P = 0; const float* last_ray_o = nullptr; bool have_render = false; int chunk_actual = 1;
    int64_t lay_global_P = 0; int lay_block = 32, lay_world = 1, lay_rank = 0;     // ra_set_ray_layout (tile sharding)
};

// ---------------------------------------------------------------------------------------------- helpers
static int grid_for(ra_handle* h, long long n, int block = 256, int per_sm = 8) {
    long long g = (n + block - 1) / block;
    long long cap = (long long)h->sms * per_sm;
    return (int)std::max(1LL, std::min(g, cap));
}

template <typename T>
static cudaError_t dalloc(ra_handle* h, T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e == cudaSuccess) { h->allocs.insert((void*)*p); e = cudaMemset(*p, 0, n * sizeof(T)); }
    // cudaMemset of device memory returns before the fill has run, and it runs on the legacy default stream: a caller that works on a
    // NON-BLOCKING stream (torch side streams: parallel.FramesInFlight) would not be ordered behind it -- the first kernels of a fresh
    // handle could be overtaken by its own zero-fill.  Allocation happens at creation and on capacity growth only: wait here.
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
}
template <typename T>
static void hfree(ra_handle* h, T*& p) {
    if (p) { h->allocs.erase((void*)p); cudaFree((void*)p); p = nullptr; }
}

// copy a (N, K) fp32 matrix (host or device) into a zero-padded (N, Kp) device buffer, optionally selecting columns
static int upload_lin(ra_handle* h, Lin& L, const float* w, const float* b, int N, int K_src, int col0, int K_use, int Kp,
                      float scale, const int* row_perm, cudaStream_t st) {
    std::vector<float> hw((size_t)N * K_src), hb(N);
    CK(cudaMemcpyAsync(hw.data(), w, hw.size() * sizeof(float), cudaMemcpyDefault, st));
    if (b) CK(cudaMemcpyAsync(hb.data(), b, hb.size() * sizeof(float), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    std::vector<float> pw((size_t)N * Kp, 0.f), pb(N, 0.f);
    for (int n = 0; n < N; n++) {
        int sn = row_perm ? row_perm[n] : n;
        for (int k = 0; k < K_use; k++) pw[(size_t)n * Kp + k] = hw[(size_t)sn * K_src + col0 + k] * scale;
        pb[n] = b ? hb[sn] : 0.f;
    }
    int Np8 = (N + 7) / 8 * 8;
    std::vector<float> pt((size_t)Kp * Np8, 0.f);
    for (int n = 0; n < N; n++)
        for (int k = 0; k < Kp; k++) pt[(size_t)k * Np8 + n] = pw[(size_t)n * Kp + k];
    hfree(h, L.w); hfree(h, L.b); hfree(h, L.wt);
    CK(dalloc(h, &L.w, pw.size())); CK(dalloc(h, &L.b, pb.size())); CK(dalloc(h, &L.wt, pt.size()));
    CK(cudaMemcpy(L.w, pw.data(), pw.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.b, pb.data(), pb.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.wt, pt.data(), pt.size() * sizeof(float), cudaMemcpyHostToDevice));
    L.N = N; L.K = Kp; L.Nt = Kp; L.Kt = Np8;
    return 0;
}

static int upload_raw(ra_handle* h, float** dst, const float* src, size_t n, cudaStream_t st) {
    hfree(h, *dst);
    CK(dalloc(h, dst, n));
    CK(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

template <int EPI>
static void gemm(ra_handle* h, cudaStream_t st, const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y,
                 int ldy, const float* aux, int ldaux, const int* count, int row0, int rows_cap, int N, int K) {
    GemmArgs a{X, ldx, W, ldw, bias, Y, ldy, aux, ldaux, count, row0, rows_cap, N, K};
    dim3 grid((rows_cap + GBM - 1) / GBM, (N + GBN - 1) / GBN);
    if (h->attr_tc == 3 && lin_tc_ok(N, K, ldx)) {
        const int Npad = (N + 31) / 32 * 32;
        const unsigned char* blob = lin_tc_pack(h->lin_tc, W, ldw, N, K, Npad, st);
        if (blob) {
            LinTcArgs la{X, ldx, blob, bias, Y, ldy, aux, ldaux, count, row0, rows_cap, N, Npad, K};
            LAUNCH(h, k_lin_tc<EPI>, std::min((rows_cap + 127) / 128, 2 * h->sms), LT_THREADS, LT_SMEM_BYTES, st, la);
            return;
        }
    }
    if (h->attr_tc >= 2 && (K % G2K) == 0) {
        LAUNCH(h, k_gemm_tf32x3_p<EPI>, grid, 256, G2_SMEM_BYTES, st, a);
    } else if (h->attr_tc) LAUNCH(h, k_gemm_tf32x3<EPI>, grid, 256, 0, st, a);
    else LAUNCH(h, k_gemm<EPI>, grid, 256, 0, st, a);
}

// Dynamic shared-memory opt-in of every GEMM instantiation.  cudaFuncSetAttribute applies to the CURRENT device only, so this runs in
// every ra_create (one handle per device; a process may hold handles on several GPUs), like tc_init / tc2_init / tc6_init.
template <int EPI>
static cudaError_t gemm_attr_one() {
    cudaError_t e = cudaFuncSetAttribute(k_lin_tc<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_gemm_tf32x3_p<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES);
}
static int gemm_init(std::string& err) {
    cudaError_t e = gemm_attr_one<EPI_NONE>();
    if (e == cudaSuccess) e = gemm_attr_one<EPI_RELU>();
    if (e == cudaSuccess) e = gemm_attr_one<EPI_SOFTPLUS>();
    if (e == cudaSuccess) e = gemm_attr_one<EPI_MUL_DRELU>();
    if (e == cudaSuccess) e = gemm_attr_one<EPI_MUL_DSOFTPLUS>();
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute(attribute GEMMs): ") + cudaGetErrorString(e); return 1; }
    return 0;
}

// ---------------------------------------------------------------------------------------------- create / destroy
extern "C" int ra_create(ra_handle** out, const ra_config* cfg) {
    if (!out || !cfg) return 1;
    ra_handle* h = new ra_handle();
    *out = h;
    h->cfg = *cfg;
    if (cfg->n_verts <= 0 || cfg->n_bones <= 0 || cfg->max_rays <= 0) { h->err = "bad config"; return 1; }
    if (cfg->env_h * cfg->env_w > RA_NLIGHT_MAX) { h->err = "env_h*env_w > 512"; return 1; }
    CK(cudaGetDevice(&h->dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->dev));
    h->sms = prop.multiProcessorCount;
    if (const char* e = getenv("RA_CELL_H")) h->cell_h = (float)atof(e);
    if (const char* e = getenv("RA_TB_SURF")) h->tb_surf = std::min(256, std::max(32, atoi(e) / 32 * 32));
    if (const char* e = getenv("RA_TB_SHADOW")) h->tb_shadow = std::min(256, std::max(32, atoi(e) / 32 * 32));
    if (const char* e = getenv("RA_GRID2_RATIO")) h->grid2_ratio = (float)atof(e);
    if (prop.major != 10) { h->err = "ra_b200 requires an sm_100 (B200) device"; return 1; }
    int64_t P = cfg->max_rays;
    int L = cfg->env_h * cfg->env_w;
    h->P_cap = P;
    h->q_cap = std::max<int64_t>(256 * P, 1 << 20);
    h->attr_cap = std::max<int64_t>((int64_t)cfg->n_samples * P, h->vol_rays * cfg->vol_samples);
    int N = cfg->n_verts;
    CK(dalloc(h, &h->fc, 1));
    CK(dalloc(h, &h->sv.pos, N)); CK(dalloc(h, &h->sv.nrm, N)); CK(dalloc(h, &h->sv.tv, N)); CK(dalloc(h, &h->sv.T, (size_t)N * 24));
    CK(dalloc(h, &h->sv.cell_start, RA_MAX_CELLS + 1));
    CK(dalloc(h, &h->sv.pos2, N)); CK(dalloc(h, &h->sv.cell_start2, RA_MAX_CELLS + 1));
    CK(dalloc(h, &h->sv.occ_lo, RA_MAX_OCC)); CK(dalloc(h, &h->sv.occ_hi, RA_MAX_OCC));
    CK(dalloc(h, &h->sv.sup_lo, RA_MAX_SUP)); CK(dalloc(h, &h->sv.sup_hi, RA_MAX_SUP));
    CK(dalloc(h, &h->sv.occ_tmp, 2 * RA_MAX_OCC)); CK(dalloc(h, &h->sv.occ_sup, RA_MAX_OCC));
    {
        for (int lv = 0; lv < RA_NB_LEVELS; lv++) {
            const int r = lv + 1;          // cells per level = how often a vertex can appear: 27, then the shells 98, 218, 386, ...
            const int shell = (r == 1) ? 27 : (2 * r + 1) * (2 * r + 1) * (2 * r + 1) - (2 * r - 1) * (2 * r - 1) * (2 * r - 1);
            CK(dalloc(h, &h->sv.nb_start[lv], RA_MAX_CELLS + 1));
            CK(dalloc(h, &h->sv.nb_pos[lv], (size_t)shell * N));
        }
        CK(dalloc(h, &h->nb_cnt, (size_t)RA_NB_LEVELS * (RA_MAX_CELLS + 1)));
        CK(dalloc(h, &h->sv.nb_mask, RA_MAX_CELLS));
    }
    CK(dalloc(h, &h->cell_count, RA_MAX_CELLS + 1)); CK(dalloc(h, &h->cell_fill, RA_MAX_CELLS + 1)); CK(dalloc(h, &h->vert_cell, N)); CK(dalloc(h, &h->vert_order, N));
    float** ssp[] = {&h->ss.t, &h->ss.occ, &h->ss.d0, &h->ss.cd, &h->ss.dt, &h->ss.st, &h->ss.off, &h->ss.rlx, &h->ss.q_smpl};
    for (auto p : ssp) CK(dalloc(h, p, P));
    CK(dalloc(h, &h->ss.q_slot, P));
    CK(dalloc(h, &h->surf, 3 * P)); CK(dalloc(h, &h->acc, P)); CK(dalloc(h, &h->depth, P)); CK(dalloc(h, &h->fg_ray, P));
    CK(dalloc(h, &h->fm.norm, 3 * P)); CK(dalloc(h, &h->fm.albedo, 3 * P)); CK(dalloc(h, &h->fm.rough, P));
    if (cfg->relight) {
        CK(dalloc(h, &h->lvis, (size_t)P * L)); CK(dalloc(h, &h->ldot, (size_t)P * L));
        int64_t S = 256 * P;
        CK(dalloc(h, &h->sr.fg, S)); CK(dalloc(h, &h->sr.light, S)); CK(dalloc(h, &h->sr.near_, S)); CK(dalloc(h, &h->sr.far_, S));
        h->sr.cap = (int)std::min<int64_t>(S, 0x7fffffff);
        CK(dalloc(h, &h->sr.t, S)); CK(dalloc(h, &h->sr.occ, S)); CK(dalloc(h, &h->sr.d0, S)); CK(dalloc(h, &h->sr.q_smpl, S)); CK(dalloc(h, &h->sr.q_slot, S));
    }
    CK(dalloc(h, &h->q.bpts, (size_t)h->q_cap * 3)); CK(dalloc(h, &h->q.net, (size_t)h->q_cap));
    if (cfg->relight) { CK(dalloc(h, &h->q2.bpts, (size_t)h->q_cap * 3 / 2 + 3)); CK(dalloc(h, &h->q2.net, (size_t)h->q_cap / 2 + 1)); }
    CK(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    if (const char* e = getenv("RA_OVERLAP")) h->overlap = atoi(e);
    if (const char* e = getenv("RA_TRACE_FINAL")) h->final_skip = atoi(e);
    if (const char* e = getenv("RA_PKT_ORDER")) h->pkt_order = atoi(e);
    if (const char* e = getenv("RA_PKT_SEARCH")) h->pkt_search = atoi(e);
    if (const char* e = getenv("RA_PKT_MIN")) { int v = atoi(e); CK(cudaMemcpyToSymbol(g_pkt_min, &v, sizeof(int))); }
    if (const char* e = getenv("RA_PKT_RHO")) { float v = (float)atof(e); CK(cudaMemcpyToSymbol(g_pkt_rho, &v, sizeof(float))); }
    CK(dalloc(h, &h->al.bpts, (size_t)h->attr_cap * 3)); CK(dalloc(h, &h->al.mats, (size_t)h->attr_cap * 18));
    CK(dalloc(h, &h->al.bvds, (size_t)h->attr_cap * 3)); CK(dalloc(h, &h->al.src, (size_t)h->attr_cap));
    CK(dalloc(h, &h->raw, (size_t)h->attr_cap * 17));
    CK(dalloc(h, &h->counters_blk, 16));
    h->cnt.n_fg = h->counters_blk; h->cnt.n_shadow = h->counters_blk + 1; h->cnt.n_attr = h->counters_blk + 2;
    h->q.count = h->counters_blk + 3; h->al.count = h->cnt.n_attr;
    h->q2.count = h->counters_blk + 4;
    h->sr.dropped = h->counters_blk + 6;          // (+5: shadow-ray list counter of the floor pass)
    h->sr.n_rays = h->counters_blk + 7;           // rays appended by the human pass and the floor pass (the list counters count padded packet slots)
    h->cnt.n_queries = (unsigned long long*)(h->counters_blk + 8); h->cnt.n_inshell = (unsigned long long*)(h->counters_blk + 10);
    CK(dalloc(h, &h->pt_smpl, (size_t)h->q_cap)); CK(dalloc(h, &h->pt_slot, (size_t)h->q_cap));
    CK(dalloc(h, &h->bg_spec, 4));
    size_t R = ATTR_CH;
    CK(dalloc(h, &h->Xr0, R * 64)); for (int i = 0; i < 8; i++) CK(dalloc(h, &h->ra_[i], R * 256));
    CK(dalloc(h, &h->Xr4, R * 320)); CK(dalloc(h, &h->z8, R * 4)); CK(dalloc(h, &h->resd_o, R * 3)); CK(dalloc(h, &h->cpts_o, R * 3));
    CK(dalloc(h, &h->Xs0, R * 64)); for (int i = 0; i < 8; i++) CK(dalloc(h, &h->sb_[i], R * 256));
    CK(dalloc(h, &h->Xs4, R * 256)); CK(dalloc(h, &h->out257, R * 264));
    CK(dalloc(h, &h->GA, R * 256)); CK(dalloc(h, &h->GB, R * 256)); CK(dalloc(h, &h->dpe0, R * 64)); CK(dalloc(h, &h->dpes, R * 64));
    CK(dalloc(h, &h->gcp, R * 3)); CK(dalloc(h, &h->u4, R * 4)); CK(dalloc(h, &h->gbp, R * 3)); CK(dalloc(h, &h->nrm_o, R * 3));
    CK(dalloc(h, &h->hd1, R * 128)); CK(dalloc(h, &h->hd2, R * 128)); CK(dalloc(h, &h->head_a, R * 4)); CK(dalloc(h, &h->head_r, R * 4));
    CK(dalloc(h, &h->Xrn, R * 288)); CK(dalloc(h, &h->rn1, R * 256)); CK(dalloc(h, &h->rn2, R * 256));
    if (tc_init(h->tc, h->err)) return 1;
    if (tc2_init(h->tc2, h->err)) return 1;
    if (tc6_init(h->err)) return 1;
    if (tc7_init(h->err)) return 1;
    if (tc8_init(h->err)) return 1;
    if (gemm_init(h->err)) return 1;
    if (const char* e = getenv("RA_TC_VARIANT")) h->tc_variant = atoi(e);
    if (const char* e = getenv("RA_TC_SKEW")) h->tc2.skew = std::max(0, std::min(17, atoi(e)));
    if (h->tc_variant != 1 && h->tc_variant != 2 && (h->tc_variant < 6 || h->tc_variant > 8)) { h->err = "RA_TC_VARIANT must be 1, 2, 6, 7 or 8"; return 1; }
    if (const char* e = getenv("RA_ATTR_TC")) h->attr_tc = atoi(e);
    return 0;
}

extern "C" void ra_destroy(ra_handle* h) {
    if (!h) return;
    for (void* p : h->allocs) cudaFree(p);          // everything allocated through dalloc()
    h->allocs.clear();
    tc_free(h->tc);
    tc2_free(h->tc2);
    tc8_free(h->tc8);
    lin_tc_clear(h->lin_tc);
    if (h->tc.dbg) cudaFree(h->tc.dbg);
    if (h->aux) cudaStreamDestroy(h->aux);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (cudaEvent_t e : h->ev_mlp) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_stage) cudaEventDestroy(e);
    delete h;
}

extern "C" const char* ra_last_error(ra_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t ra_launch_count(ra_handle* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------------------------- weights
extern "C" int ra_upload_weights(ra_handle* h, const ra_weights* w, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const float rs2 = (float)(1.0 / std::sqrt(2.0));
    // residual deformation MLP: (63+C) -> 256 x4 -> (256+63+C) -> 256 x3 -> 3; the C = 3 * n_bones pose condition (156 for SMPL-H,
    // 72 for SMPL; config.py:465-466) is folded into biases per frame
    const int C = 3 * h->cfg.n_bones;
    const int rK[9] = {63 + C, 256, 256, 256, 256 + 63 + C, 256, 256, 256, 256};
    for (int l = 0; l < 9; l++) {
        int N = (l == 8) ? 3 : 256;
        if (l == 0) { if (upload_lin(h, h->resd[l], w->resd_w[l], w->resd_b[l], N, rK[l], 0, 63, 64, 1.f, nullptr, st)) return 1; }
        else if (l == 4) {
            // columns [h3 (256) | PE10 (63)] -> (256, 320)
            if (upload_lin(h, h->resd[l], w->resd_w[l], w->resd_b[l], N, rK[l], 0, 256 + 63, 320, 1.f, nullptr, st)) return 1;
        } else if (upload_lin(h, h->resd[l], w->resd_w[l], w->resd_b[l], N, rK[l], 0, rK[l], 256, 1.f, nullptr, st)) return 1;
    }
    if (upload_raw(h, &h->resd_w0_raw, w->resd_w[0], (size_t)256 * rK[0], st)) return 1;
    if (upload_raw(h, &h->resd_b0_raw, w->resd_b[0], 256, st)) return 1;
    if (upload_raw(h, &h->resd_w4_raw, w->resd_w[4], (size_t)256 * rK[4], st)) return 1;
    if (upload_raw(h, &h->resd_b4_raw, w->resd_b[4], 256, st)) return 1;
    // SDF MLP: 51 -> 256,256,256,205 -> cat(205,51)/sqrt2 -> 256 x3 -> 257 ; rows of the last layer permuted to [feat(256), sdf]
    static const int sN[9] = {256, 256, 256, 205, 256, 256, 256, 256, 257};
    static const int sK[9] = {51, 256, 256, 256, 256, 256, 256, 256, 256};
    std::vector<int> perm(257);
    for (int i = 0; i < 256; i++) perm[i] = i + 1;
    perm[256] = 0;
    for (int l = 0; l < 9; l++) {
        int Kp = (l == 0) ? 64 : 256;
        if (upload_lin(h, h->sdf[l], w->sdf_w[l], w->sdf_b[l], sN[l], sK[l], 0, sK[l], Kp, l == 4 ? rs2 : 1.f,
                       l == 8 ? perm.data() : nullptr, st)) return 1;
    }
    h->beta = w->sdf_beta;
    if (w->render_w[0]) {
        const int nK[5] = {286, 256, 256, 256 + C, 256};
        for (int l = 0; l < 5; l++) {
            int N = (l == 4) ? 3 : 256;
            int use = (l == 3) ? 256 : nK[l];
            int Kp = (l == 0) ? 288 : 256;
            if (upload_lin(h, h->rend[l], w->render_w[l], w->render_b[l], N, nK[l], 0, use, Kp, 1.f, nullptr, st)) return 1;
        }
        if (upload_raw(h, &h->rend_w3_raw, w->render_w[3], (size_t)256 * nK[3], st)) return 1;
        if (upload_raw(h, &h->rend_b3_raw, w->render_b[3], 256, st)) return 1;
    }
    if (h->cfg.relight) {
        if (!w->albedo_w[0] || !w->rough_w[0] || !w->env_main || !w->light_xyz) { h->err = "relight weights missing"; return 1; }
        static const int aK[3] = {256, 128, 128};
        for (int l = 0; l < 3; l++) {
            if (upload_lin(h, h->alb[l], w->albedo_w[l], w->albedo_b[l], l == 2 ? 3 : 128, aK[l], 0, aK[l], aK[l], 1.f, nullptr, st)) return 1;
            if (upload_lin(h, h->rgh[l], w->rough_w[l], w->rough_b[l], l == 2 ? 1 : 128, aK[l], 0, aK[l], aK[l], 1.f, nullptr, st)) return 1;
        }
        h->emh = w->env_main_h; h->emw = w->env_main_w;
        if (upload_raw(h, &h->env_main, w->env_main, (size_t)h->emh * h->emw * 3, st)) return 1;
        int L = h->cfg.env_h * h->cfg.env_w;
        if (upload_raw(h, &h->lxyz, w->light_xyz, (size_t)L * 3, st)) return 1;
        if (upload_raw(h, &h->larea, w->light_area, L, st)) return 1;
        if (upload_raw(h, &h->lsharp, w->light_sharp, L, st)) return 1;
        // shadow ray directions: normalize(xyz) with the reference's normalize (x / (|x| + 1e-8))
        std::vector<float> xyz((size_t)L * 3), dir((size_t)L * 3);
        CK(cudaMemcpy(xyz.data(), h->lxyz, xyz.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (int l = 0; l < L; l++) {
            float x = xyz[l * 3], y = xyz[l * 3 + 1], z = xyz[l * 3 + 2];
            float n = sqrtf(x * x + y * y + z * z) + 1e-8f;
            dir[l * 3] = x / n; dir[l * 3 + 1] = y / n; dir[l * 3 + 2] = z / n;
        }
        if (upload_raw(h, &h->ldir, dir.data(), dir.size(), st)) return 1;
    }
    lin_tc_clear(h->lin_tc);         // the packed GEMM images refer to the previous weights
    if (tc_upload(h->tc, w, C, h->err, st)) return 1;
    if (tc2_upload(h->tc2, w, C, h->err, st)) return 1;
    if (h->tc_variant == 8 && tc8_upload(h->tc8, w, C, h->err, st)) return 1;
    // the uploads above are pageable host -> device copies on the legacy stream (they return once staged): make them visible to callers
    // that render on non-blocking streams
    CK(cudaStreamSynchronize(cudaStreamLegacy));
    CK(cudaStreamSynchronize(st));
    h->have_weights = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------- frame
extern "C" int ra_set_frame(ra_handle* h, const ra_frame* f, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->have_weights) { h->err = "ra_set_frame before ra_upload_weights"; return 1; }
    h->frame = *f;
    int N = h->cfg.n_verts;
    LAUNCH(h, k_frame_prep, 1, 1024, 0, st, h->fc, f->R, f->Th, f->pverts, N, f->wbounds, f->poses, f->mat_cond,
           3 * h->cfg.n_bones, h->resd_w0_raw, h->resd_b0_raw, h->resd_w4_raw, h->resd_b4_raw, h->rend_w3_raw, h->rend_b3_raw, h->cell_count, h->cell_h, h->grid2_ratio);
    LAUNCH(h, k_grid_count, (N + 255) / 256, 256, 0, st, h->fc, 0, f->pverts, (const float4*)nullptr, N, h->cell_count, h->vert_cell);
    LAUNCH(h, k_grid_scan, 1, 1024, 0, st, h->fc, 0, h->cell_count, h->sv.cell_start, h->cell_fill);
    LAUNCH(h, k_grid_order, (N + 255) / 256, 256, 0, st, N, h->vert_cell, h->sv.cell_start, h->cell_fill, h->vert_order);
    LAUNCH(h, k_grid_fill, (N + 127) / 128, 128, 0, st, h->fc, f->pverts, f->pnorm, f->tverts, f->weights, f->A, f->big_A, N,
           h->cfg.n_bones, h->vert_cell, h->sv.cell_start, h->vert_order, h->sv);
    // per-cell neighbourhood lists (3x3x3 block + radius-2 / radius-3 shells) for the near phase of the 3-NN search
    LAUNCH(h, k_nb_count, h->sms * 2, 256, 0, st, h->fc, h->sv.cell_start, h->nb_cnt, h->sv.nb_mask);
    LAUNCH(h, k_nb_scan, RA_NB_LEVELS, 1024, 0, st, h->fc, h->nb_cnt, h->sv);
    LAUNCH(h, k_nb_fill, h->sms * 4, 256, 0, st, h->fc, h->sv.cell_start, (const float4*)h->sv.pos, h->sv);
    // coarse second level over the cell-sorted vertices
    LAUNCH(h, k_grid_count, (N + 255) / 256, 256, 0, st, h->fc, 1, f->pverts, (const float4*)h->sv.pos, N, h->cell_count, h->vert_cell);
    LAUNCH(h, k_grid_scan, 1, 1024, 0, st, h->fc, 1, h->cell_count, h->sv.cell_start2, h->cell_fill);
    LAUNCH(h, k_grid_order, (N + 255) / 256, 256, 0, st, N, h->vert_cell, h->sv.cell_start2, h->cell_fill, h->vert_order);
    LAUNCH(h, k_grid_fill2, (N + 255) / 256, 256, 0, st, (const float4*)h->sv.pos, N, h->vert_cell, h->sv.cell_start2, h->vert_order, h->sv.pos2);
    LAUNCH(h, k_grid_occ, 1, 1024, 0, st, h->fc, h->sv.cell_start2, (const float4*)h->sv.pos2, h->sv);
    if (h->cfg.precision == RA_PRECISION_TC) {
        if (h->tc_variant == 2 || h->tc_variant >= 6) tc2_set_frame(h->tc2, h->fc, st, h->launches);
        if (h->tc_variant == 8) tc8_set_frame(h->tc8, h->fc, st, h->launches);
        else tc_set_frame(h->tc, h->fc, st, h->launches);
    }
    CK(cudaGetLastError());
    h->have_frame = true; h->have_render = false;
    return 0;
}

// ---------------------------------------------------------------------------------------------- fp32 MLP passes
// forward of both MLPs on rows [0, rows_cap) of a chunk whose points are `bpts` (chunk-local pointer).
// full == false: distance only (sdf written to net_out).
static void mlp_forward_fp32(ra_handle* h, cudaStream_t st, const float* bpts, const int* count, int row0, int rows, bool full,
                             float* net_out) {
    int g = grid_for(h, rows);
    const float* b0 = &h->fc->resd_b0[0];
    const float* b4 = &h->fc->resd_b4[0];
    LAUNCH(h, k_encode, grid_for(h, (long long)rows * 64), 256, 0, st, bpts, 10, h->Xr0, 64, 64, h->Xr4, 320, 256, 64, count, row0, rows);
    gemm<EPI_RELU>(h, st, h->Xr0, 64, h->resd[0].w, 64, b0, h->ra_[0], 256, nullptr, 0, count, row0, rows, 256, 64);
    gemm<EPI_RELU>(h, st, h->ra_[0], 256, h->resd[1].w, 256, h->resd[1].b, h->ra_[1], 256, nullptr, 0, count, row0, rows, 256, 256);
    gemm<EPI_RELU>(h, st, h->ra_[1], 256, h->resd[2].w, 256, h->resd[2].b, h->ra_[2], 256, nullptr, 0, count, row0, rows, 256, 256);
    gemm<EPI_RELU>(h, st, h->ra_[2], 256, h->resd[3].w, 256, h->resd[3].b, h->Xr4, 320, nullptr, 0, count, row0, rows, 256, 256);
    gemm<EPI_RELU>(h, st, h->Xr4, 320, h->resd[4].w, 320, b4, h->ra_[4], 256, nullptr, 0, count, row0, rows, 256, 320);
    for (int l = 5; l < 8; l++)
        gemm<EPI_RELU>(h, st, h->ra_[l - 1], 256, h->resd[l].w, 256, h->resd[l].b, h->ra_[l], 256, nullptr, 0, count, row0, rows, 256, 256);
    LAUNCH(h, k_skinny, grid_for(h, (long long)rows * 32), 256, 0, st, h->ra_[7], 256, h->resd[8].w, 256, h->resd[8].b, h->z8, 4, count, row0, rows, 3, 256);
    LAUNCH(h, k_resd_finish, g, 256, 0, st, h->z8, 4, bpts, h->cfg.resd_limit, h->resd_o, h->cpts_o, count, row0, rows);
    LAUNCH(h, k_encode, grid_for(h, (long long)rows * 64), 256, 0, st, h->cpts_o, 8, h->Xs0, 64, 64, h->Xs4, 256, 205, 51, count, row0, rows);
    gemm<EPI_SOFTPLUS>(h, st, h->Xs0, 64, h->sdf[0].w, 64, h->sdf[0].b, h->sb_[0], 256, nullptr, 0, count, row0, rows, 256, 64);
    gemm<EPI_SOFTPLUS>(h, st, h->sb_[0], 256, h->sdf[1].w, 256, h->sdf[1].b, h->sb_[1], 256, nullptr, 0, count, row0, rows, 256, 256);
    gemm<EPI_SOFTPLUS>(h, st, h->sb_[1], 256, h->sdf[2].w, 256, h->sdf[2].b, h->sb_[2], 256, nullptr, 0, count, row0, rows, 256, 256);
    gemm<EPI_SOFTPLUS>(h, st, h->sb_[2], 256, h->sdf[3].w, 256, h->sdf[3].b, h->Xs4, 256, nullptr, 0, count, row0, rows, 205, 256);
    gemm<EPI_SOFTPLUS>(h, st, h->Xs4, 256, h->sdf[4].w, 256, h->sdf[4].b, h->sb_[4], 256, nullptr, 0, count, row0, rows, 256, 256);
    for (int l = 5; l < 8; l++)
        gemm<EPI_SOFTPLUS>(h, st, h->sb_[l - 1], 256, h->sdf[l].w, 256, h->sdf[l].b, h->sb_[l], 256, nullptr, 0, count, row0, rows, 256, 256);
    if (full)
    {   // 257 outputs = 256 feature columns (one tensor-core tile) + the sdf column (rows were permuted to [feat(256), sdf])
        gemm<EPI_NONE>(h, st, h->sb_[7], 256, h->sdf[8].w, 256, h->sdf[8].b, h->out257, 264, nullptr, 0, count, row0, rows, 256, 256);
        LAUNCH(h, k_skinny, grid_for(h, (long long)rows * 32), 256, 0, st, h->sb_[7], 256, h->sdf[8].w + 256 * 256, 256, h->sdf[8].b + 256,
               h->out257 + 256, 264, count, row0, rows, 1, 256);
    }
    else
        LAUNCH(h, k_skinny, grid_for(h, (long long)rows * 32), 256, 0, st, h->sb_[7], 256, h->sdf[8].w + 256 * 256, 256, h->sdf[8].b + 256,
               net_out, 1, count, row0, rows, 1, 256);
}

// analytic d sdf / d bpts through both MLPs (what autograd does at base_network.py:463-468)
static void mlp_backward_fp32(ra_handle* h, cudaStream_t st, const float* bpts, const int* count, int row0, int rows) {
    int g = grid_for(h, rows);
    int ge = grid_for(h, (long long)rows * 256);
    // ---- SDF net
    LAUNCH(h, k_outer_small<EPI_MUL_DSOFTPLUS>, ge, 256, 0, st, (const float*)nullptr, 0, h->sdf[8].w + 256 * 256, 256, 1, 256,
           h->sb_[7], 256, h->GA, 256, count, row0, rows);
    float *cur = h->GA, *nxt = h->GB;
    for (int l = 7; l >= 5; l--) {   // dZ_{l-1} = (dZ_l W_l) * dsp(a_{l-1})
        gemm<EPI_MUL_DSOFTPLUS>(h, st, cur, 256, h->sdf[l].wt, h->sdf[l].Kt, nullptr, nxt, 256, h->sb_[l - 1], 256, count, row0, rows, 256, 256);
        std::swap(cur, nxt);
    }
    // layer 4 input = [a3 (205) | PE8 (51)]
    gemm<EPI_MUL_DSOFTPLUS>(h, st, cur, 256, h->sdf[4].wt, h->sdf[4].Kt, nullptr, nxt, 256, h->Xs4, 256, count, row0, rows, 205, 256);
    gemm<EPI_NONE>(h, st, cur, 256, h->sdf[4].wt + (size_t)205 * h->sdf[4].Kt, h->sdf[4].Kt, nullptr, h->dpes, 64, nullptr, 0, count, row0, rows, 51, 256);
    std::swap(cur, nxt);
    // dZ_2 = (dZ_3 W_3) * dsp(a_2): K = 205 padded to 208 (zero weight columns)
    gemm<EPI_MUL_DSOFTPLUS>(h, st, cur, 256, h->sdf[3].wt, h->sdf[3].Kt, nullptr, nxt, 256, h->sb_[2], 256, count, row0, rows, 256, 208);
    std::swap(cur, nxt);
    for (int l = 2; l >= 1; l--) {
        gemm<EPI_MUL_DSOFTPLUS>(h, st, cur, 256, h->sdf[l].wt, h->sdf[l].Kt, nullptr, nxt, 256, h->sb_[l - 1], 256, count, row0, rows, 256, 256);
        std::swap(cur, nxt);
    }
    gemm<EPI_NONE>(h, st, cur, 256, h->sdf[0].wt, h->sdf[0].Kt, nullptr, h->dpe0, 64, nullptr, 0, count, row0, rows, 51, 256);
    LAUNCH(h, k_sdf_grad_to_cp, g, 256, 0, st, h->cpts_o, h->dpe0, 64, h->dpes, 64, 0, h->z8, 4, h->cfg.resd_limit, h->gcp, h->u4, 4, count, row0, rows);
    // ---- residual net
    LAUNCH(h, k_outer_small<EPI_MUL_DRELU>, ge, 256, 0, st, h->u4, 4, h->resd[8].w, 256, 3, 256, h->ra_[7], 256, h->GA, 256, count, row0, rows);
    cur = h->GA; nxt = h->GB;
    for (int l = 7; l >= 5; l--) {
        gemm<EPI_MUL_DRELU>(h, st, cur, 256, h->resd[l].wt, h->resd[l].Kt, nullptr, nxt, 256, h->ra_[l - 1], 256, count, row0, rows, 256, 256);
        std::swap(cur, nxt);
    }
    gemm<EPI_MUL_DRELU>(h, st, cur, 256, h->resd[4].wt, h->resd[4].Kt, nullptr, nxt, 256, h->Xr4, 320, count, row0, rows, 256, 256);
    gemm<EPI_NONE>(h, st, cur, 256, h->resd[4].wt + (size_t)256 * h->resd[4].Kt, h->resd[4].Kt, nullptr, h->dpes, 64, nullptr, 0, count, row0, rows, 63, 256);
    std::swap(cur, nxt);
    for (int l = 3; l >= 1; l--) {
        gemm<EPI_MUL_DRELU>(h, st, cur, 256, h->resd[l].wt, h->resd[l].Kt, nullptr, nxt, 256, h->ra_[l - 1], 256, count, row0, rows, 256, 256);
        std::swap(cur, nxt);
    }
    gemm<EPI_NONE>(h, st, cur, 256, h->resd[0].wt, h->resd[0].Kt, nullptr, h->dpe0, 64, nullptr, 0, count, row0, rows, 63, 256);
    LAUNCH(h, k_resd_grad_to_bp, g, 256, 0, st, bpts, h->dpe0, 64, h->dpes, 64, 0, h->gcp, h->gbp, count, row0, rows);
}

// a3 of the residual net lives in Xr4[:, :256] (ld 320); a3 of the SDF net in Xs4[:, :205] (ld 256).
// ra_[3] / sb_[3] are unused placeholders.

// forward + gradient + heads + raw assembly for the attribute list, chunk by chunk
static int attr_pass(ra_handle* h, cudaStream_t st, int64_t max_rows, bool sync_count) {
    int64_t rows_total = max_rows;
    if (sync_count) {
        int c = 0;
        CK(cudaMemcpyAsync(&c, h->al.count, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        rows_total = c;
    }
    const int* count = h->al.count;
    for (int64_t row0 = 0; row0 < rows_total; row0 += ATTR_CH) {
        int rows = (int)std::min<int64_t>(ATTR_CH, rows_total - row0);
        const float* bp = h->al.bpts + row0 * 3;
        int g = grid_for(h, rows);
        mlp_forward_fp32(h, st, bp, count, (int)row0, rows, true, nullptr);
        mlp_backward_fp32(h, st, bp, count, (int)row0, rows);
        LAUNCH(h, k_attr_normals, g, 256, 0, st, h->fc, h->gbp, h->al.mats + row0 * 18, h->nrm_o, count, (int)row0, rows);
        if (h->cfg.relight) {
            gemm<EPI_SOFTPLUS>(h, st, h->out257, 264, h->alb[0].w, 256, h->alb[0].b, h->hd1, 128, nullptr, 0, count, (int)row0, rows, 128, 256);
            gemm<EPI_SOFTPLUS>(h, st, h->hd1, 128, h->alb[1].w, 128, h->alb[1].b, h->hd2, 128, nullptr, 0, count, (int)row0, rows, 128, 128);
            LAUNCH(h, k_skinny, grid_for(h, (long long)rows * 32), 256, 0, st, h->hd2, 128, h->alb[2].w, 128, h->alb[2].b, h->head_a, 4, count, (int)row0, rows, 3, 128);
            gemm<EPI_SOFTPLUS>(h, st, h->out257, 264, h->rgh[0].w, 256, h->rgh[0].b, h->hd1, 128, nullptr, 0, count, (int)row0, rows, 128, 256);
            gemm<EPI_SOFTPLUS>(h, st, h->hd1, 128, h->rgh[1].w, 128, h->rgh[1].b, h->hd2, 128, nullptr, 0, count, (int)row0, rows, 128, 128);
            LAUNCH(h, k_skinny, grid_for(h, (long long)rows * 32), 256, 0, st, h->hd2, 128, h->rgh[2].w, 128, h->rgh[2].b, h->head_r, 4, count, (int)row0, rows, 1, 128);
        } else {
            if (!h->rend[0].w) { h->err = "render_network weights missing for AniSDF"; return 1; }
            LAUNCH(h, k_render_input, grid_for(h, (long long)rows * 288), 256, 0, st, h->al.bvds + row0 * 3, h->nrm_o, h->out257, 264, h->Xrn, 288, count, (int)row0, rows);
            gemm<EPI_RELU>(h, st, h->Xrn, 288, h->rend[0].w, 288, h->rend[0].b, h->rn1, 256, nullptr, 0, count, (int)row0, rows, 256, 288);
            gemm<EPI_RELU>(h, st, h->rn1, 256, h->rend[1].w, 256, h->rend[1].b, h->rn2, 256, nullptr, 0, count, (int)row0, rows, 256, 256);
            gemm<EPI_RELU>(h, st, h->rn2, 256, h->rend[2].w, 256, h->rend[2].b, h->rn1, 256, nullptr, 0, count, (int)row0, rows, 256, 256);
            gemm<EPI_RELU>(h, st, h->rn1, 256, h->rend[3].w, 256, &h->fc->rend_b3[0], h->rn2, 256, nullptr, 0, count, (int)row0, rows, 256, 256);
            LAUNCH(h, k_skinny, grid_for(h, (long long)rows * 32), 256, 0, st, h->rn2, 256, h->rend[4].w, 256, h->rend[4].b, h->head_a, 4, count, (int)row0, rows, 3, 256);
        }
        LAUNCH(h, k_attr_finish, g, 256, 0, st, h->cfg.relight, bp, h->cpts_o, h->resd_o, h->out257 + 256, 264, h->nrm_o, h->head_a, h->head_r,
               h->beta, h->cfg.albedo_slope, h->cfg.albedo_bias, h->cfg.rough_slope, h->cfg.rough_bias, h->al.src + row0, h->raw, count, (int)row0, rows);
    }
    CK(cudaGetLastError());
    return 0;
}

// distance MLPs over the query list (net sdf into q.net)
static cudaEvent_t prof_event(std::vector<cudaEvent_t>& pool, size_t& used) {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
}
static void prof_stage(ra_handle* h, cudaStream_t st) {
    if (h->prof) cudaEventRecord(prof_event(h->ev_stage, h->ev_stage_used), st);
}

// `rows_bound` (fp32 mode only): a host-side upper bound of the work list's length.  The GEMM chain runs chunk by chunk over
// [0, rows_bound); every kernel reads the true count on the device and returns early past it, so no counter is read back per
// tracing iteration (the surface stage is bounded by P, the shadow stages by the shadow-ray count read ONCE per stage).
static int distance_pass(ra_handle* h, cudaStream_t st, int64_t rows_bound, const QueryList* ql = nullptr) {
    const QueryList& q = ql ? *ql : h->q;
    if (h->cfg.precision == RA_PRECISION_TC) {
        if (h->prof) cudaEventRecord(prof_event(h->ev_mlp, h->ev_mlp_used), st);
        if (h->tc_variant == 8) tc8_distance(h->tc8, q.bpts, q.net, q.count, h->cfg.resd_limit, h->sms, st, h->launches);
        else if (h->tc_variant == 7) tc7_distance(h->tc2, q.bpts, q.net, q.count, h->cfg.resd_limit, h->sms, st, h->launches);
        else if (h->tc_variant == 6) tc6_distance(h->tc2, q.bpts, q.net, q.count, h->cfg.resd_limit, h->sms, st, h->launches);
        else if (h->tc_variant == 2) tc2_distance(h->tc2, q.bpts, q.net, q.count, h->cfg.resd_limit, h->sms, st, h->launches);
        else tc_distance(h->tc, q.bpts, q.net, q.count, h->cfg.resd_limit, h->sms, st, h->launches);
        if (h->prof) cudaEventRecord(prof_event(h->ev_mlp, h->ev_mlp_used), st);
        return 0;
    }
    const int64_t c = std::min<int64_t>(rows_bound, h->q_cap);
    for (int64_t row0 = 0; row0 < c; row0 += ATTR_CH) {
        int rows = (int)std::min<int64_t>(ATTR_CH, c - row0);
        mlp_forward_fp32(h, st, q.bpts + row0 * 3, q.count, (int)row0, rows, false, q.net + row0);
    }
    return 0;
}

// one device counter -> host (fp32 mode: once per shadow stage; never on the tensor-core product path)
static int read_counter(ra_handle* h, const int* dev, cudaStream_t st, int64_t* out) {
    int c = 0;
    CK(cudaMemcpyAsync(&c, dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *out = c;
    return 0;
}

static Counters no_count(ra_handle* h) { return h->cnt; }

// ---------------------------------------------------------------------------------------------- render: sphere tracing
static int zero_outputs(ra_handle* h, const ra_outputs* o, int64_t P, cudaStream_t st) {
    int L = h->cfg.env_h * h->cfg.env_w;
    struct { float* p; size_t n; } m[] = {{o->rgb_map, 3}, {o->acc_map, 1}, {o->depth_map, 1}, {o->surf_map, 3}, {o->norm_map, 3},
                                          {o->cpts_map, 3}, {o->bpts_map, 3}, {o->resd_map, 3}, {o->albedo_map, 3}, {o->roughness_map, 1},
                                          {o->shade_map, 3}, {o->lvis_map, (size_t)L}, {o->ldot_map, (size_t)L}, {o->spec_map, 3}};
    for (auto& e : m)
        if (e.p) CK(cudaMemsetAsync(e.p, 0, e.n * P * sizeof(float), st));
    return 0;
}

static int render_trace(ra_handle* h, const float* ray_o, const float* ray_d, const float* near_, const float* far_, int64_t P,
                        const ra_outputs* out, cudaStream_t st) {
    if (!h->have_frame) { h->err = "render before ra_set_frame"; return 1; }
    if (P > h->P_cap) { h->err = "P exceeds ra_config.max_rays"; return 1; }
    const ra_config& c = h->cfg;
    if (zero_outputs(h, out, P, st)) return 1;
    CK(cudaMemsetAsync(h->counters_blk, 0, 16 * sizeof(int), st));
    h->last_P = P; h->last_ray_o = ray_o; h->have_render = true;      // P == 0 is a valid (empty) render: ray_o may be null
    const int64_t Pg = (h->lay_world > 1 && h->lay_global_P > 0) ? h->lay_global_P : P;     // the reference chunks the WHOLE frame's rays
    int n_chunks = std::max<int64_t>((Pg + c.render_chunk - 1) / c.render_chunk, 1);
    h->chunk_actual = Pg ? (int)((Pg + n_chunks - 1) / n_chunks) : 1;        // chunkify's equalised size, net_utils.py:323
    if (P == 0) return 0;
    TraceCfg tc{c.st_iter, c.st_tan_i, c.st_relax, c.st_offset, c.st_eps, c.st_skip, c.dist_th, c.blend_radius};
    int N = c.n_verts;
    int g = grid_for(h, P, h->tb_surf, 8 * 256 / h->tb_surf);
    prof_stage(h, st);
    for (int it = 0; it <= c.st_iter; it++) {
        CK(cudaMemsetAsync(h->q.count, 0, sizeof(int), st));
        LAUNCH(h, k_trace_surface, g, h->tb_surf, 0, st, it, tc, h->fc, h->sv, N, ray_o, ray_d, near_, far_, (int)P, h->ss, h->q, h->cnt,
               h->surf, h->acc, h->depth, h->fg_ray, (h->pkt_search >> 2) & 1);
        if (it < c.st_iter && distance_pass(h, st, P)) return 1;
    }
    prof_stage(h, st);
    // surface samples -> attributes
    int C = c.relight ? 17 : 16;
    CK(cudaMemsetAsync(h->raw, 0, (size_t)P * c.n_samples * C * sizeof(float), st));
    LAUNCH(h, k_attr_front, grid_for(h, P * c.n_samples, 256, 8), 256, 0, st, 1, h->fc, h->sv, N, c.dist_th, c.blend_radius,
           (const float*)nullptr, (const float*)nullptr, 0LL, h->cnt.n_fg, h->fg_ray, h->surf, ray_o, ray_d, near_, far_, c.n_samples,
           c.surf_sample_range, c.clip_near, c.clip_far, 0LL, 0LL, h->al, h->cnt, 0);
    if (attr_pass(h, st, P * c.n_samples, h->cfg.precision == RA_PRECISION_FP32)) return 1;
    OutMaps om{out->rgb_map, out->acc_map, out->depth_map, out->surf_map, out->norm_map, out->cpts_map, out->bpts_map, out->resd_map,
               out->albedo_map, out->roughness_map, out->shade_map};
    LAUNCH(h, k_surface_blend, grid_for(h, P), 256, 0, st, c.relight, h->cnt.n_fg, h->fg_ray, h->raw, c.n_samples, h->acc, h->surf, h->depth,
           c.albedo_slope, c.albedo_bias, c.rough_slope, c.rough_bias, c.albedo_multiplier > 0.f ? c.albedo_multiplier : 1.f /* <= 0 means 'off' (sphere_tracing_renderer.py:653) */, h->fm, om);
    if (!c.relight) { CK(cudaGetLastError()); return 0; }
    prof_stage(h, st);
    // light visibility (DFSS)
    int L = c.env_h * c.env_w;
    LAUNCH(h, k_shadow_gen, grid_for(h, P * L / 4, 256, 16), 256, 0, st, h->fc, h->cnt.n_fg, h->fg_ray, h->surf, h->fm.norm, h->ldir, L,
           c.lv_near, c.bbox_margin, h->chunk_actual, h->lay_block, h->lay_world, h->lay_rank, c.visibility_mode, h->lvis, h->ldot, h->sr, h->cnt.n_shadow, (h->pkt_order >> 1) & 1);
    TraceCfg sc{c.lv_iter, 1.f, c.lv_relax, c.lv_offset, c.st_eps, c.st_skip, c.lv_dist_th, c.blend_radius};
    const bool split = h->overlap && h->cfg.precision == RA_PRECISION_TC && (h->tc_variant == 1 || h->tc_variant >= 6);
    int64_t n_sh = h->sr.cap;
    if (c.visibility_mode) n_sh = 0;          // cfg.local_visibility / cfg.no_visibility: k_shadow_gen already wrote the final lvis, nothing to trace
    else if (h->cfg.precision == RA_PRECISION_FP32 && read_counter(h, h->cnt.n_shadow, st, &n_sh)) return 1;
    if (c.visibility_mode) {
    } else if (!split) {
        int gs = grid_for(h, P * 64, h->tb_shadow, 8 * 256 / h->tb_shadow);
        for (int it = 0; it <= c.lv_iter; it++) {
            CK(cudaMemsetAsync(h->q.count, 0, sizeof(int), st));
            LAUNCH(h, k_trace_shadow, gs, h->tb_shadow, 0, st, it, sc, h->fc, h->sv, N, h->cnt.n_shadow, h->fg_ray, h->surf, h->ldir, h->lsharp, L, h->sr,
                   h->q, h->cnt, h->lvis, 0, 1, (h->pkt_search >> 1) & 1, h->final_skip);
            if (it < c.lv_iter && distance_pass(h, st, n_sh)) return 1;
        }
    } else {
        // Two halves of the shadow rays on two streams: while the fused MLP kernel (tensor pipe, 2 CTAs x 80 regs per SM)
        // works on one half, the CUDA-core tracing kernel of the other half runs in the register space it leaves free.
        int gs = grid_for(h, P * 32, 128, 8);
        CK(cudaEventRecord(h->ev_fork, st));
        CK(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
        for (int it = 0; it <= c.lv_iter; it++) {
            for (int part = 0; part < 2; part++) {
                cudaStream_t ps = part ? h->aux : st;
                const QueryList& ql = part ? h->q2 : h->q;
                CK(cudaMemsetAsync(ql.count, 0, sizeof(int), ps));
                LAUNCH(h, k_trace_shadow, gs, 128, 0, ps, it, sc, h->fc, h->sv, N, h->cnt.n_shadow, h->fg_ray, h->surf, h->ldir, h->lsharp, L,
                       h->sr, ql, h->cnt, h->lvis, part, 2, (h->pkt_search >> 1) & 1, h->final_skip);
                if (it < c.lv_iter && distance_pass(h, ps, n_sh, &ql)) return 1;
            }
        }
        CK(cudaEventRecord(h->ev_join, h->aux));
        CK(cudaStreamWaitEvent(st, h->ev_join, 0));
    }
    prof_stage(h, st);
    LAUNCH(h, k_shade, grid_for(h, P * 32, 256, 8), 256, 0, st, h->cnt.n_fg, h->fg_ray, ray_o, h->surf, h->acc, h->fm, h->lvis, h->ldot,
           h->lxyz, h->larea, L, h->main_light ? h->main_light : h->env_main, h->main_light ? h->mlh : h->emh, h->main_light ? h->mlw : h->emw,
           c.fresnel_f0, c.shading_albedo, 0, 1, c.tonemapping, c.brdf_mode, out->rgb_map, out->shade_map, out->spec_map);
    if (out->lvis_map || out->ldot_map)
        LAUNCH(h, k_scatter_lmaps, grid_for(h, P * L / 4), 256, 0, st, h->cnt.n_fg, h->fg_ray, h->acc, h->lvis, h->ldot, L, out->lvis_map, out->ldot_map);
    prof_stage(h, st);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_render_relight(ra_handle* h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                                 int64_t P, const ra_outputs* out, void* stream) {
    if (!h->cfg.relight) { h->err = "handle was created with relight=0"; return 1; }
    return render_trace(h, ray_o, ray_d, near_, far_, P, out, (cudaStream_t)stream);
}

extern "C" int ra_render_anisdf_trace(ra_handle* h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                                      int64_t P, const ra_outputs* out, void* stream) {
    if (h->cfg.relight) { h->err = "handle was created with relight=1"; return 1; }
    return render_trace(h, ray_o, ray_d, near_, far_, P, out, (cudaStream_t)stream);
}

static int relight_envmaps_impl(ra_handle* h, const float* probes, int32_t n_env, float* rgb, float* shade, float* spec, void* stream, int raw) {
    cudaStream_t st = (cudaStream_t)stream;
    const ra_config& c = h->cfg;
    if (!c.relight || !h->have_render) { h->err = "ra_relight_envmaps needs a preceding ra_render_relight"; return 1; }
    int64_t P = h->last_P;
    int L = c.env_h * c.env_w;
    for (int e = 0; e < n_env; e++) {      // background pixels: rgb = shade = 0, spec = probe-dependent constant (all-zero inputs)
        const float* probe = probes + (size_t)e * L * 3;
        float* r = rgb ? rgb + (size_t)e * P * 3 : nullptr;
        float* s = shade ? shade + (size_t)e * P * 3 : nullptr;
        float* p = spec ? spec + (size_t)e * P * 3 : nullptr;
        if (r) CK(cudaMemsetAsync(r, 0, (size_t)P * 3 * sizeof(float), st));
        if (s) CK(cudaMemsetAsync(s, 0, (size_t)P * 3 * sizeof(float), st));
        if (p && P && raw) CK(cudaMemsetAsync(p, 0, (size_t)P * 3 * sizeof(float), st));      // times acc = 0
        if (p && P && !raw) {
            LAUNCH(h, k_bg_spec, 1, 32, 0, st, h->lxyz, h->larea, L, probe, c.env_h, c.env_w, c.fresnel_f0, c.brdf_mode, h->bg_spec);
            LAUNCH(h, k_fill3, grid_for(h, P * 3), 256, 0, st, p, h->bg_spec, (long long)P);
        }
    }
    for (int e0 = 0; e0 < n_env && P; e0 += RA_SHADE_MULTI) {      // foreground: up to 4 probes per pass over the visibility maps
        int ne = std::min(RA_SHADE_MULTI, n_env - e0);
        LAUNCH(h, k_shade_multi, grid_for(h, P * 32, 256, 8), 256, 0, st, h->cnt.n_fg, h->fg_ray, h->last_ray_o, h->surf, h->acc, h->fm, h->lvis,
               h->ldot, h->lxyz, h->larea, L, probes + (size_t)e0 * L * 3, ne, c.env_h, c.env_w, c.fresnel_f0, c.shading_albedo,
               rgb ? rgb + (size_t)e0 * P * 3 : nullptr, shade ? shade + (size_t)e0 * P * 3 : nullptr,
               spec ? spec + (size_t)e0 * P * 3 : nullptr, (long long)P, raw ? 0 : 1, raw ? 1 : 0, 1, c.brdf_mode);      // the novel-light re-shade maps unconditionally (novel_light_sphere_tracing.py:48)
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_allgather(ra_handle* h, void* comm, const float* send, int64_t n_floats, float* recv, void* stream) {
    // ncclResult_t ncclAllGather(const void* sendbuff, void* recvbuff, size_t sendcount, ncclDataType_t datatype, ncclComm_t comm, cudaStream_t stream)
    typedef int (*allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
    static allgather_fn fn = nullptr;
    if (!fn) {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // the copy already loaded by the host process, if any
        if (!lib) { h->err = std::string("ra_allgather: dlopen(libnccl.so.2): ") + dlerror(); return 1; }
        fn = (allgather_fn)dlsym(lib, "ncclAllGather");
        if (!fn) { h->err = "ra_allgather: ncclAllGather not found"; return 1; }
    }
    const int nccl_float32 = 7;      // ncclFloat32 (nccl.h)
    int rc = fn(send, recv, (size_t)n_floats, nccl_float32, comm, (cudaStream_t)stream);
    if (rc != 0) { h->err = "ra_allgather: ncclAllGather returned " + std::to_string(rc); return 1; }
    return 0;
}

extern "C" int ra_set_ray_layout(ra_handle* h, int64_t global_P, int32_t block, int32_t world, int32_t rank) {
    if (world < 1 || rank < 0 || rank >= world || block < 1 || global_P < 0) { h->err = "ra_set_ray_layout: bad arguments"; return 1; }
    h->lay_global_P = global_P; h->lay_block = block; h->lay_world = world; h->lay_rank = rank;
    return 0;
}

extern "C" int ra_relight_envmaps(ra_handle* h, const float* probes, int32_t n_env, float* rgb, float* shade, float* spec, void* stream) {
    return relight_envmaps_impl(h, probes, n_env, rgb, shade, spec, stream, 0);
}
extern "C" int ra_relight_envmaps_raw(ra_handle* h, const float* probes, int32_t n_env, float* rgb, float* shade, float* spec, void* stream) {
    return relight_envmaps_impl(h, probes, n_env, rgb, shade, spec, stream, 1);
}

// ---------------------------------------------------------------------------------------------- batch preparation (row f1)
template <typename T>
static int upload_any(ra_handle* h, T** dst, const T* src, size_t n, cudaStream_t st) {
    hfree(h, *dst);
    CK(dalloc(h, dst, n));
    CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int ra_upload_body(ra_handle* h, const ra_body* b, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->cfg.n_verts, J = h->cfg.n_bones;
    if (!b->tjoints || !b->parents || !b->rverts || !b->weights || (!b->rnorm && !b->faces)) { h->err = "ra_upload_body: tjoints, parents, rverts, weights and rnorm-or-faces are required"; return 1; }
    if (upload_any(h, &h->body.tjoints, b->tjoints, (size_t)J * 3, st)) return 1;
    if (upload_any(h, &h->body.parents, (const int*)b->parents, (size_t)J, st)) return 1;
    if (upload_any(h, &h->body.rverts, b->rverts, (size_t)N * 3, st)) return 1;
    if (upload_any(h, &h->body.weights, b->weights, (size_t)N * J, st)) return 1;
    if (b->rnorm) { if (upload_any(h, &h->body.rnorm, b->rnorm, (size_t)N * 3, st)) return 1; }
    else hfree(h, h->body.rnorm);
    h->body.n_faces = 0;
    if (b->faces && b->n_faces > 0) {
        if (upload_any(h, &h->body.faces, (const int*)b->faces, (size_t)b->n_faces * 3, st)) return 1;
        h->body.n_faces = b->n_faces;
    }
    if (!h->prep_mm) { CK(dalloc(h, &h->prep_mm, 16)); CK(dalloc(h, &h->prep_nacc, (size_t)N * 3)); }
    h->body.ready = true;
    return 0;
}

extern "C" int ra_prepare_pose(ra_handle* h, const float* poses, const float* Rh, const float* Th, float bounds_pad, const ra_pose_outputs* o, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->body.ready) { h->err = "ra_prepare_pose before ra_upload_body"; return 1; }
    if (!o->A || !o->R || !o->pverts || !o->pnorm) { h->err = "ra_prepare_pose: A, R, pverts, pnorm outputs are required"; return 1; }
    const int N = h->cfg.n_verts, J = h->cfg.n_bones;
    if (J + 1 > 1024) { h->err = "ra_prepare_pose: too many bones"; return 1; }
    int threads = ((J + 1 + 31) / 32) * 32;
    LAUNCH(h, k_prep_pose, 1, threads, (size_t)J * 21 * sizeof(float), st, poses, Rh, h->body.tjoints, h->body.parents, J, o->A, (float*)nullptr, o->R);
    LAUNCH(h, k_prep_bounds_init, 1, 32, 0, st, h->prep_mm);
    const bool mesh_normals = h->body.n_faces > 0;
    LAUNCH(h, k_prep_verts, (N + 127) / 128, 128, 0, st, h->body.rverts, mesh_normals ? (const float*)nullptr : h->body.rnorm, h->body.weights, o->A, N, J,
           o->R, Th, o->pverts, o->pnorm, o->wverts, o->wnorm, h->prep_mm);
    if (mesh_normals) {
        CK(cudaMemsetAsync(h->prep_nacc, 0, (size_t)N * 3 * sizeof(float), st));
        LAUNCH(h, k_prep_face_normals, (h->body.n_faces + 127) / 128, 128, 0, st, o->pverts, h->body.faces, h->body.n_faces, h->prep_nacc);
        LAUNCH(h, k_prep_normals_finish, (N + 127) / 128, 128, 0, st, h->prep_nacc, N, o->R, o->pnorm, o->wnorm);
    }
    LAUNCH(h, k_prep_bounds_finish, 1, 32, 0, st, h->prep_mm, bounds_pad, o->pbounds, o->wbounds);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_prepare_rays(ra_handle* h, const float* K, const float* R, const float* T, int32_t H, int32_t W, const float* wbounds,
                               float* ray_o, float* ray_d, float* near_, float* far_, unsigned char* mask_at_box, int32_t* n_rays, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (H <= 0 || W <= 0) { h->err = "ra_prepare_rays: bad image size"; return 1; }
    CamDev c;
    // inv(K) by the adjugate in double (np.linalg.inv on the host in the reference, data_utils.py:839)
    double k[9]; for (int i = 0; i < 9; i++) k[i] = K[i];
    double det = k[0] * (k[4] * k[8] - k[5] * k[7]) - k[1] * (k[3] * k[8] - k[5] * k[6]) + k[2] * (k[3] * k[7] - k[4] * k[6]);
    if (det == 0.0) { h->err = "ra_prepare_rays: singular K"; return 1; }
    double inv[9] = {(k[4] * k[8] - k[5] * k[7]), -(k[1] * k[8] - k[2] * k[7]), (k[1] * k[5] - k[2] * k[4]),
                     -(k[3] * k[8] - k[5] * k[6]), (k[0] * k[8] - k[2] * k[6]), -(k[0] * k[5] - k[2] * k[3]),
                     (k[3] * k[7] - k[4] * k[6]), -(k[0] * k[7] - k[1] * k[6]), (k[0] * k[4] - k[1] * k[3])};
    for (int i = 0; i < 9; i++) { c.Kinv[i] = (float)(inv[i] / det); c.R[i] = R[i]; }
    for (int a = 0; a < 3; a++) { c.T[a] = T[a]; c.o[a] = -(R[a] * T[0] + R[3 + a] * T[1] + R[6 + a] * T[2]); }       // -R^T T
    int n = H * W, nb = (n + 255) / 256;
    h->g_W = W; h->g_H = H;
    if (nb > h->blk_cap) { hfree(h, h->blk_cnt); CK(dalloc(h, &h->blk_cnt, (size_t)nb)); h->blk_cap = nb; }
    LAUNCH(h, k_prep_rays_count, nb, 256, 0, st, c, wbounds, H, W, mask_at_box, h->blk_cnt);
    LAUNCH(h, k_scan_blocks, 1, 1024, 0, st, h->blk_cnt, nb);
    LAUNCH(h, k_prep_rays_write, nb, 256, 0, st, c, wbounds, H, W, h->blk_cnt, nb, ray_o, ray_d, near_, far_, (int*)n_rays);
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- ground plane (row f2)
static GroundCfg ground_cfg(ra_handle* h, const ra_ground_config* g) {
    GroundCfg c{};
    for (int i = 0; i < 3; i++) { c.normal[i] = g->normal[i]; c.origin[i] = g->origin[i]; c.albedo[i] = g->albedo[i]; }
    c.attach_envmap = g->attach_envmap; c.shading_albedo = h->cfg.shading_albedo; c.multiplier = g->shading_multiplier;
    c.env_r = h->cfg.env_r; c.near_offset = g->near_offset; c.bbox_margin = h->cfg.bbox_margin; c.tonemap = h->cfg.tonemapping;
    return c;
}

extern "C" int ra_ground_begin(ra_handle* h, const unsigned char* mask_at_box, int32_t H, int32_t W, const float* acc_map, float* acc_g, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int n = H * W, nb = (n + 255) / 256;
    h->g_W = W; h->g_H = H;
    if (nb > h->blk_cap) { hfree(h, h->blk_cnt); CK(dalloc(h, &h->blk_cnt, (size_t)nb)); h->blk_cap = nb; }
    if (n > h->pix_cap) {
        hfree(h, h->pix2ray); hfree(h, h->g_weight);
        CK(dalloc(h, &h->pix2ray, (size_t)n)); CK(dalloc(h, &h->g_weight, (size_t)n)); h->pix_cap = n;
    }
    LAUNCH(h, k_mask_count, nb, 256, 0, st, mask_at_box, n, h->blk_cnt);
    LAUNCH(h, k_scan_blocks, 1, 1024, 0, st, h->blk_cnt, nb);
    LAUNCH(h, k_ground_pix2ray, nb, 256, 0, st, mask_at_box, n, h->blk_cnt, acc_map, h->pix2ray, acc_g);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_render_ground(ra_handle* h, const ra_ground_config* g, const float* ray_o, const float* ray_d, const float* acc_g, int64_t F,
                                const float* probe, int32_t ph, int32_t pw, const float* albedo_image, int32_t ih, int32_t iw,
                                const ra_ground_outputs* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const ra_config& c = h->cfg;
    if (!c.relight || !h->have_frame) { h->err = "ra_render_ground needs a relight handle with a frame set"; return 1; }
    if (F > h->pix_cap) { h->err = "ra_render_ground: call ra_ground_begin for this image size first"; return 1; }
    if (!out->rgb_map || !out->surf_map || !out->albedo_map || !out->lvis_map || !out->ldot_map) { h->err = "ra_render_ground: rgb/surf/albedo/lvis/ldot outputs are required"; return 1; }
    if (F == 0) return 0;
    const int L = c.env_h * c.env_w, N = c.n_verts;
    GroundCfg gc = ground_cfg(h, g);
    if (!h->g_light) CK(dalloc(h, &h->g_light, (size_t)RA_NLIGHT_MAX * 3));
    LAUNCH(h, k_ground_light_table, 2, 256, 0, st, h->lxyz, L, probe, ph, pw, h->g_light);
    const float* img = albedo_image ? albedo_image : probe;
    if (!albedo_image) { ih = ph; iw = pw; }
    // reference chunking (only the bbox growth depends on it): human pass chunks so far, equalised floor chunk size
    const int human_chunks = (int)std::max<int64_t>((h->last_P + c.render_chunk - 1) / c.render_chunk, 1);
    const int64_t n_chunks = std::max<int64_t>((F + c.render_chunk - 1) / c.render_chunk, 1);
    const int chunk_actual = (int)((F + n_chunks - 1) / n_chunks);
    LAUNCH(h, k_ground_setup, grid_for(h, F), 256, 0, st, gc, ray_o, ray_d, 0LL, (long long)F, img, ih, iw, out->surf_map, out->depth_map,
           out->norm_map, out->albedo_map, out->roughness_map, h->g_weight);
    TraceCfg sc{g->iter, 1.f, g->relax, g->offset, c.st_eps, c.st_skip, g->dist_th, c.blend_radius};
    int* n_gshadow = h->counters_blk + 5;
    // The floor's rays get a list of their own, sized for the whole image (256 entries of 34 B + a 16 B query slot per pixel: 3.4 GB at
    // 512^2) up to RA_GROUND_SLOTS_MAX: one batch of 17 tracer + 16 MLP launches instead of one batch per max_rays pixels (at 512^2: four
    // batches, two of them over the nearly ray-less upper half of the image -- 34 launches of ~0.15 ms and 48 tiny MLP launches for nothing).
    {
        const int64_t want = std::max<int64_t>(std::min<int64_t>(256 * F, RA_GROUND_SLOTS_MAX), 256 * 32);
        if (h->g_slots < want) {
            ShadowRays& g2 = h->sr_g;
            hfree(h, g2.fg); hfree(h, g2.light); hfree(h, g2.near_); hfree(h, g2.far_); hfree(h, g2.t); hfree(h, g2.occ); hfree(h, g2.d0); hfree(h, g2.q_smpl);
            hfree(h, g2.q_slot); hfree(h, h->q_g.bpts); hfree(h, h->q_g.net);
            h->g_slots = 0;
            CK(dalloc(h, &g2.fg, want)); CK(dalloc(h, &g2.light, want)); CK(dalloc(h, &g2.near_, want)); CK(dalloc(h, &g2.far_, want));
            CK(dalloc(h, &g2.t, want)); CK(dalloc(h, &g2.occ, want)); CK(dalloc(h, &g2.d0, want)); CK(dalloc(h, &g2.q_smpl, want)); CK(dalloc(h, &g2.q_slot, want));
            CK(dalloc(h, &h->q_g.bpts, (size_t)want * 3)); CK(dalloc(h, &h->q_g.net, (size_t)want));
            g2.cap = (int)want; g2.dropped = h->sr.dropped; g2.n_rays = h->sr.n_rays;
            h->q_g.count = h->q.count;
            h->g_slots = want;
        }
    }
    // processing granularity: as many pixels as the floor's ray list holds (256 rays per pixel)
    int64_t step = std::max<int64_t>(h->g_slots / 256, 1);
    if (step >= 32) step &= ~(int64_t)31;          // whole 32-pixel packets per batch (k_ground_rays pads the last one)
    // packets of 8 x 4 pixels (about half the radius of 32 x 1 on the floor) when the batches can be whole groups of 4 image rows
    int tile_w = 0;
    if ((int64_t)h->g_W * h->g_H == F && h->g_W % 8 == 0 && h->g_H % 4 == 0 && step >= 4 * h->g_W) { tile_w = h->g_W; step = step / (4 * tile_w) * (4 * tile_w); }
    for (int64_t p0 = 0; p0 < F; p0 += step) {
        const int64_t n = std::min<int64_t>(step, F - p0);
        CK(cudaMemsetAsync(n_gshadow, 0, sizeof(int), st));
        LAUNCH(h, k_ground_vis_init, grid_for(h, n * L / 4, 256, 16), 256, 0, st, gc, acc_g, (long long)p0, (long long)n, h->ldir, L, out->lvis_map);
        LAUNCH(h, k_ground_rays, (h->pkt_order & 1) ? grid_for(h, n, 128, 16) : grid_for(h, n * L / 4, 256, 16), (h->pkt_order & 1) ? 128 : 256, 0, st, h->fc, gc,
               out->surf_map, acc_g, (long long)p0, (long long)n, h->ldir, L, human_chunks, chunk_actual, h->sr_g, n_gshadow, h->pkt_order & 1, tile_w);
        const int gs = grid_for(h, n * 64, 256, 8);
        for (int it = 0; it <= g->iter; it++) {
            CK(cudaMemsetAsync(h->q.count, 0, sizeof(int), st));
            LAUNCH(h, k_trace_shadow, gs, 256, 0, st, it, sc, h->fc, h->sv, N, n_gshadow, (const int*)nullptr, out->surf_map, h->ldir, h->lsharp, L,
                   h->sr_g, h->q_g, h->cnt, out->lvis_map, 0, 1, h->pkt_search & 1, h->final_skip);
            // fp32 (reference-precision) mode chunks its GEMM chain on the host: the floor's list has tens of millions of entries of which
            // a few thousand are in the 5 mm shell, so the bound is the query count itself, read back per iteration (this mode synchronises
            // anyway; the tensor-core path reads every count on the device)
            int64_t n_q = h->sr_g.cap;
            if (it < g->iter && h->cfg.precision == RA_PRECISION_FP32 && read_counter(h, h->q.count, st, &n_q)) return 1;
            if (it < g->iter && distance_pass(h, st, n_q, &h->q_g)) return 1;
        }
    }
    LAUNCH(h, k_ground_shade, grid_for(h, F * 32, 256, 8), 256, 0, st, gc, 1, 0LL, (long long)F, h->g_weight, h->ldir, h->larea, L, h->g_light,
           out->lvis_map, out->ldot_map, out->albedo_map, c.shading_albedo, g->shading_multiplier, out->rgb_map, out->shade_map, out->spec_map);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_relight_ground(ra_handle* h, const ra_ground_config* g, const float* probe, int32_t ph, int32_t pw, const float* albedo_image,
                                 int32_t ih, int32_t iw, const float* ray_d, const float* albedo_in, float* lvis_map, float* ldot_map, int64_t F,
                                 float* rgb, float* albedo_out, float* shade, float* spec, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const ra_config& c = h->cfg;
    if (!c.relight) { h->err = "ra_relight_ground needs a relight handle"; return 1; }
    if (F == 0) return 0;
    const int L = c.env_h * c.env_w;
    GroundCfg gc = ground_cfg(h, g);
    if (!h->g_light) CK(dalloc(h, &h->g_light, (size_t)RA_NLIGHT_MAX * 3));
    LAUNCH(h, k_ground_light_table, 2, 256, 0, st, h->lxyz, L, probe, ph, pw, h->g_light);
    const float* alb = albedo_in;
    if (g->attach_envmap) {
        if (!albedo_out) { h->err = "ra_relight_ground: albedo_out required with attach_envmap"; return 1; }
        const float* img = albedo_image ? albedo_image : probe;
        if (!albedo_image) { ih = ph; iw = pw; }
        LAUNCH(h, k_ground_albedo, grid_for(h, F), 256, 0, st, ray_d, (long long)F, img, ih, iw, albedo_out);
        alb = albedo_out;
    } else if (albedo_out && albedo_out != albedo_in) {
        CK(cudaMemcpyAsync(albedo_out, albedo_in, (size_t)F * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    // the novel-light floor shade is sum / pi without shading_albedo and without the multiplier (:93-96)
    LAUNCH(h, k_ground_shade, grid_for(h, F * 32, 256, 8), 256, 0, st, gc, 0, 0LL, (long long)F, (const float*)nullptr, h->ldir, h->larea, L, h->g_light,
           lvis_map, ldot_map, alb, 1.f, 1.f, rgb, shade, spec);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_blend_ground(ra_handle* h, const float* acc_g, const float* ground, const float* human, int32_t human_premul, int32_t C,
                               int64_t F, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (F > h->pix_cap || !h->pix2ray) { h->err = "ra_blend_ground: call ra_ground_begin first"; return 1; }
    if (F == 0) return 0;
    LAUNCH(h, k_ground_blend, grid_for(h, F * C), 256, 0, st, h->pix2ray, acc_g, ground, human, human_premul, C, (long long)F, out);
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- render: volume (config 2)
extern "C" int ra_render_anisdf_volume(ra_handle* h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                                       int64_t P, const ra_outputs* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const ra_config& c = h->cfg;
    if (c.relight) { h->err = "volume renderer is the AniSDF path (relight=0)"; return 1; }
    if (!h->have_frame) { h->err = "render before ra_set_frame"; return 1; }
    if (zero_outputs(h, out, P, st)) return 1;
    CK(cudaMemsetAsync(h->counters_blk, 0, 16 * sizeof(int), st));
    OutMaps om{out->rgb_map, out->acc_map, out->depth_map, nullptr, out->norm_map, out->cpts_map, out->bpts_map, out->resd_map, nullptr, nullptr, nullptr};
    int S = c.vol_samples;
    for (int64_t r0 = 0; r0 < P; r0 += h->vol_rays) {
        int64_t nr = std::min<int64_t>(h->vol_rays, P - r0);
        CK(cudaMemsetAsync(h->al.count, 0, sizeof(int), st));
        CK(cudaMemsetAsync(h->raw, 0, (size_t)nr * S * 16 * sizeof(float), st));
        LAUNCH(h, k_attr_front, grid_for(h, nr * S, 256, 8), 256, 0, st, 2, h->fc, h->sv, c.n_verts, c.dist_th, c.blend_radius,
               (const float*)nullptr, (const float*)nullptr, 0LL, h->cnt.n_fg, h->fg_ray, h->surf, ray_o, ray_d, near_, far_, S,
               c.surf_sample_range, c.clip_near, c.clip_far, (long long)r0, (long long)nr, h->al, h->cnt, (h->pkt_search >> 2) & 1);
        if (attr_pass(h, st, nr * S, true)) return 1;
        LAUNCH(h, k_volume_blend, grid_for(h, nr * 16, 128, 16), 128, 0, st, h->raw, 16, S, near_, far_, c.clip_near, c.clip_far, (long long)r0, (long long)nr, om);
    }
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- point queries
extern "C" int ra_query_sdf(ra_handle* h, const float* x, int64_t n, float dist_th, int32_t smooth, float* sdf, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->have_frame) { h->err = "query before ra_set_frame"; return 1; }
    if (n > h->q_cap) { h->err = "too many query points"; return 1; }
    if (n == 0) return 0;
    CK(cudaMemsetAsync(h->q.count, 0, sizeof(int), st));
    LAUNCH(h, k_points_front, grid_for(h, n, 256, 8), 256, 0, st, h->fc, h->sv, h->cfg.n_verts, x, (int)n, dist_th, h->cfg.blend_radius,
           h->pt_smpl, h->pt_slot, h->q, h->cnt);
    if (distance_pass(h, st, n)) return 1;
    LAUNCH(h, k_points_finish, grid_for(h, n), 256, 0, st, h->pt_smpl, h->pt_slot, h->q.net, (int)n, dist_th, smooth, sdf);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_query_raw(ra_handle* h, const float* x, const float* v, int64_t n, float* raw, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->have_frame) { h->err = "query before ra_set_frame"; return 1; }
    if (n > h->attr_cap) { h->err = "too many query points"; return 1; }
    int C = h->cfg.relight ? 17 : 16;
    if (n == 0) return 0;
    CK(cudaMemsetAsync(h->al.count, 0, sizeof(int), st));
    CK(cudaMemsetAsync(h->raw, 0, (size_t)n * C * sizeof(float), st));
    LAUNCH(h, k_attr_front, grid_for(h, n, 256, 8), 256, 0, st, 0, h->fc, h->sv, h->cfg.n_verts, h->cfg.dist_th, h->cfg.blend_radius, x, v,
           (long long)n, h->cnt.n_fg, h->fg_ray, h->surf, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr,
           (const float*)nullptr, 1, 0.f, 0.f, 0.f, 0LL, 0LL, h->al, h->cnt, 0);
    if (attr_pass(h, st, n, true)) return 1;
    CK(cudaMemcpyAsync(raw, h->raw, (size_t)n * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int ra_get_stats(ra_handle* h, ra_stats* out) {
    CK(cudaDeviceSynchronize());
    int c[16];
    CK(cudaMemcpy(c, h->counters_blk, sizeof(c), cudaMemcpyDeviceToHost));
    out->n_rays = h->last_P;
    out->n_fg = c[0]; out->n_shadow_rays = c[7]; out->n_attr_samples = c[2];
    unsigned long long q[2];
    memcpy(q, &c[8], 16);
    out->n_queries = (int64_t)q[0]; out->n_queries_in_shell = (int64_t)q[1];
    out->n_dropped_shadow_rays = c[6];
    out->n_shadow_slots = (int64_t)c[1] + c[5];
    return 0;
}

extern "C" int ra_profile_enable(ra_handle* h, int32_t on) {
    h->prof = on != 0;
    h->ev_mlp_used = 0; h->ev_stage_used = 0;
    return 0;
}

extern "C" int ra_profile_read(ra_handle* h, double* mlp_ms, int64_t* mlp_launches, double* stage_ms) {
    CK(cudaDeviceSynchronize());
    double ms = 0;
    for (size_t i = 0; i + 1 < h->ev_mlp_used; i += 2) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, h->ev_mlp[i], h->ev_mlp[i + 1]));
        ms += t;
    }
    if (mlp_ms) *mlp_ms = ms;
    if (mlp_launches) *mlp_launches = (int64_t)(h->ev_mlp_used / 2);
    if (stage_ms) {
        for (int k = 0; k < 4; k++) stage_ms[k] = 0;
        for (size_t i = 0; i + 4 < h->ev_stage_used; i += 5)
            for (int k = 0; k < 4; k++) {
                float t = 0;
                CK(cudaEventElapsedTime(&t, h->ev_stage[i + k], h->ev_stage[i + k + 1]));
                stage_ms[k] += t;
            }
    }
    h->ev_mlp_used = 0; h->ev_stage_used = 0;
    return 0;
}

#ifdef RA_KNN_STATS
extern "C" int ra_debug_knn_stats(unsigned long long* out12, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out12, g_knn_stats, sizeof(unsigned long long) * 12);
    if (reset) { unsigned long long z[12] = {0}; cudaMemcpyToSymbol(g_knn_stats, z, sizeof(z)); }
    return 0;
}
#endif

// debug: per-layer clock64 timeline of CTA 0 of the fused MLP kernel (tools/tc_timeline.py)
extern "C" int ra_debug_tc_timeline(ra_handle* h, unsigned long long* out, int enable) {
    if (enable) {
        if (!h->tc.dbg) { CK(cudaMalloc((void**)&h->tc.dbg, 512 * sizeof(unsigned long long))); h->tc2.dbg = h->tc.dbg; h->tc8.dbg = h->tc.dbg; }
        CK(cudaMemset(h->tc.dbg, 0, 512 * sizeof(unsigned long long)));
        return 0;
    }
    CK(cudaDeviceSynchronize());
    // 18 x 8 entries for the single-CTA kernel; 36 x 8 ((layer, slot) x 8) for the pair kernel k_mlp_tc6
    if (h->tc.dbg && out) CK(cudaMemcpy(out, h->tc.dbg, (h->tc_variant >= 6 ? 512 : 18 * 8) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

/* rotate_envmap (relight_utils.py:55-103) for a sweep: out (n_rot,16,32,3) = probe shifted by j0..j0+n_rot-1 of env_w*repeat steps */
extern "C" int ra_rotate_probes(ra_handle* h, const float* probe, int32_t repeat, int32_t j0, int32_t n_rot, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (repeat <= 0 || n_rot <= 0) { h->err = "ra_rotate_probes: repeat and n_rot must be positive"; return 1; }
    LAUNCH(h, k_shift_probe, grid_for(h, (long long)n_rot * h->cfg.env_h * h->cfg.env_w), 256, 0, st, probe, h->cfg.env_h, h->cfg.env_w, repeat, j0, n_rot, out);
    CK(cudaGetLastError());
    return 0;
}

/* Visualizer.generate_image's ray -> image scatter (base_visualizer.py:182-202) + alpha channel + optional 8-bit quantisation.
 * mask_at_box: H*W bytes (0/1); rgb_map (P,3), acc_map (P) in ray order; out_f (H,W,4) float and/or out_u8 (H,W,4) may be NULL. */
extern "C" int ra_assemble_image(ra_handle* h, const float* rgb_map, const float* acc_map, const unsigned char* mask_at_box, int32_t H,
                                 int32_t W, float bg_brightness, float* out_f, unsigned char* out_u8, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int n = H * W, nb = (n + 255) / 256;
    h->g_W = W; h->g_H = H;
    if (nb > h->blk_cap) { hfree(h, h->blk_cnt); CK(dalloc(h, &h->blk_cnt, (size_t)nb)); h->blk_cap = nb; }
    LAUNCH(h, k_mask_count, nb, 256, 0, st, mask_at_box, n, h->blk_cnt);
    LAUNCH(h, k_scan_blocks, 1, 1024, 0, st, h->blk_cnt, nb);
    LAUNCH(h, k_assemble, nb, 256, 0, st, mask_at_box, n, h->blk_cnt, rgb_map, acc_map, bg_brightness, out_f, out_u8);
    CK(cudaGetLastError());
    return 0;
}

/* ---- row f3 (remainder): Visualizer.generate_image on the device ------------------------------------------------------------ */
/* rotate_envmap's shift_image for an image of any size (the env-map image attached to the floor, relight_utils.py:74-75,103) */
extern "C" int ra_rotate_image(ra_handle* h, const float* image, int32_t H, int32_t W, double step, int32_t j0, int32_t n_rot, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (H <= 0 || W <= 0 || n_rot <= 0) { h->err = "ra_rotate_image: bad arguments"; return 1; }
    LAUNCH(h, k_shift_image, grid_for(h, (long long)n_rot * H * W), 256, 0, st, image, H, W, step, j0, n_rot, out);
    CK(cudaGetLastError());
    return 0;
}

// k-th smallest (largest == 0) / largest element of the selection domain -> h->kth[slot].result (stays on the device)
static int kth_select(ra_handle* h, int slot, int mode, int largest, const float* a, const float* b, long long n, long long k, cudaStream_t st) {
    if (!h->kth) CK(dalloc(h, &h->kth, 2));
    if (k <= 0 || k > n) { h->err = "percentile of too few elements (the reference's topk(0).max() raises as well)"; return 1; }
    KthState* s = h->kth + slot;
    LAUNCH(h, k_kth_init, 1, 256, 0, st, s, (unsigned)k);
    for (int pass = 0; pass < 4; pass++) {
        LAUNCH(h, k_kth_hist, grid_for(h, n, 256, 4), 256, 0, st, s, pass, mode, largest, a, b, n);
        LAUNCH(h, k_kth_pick, 1, 32, 0, st, s, pass, largest);
    }
    return 0;
}

extern "C" int ra_visual_map(ra_handle* h, int32_t type, const ra_visual_inputs* in, int64_t n, const ra_visual_config* vc, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!in || !vc || !out) { h->err = "ra_visual_map: null argument"; return 1; }
    if (n == 0) return 0;
    VisualIn v{in->rgb_map, in->acc_map, in->norm_map, in->depth_map, in->shade_map, in->albedo_map, in->roughness_map, in->cpts_map, in->bpts_map,
               in->surf_map, in->spec_map, in->cam_R, in->tbounds, nullptr, nullptr, vc->min_clip, 0, vc->tonemapping_albedo};
    auto need = [&](const void* p, const char* what) { if (!p) { h->err = std::string("ra_visual_map: ") + what + " is required for this output type"; return true; } return false; };
    switch (type) {
    case RA_VIS_RENDERING: if (need(in->rgb_map, "rgb_map")) return 1; break;
    case RA_VIS_NORMAL: if (need(in->norm_map, "norm_map") || need(in->acc_map, "acc_map") || need(in->cam_R, "cam_R")) return 1; break;
    case RA_VIS_ALPHA: if (need(in->acc_map, "acc_map")) return 1; break;
    case RA_VIS_DEPTH: {
        if (need(in->depth_map, "depth_map") || need(in->acc_map, "acc_map")) return 1;
        const long long k = (long long)(0.01 * (double)n);         // int(percentile * depth_map.numel())   :106
        if (kth_select(h, 0, 2, 0, in->depth_map, in->acc_map, n, k, st) || kth_select(h, 1, 2, 1, in->depth_map, in->acc_map, n, k, st)) return 1;
    } break;
    case RA_VIS_SHADING: case RA_VIS_SPECULAR: {
        const float* m = type == RA_VIS_SHADING ? in->shade_map : in->spec_map;
        if (need(m, type == RA_VIS_SHADING ? "shade_map" : "spec_map")) return 1;
        v.normalize = type == RA_VIS_SHADING ? vc->normalize_shading : vc->normalize_specular;
        if (v.normalize && kth_select(h, 1, 0, 1, m, nullptr, n * 3, (long long)(0.005 * (double)(n * 3)), st)) return 1;
    } break;
    case RA_VIS_ALBEDO: if (need(in->albedo_map, "albedo_map")) return 1; break;
    case RA_VIS_ROUGHNESS: if (need(in->roughness_map, "roughness_map")) return 1; break;
    case RA_VIS_SURFACE: if (need(in->cpts_map ? in->cpts_map : in->surf_map, "cpts_map or surf_map") || need(in->acc_map, "acc_map") || need(in->tbounds, "tbounds")) return 1; break;
    case RA_VIS_RESIDUAL:
        if (need(in->cpts_map, "cpts_map") || need(in->bpts_map, "bpts_map") || need(in->acc_map, "acc_map")) return 1;
        if (kth_select(h, 1, 1, 1, in->cpts_map, in->bpts_map, n * 3, (long long)(0.005 * (double)(n * 3)), st)) return 1;
        break;
    default: h->err = "ra_visual_map: unknown output type"; return 1;
    }
    v.lo = h->kth; v.hi = h->kth ? h->kth + 1 : nullptr;
    LAUNCH(h, k_visual_map, grid_for(h, n), 256, 0, st, type, v, (long long)n, out);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ra_assemble_visual(ra_handle* h, const float* map, const float* acc_map, const unsigned char* mask_at_box, int32_t H, int32_t W,
                                  const ra_image_config* ic, float* out_f, unsigned char* out_u8, unsigned short* out_u16, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!ic || (ic->channels != 3 && ic->channels != 4)) { h->err = "ra_assemble_visual: channels must be 3 or 4"; return 1; }
    if (ic->probe && (!ic->probe_dirs || ic->uH > H || ic->uW > W)) { h->err = "ra_assemble_visual: probe overlay needs probe_dirs and uH <= H, uW <= W"; return 1; }
    int n = H * W, nb = (n + 255) / 256;
    h->g_W = W; h->g_H = H;
    if (nb > h->blk_cap) { hfree(h, h->blk_cnt); CK(dalloc(h, &h->blk_cnt, (size_t)nb)); h->blk_cap = nb; }
    LAUNCH(h, k_mask_count, nb, 256, 0, st, mask_at_box, n, h->blk_cnt);
    LAUNCH(h, k_scan_blocks, 1, 1024, 0, st, h->blk_cnt, nb);
    AssembleArgs a{mask_at_box, n, W, h->blk_cnt, map, acc_map, ic->bg_brightness, ic->channels, ic->bgr, ic->probe, ic->eh, ic->ew,
                   ic->probe_dirs, ic->probe ? ic->uH : 0, ic->probe ? ic->uW : 0, out_f, out_u8, out_u16};
    LAUNCH(h, k_assemble2, nb, 256, 0, st, a);
    CK(cudaGetLastError());
    return 0;
}

/* cfg.replace_light (sphere_tracing_renderer.py:1068-1069): the main pass is lit by this env-map instead of the learned one.
 * probe: device pointer (ph,pw,3), copied; NULL restores the learned light. */
extern "C" int ra_set_main_light(ra_handle* h, const float* probe, int32_t ph, int32_t pw, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!probe) { h->main_light = nullptr; return 0; }
    if (ph <= 0 || pw <= 0) { h->err = "ra_set_main_light: bad probe size"; return 1; }
    static_assert(sizeof(float) == 4, "");
    float* buf = nullptr;
    CK(dalloc(h, &buf, (size_t)ph * pw * 3));
    CK(cudaMemcpyAsync(buf, probe, (size_t)ph * pw * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    float* old = h->main_light;
    h->main_light = buf; h->mlh = ph; h->mlw = pw;
    if (old) { CK(cudaStreamSynchronize(st)); hfree(h, old); }
    return 0;
}

/* The exact 3-NN of row a4 on its own (pytorch3d.ops.knn_points, K=3; sample_utils.py:122): x (n,3) world points ->
 * ids (n,3) original vertex indices nearest first, d2 (n,3) squared pose-space distances.  Same search the renderers use. */
static int query_knn(ra_handle* h, const float* x, int64_t n, int32_t* ids, float* d2, int packets, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->have_frame) { h->err = "query before ra_set_frame"; return 1; }
    if (n == 0) return 0;
    LAUNCH(h, k_points_knn, grid_for(h, n, 256, 8), 256, 0, st, h->fc, h->sv, h->cfg.n_verts, x, (int)n, (int*)ids, d2, packets);
    CK(cudaGetLastError());
    return 0;
}
/* The same with the points taken as packets of 32 consecutive entries (the shadow tracer's warps: parallel rays of neighbouring
 * pixels): far-field lanes of a packet search the box hierarchy together (hdq.cuh: knn3_packet).  Any input is answered exactly;
 * the grouping only decides which search runs. */
extern "C" int ra_query_knn_packets(ra_handle* h, const float* x, int64_t n, int32_t* ids, float* d2, void* stream) {
    return query_knn(h, x, n, ids, d2, 1, stream);
}
extern "C" int ra_query_knn(ra_handle* h, const float* x, int64_t n, int32_t* ids, float* d2, void* stream) {
    return query_knn(h, x, n, ids, d2, 0, stream);
}

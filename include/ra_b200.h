/*
 * ra_b200.h -- C-ABI of the B200-native RelightableAvatar inference renderer.
 *
 * Drop-in boundary: the reference selects its renderer with
 *   importlib.import_module(cfg.renderer_module).Renderer(network)      lib/networks/renderer/make_renderer.py:5-8
 * and calls `renderer.render(batch) -> dotdict` once per frame             run.py:45,80
 * The Python plugin `relightableavatar_b200.renderer.Renderer` mirrors that surface and forwards
 * to the entry points below through ctypes.  No torch types cross this boundary: plain pointers
 * and sizes only.  All tensors are fp32, row-major, with the reference's batch dimension B=1
 * squeezed away.  Pointers are DEVICE pointers unless the parameter says "host or device".
 *
 * Conventions: every function returns 0 on success, non-zero on error (message via ra_last_error);
 * no exceptions cross the ABI; the caller allocates every output; the library owns only its
 * workspace and its packed weight copies; all work is enqueued on the passed cudaStream_t
 * (void*), no internal device synchronisation on the product path; one handle per device.
 */
#ifndef RA_B200_H
#define RA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ra_handle ra_handle;

enum { RA_PRECISION_FP32 = 0,   /* CUDA-core fp32 MLPs everywhere (reference precision) */
       RA_PRECISION_TC = 1 };   /* distance-query MLPs on tcgen05 (fp16 operands, fp32 accumulate) */

/* Values the reference reads from its global `cfg` (lib/config/config.py; SURVEY.md 8 notation). */
typedef struct ra_config {
    int32_t relight;            /* 1: relight_network (raw 17 ch), 0: AniSDF base_network (raw 16 ch) */
    int32_t precision;          /* RA_PRECISION_* */
    int32_t max_rays;           /* capacity: max P of any render call */
    int32_t n_verts, n_bones;   /* 6890; 52 (SMPL-H, xuzhen_12v_geo) or 24 (SMPL subjects).  The pose-condition width of the residual and
                                   render MLPs is 3 * n_bones (cfg.cond_dim, config.py:465-466): 156 or 72 */
    float dist_th;              /* net.dist_th: 0.125 relight / 0.1 AniSDF      base_network.py:177 */
    float blend_radius;         /* 0.075                                        config.py:191 */
    float resd_limit;           /* 0.05                                         config.py:224 */
    int32_t st_iter;            /* 16    cfg.sphere_tracing.*                   config.py:116-124 */
    float st_tan_i, st_relax, st_offset, st_eps;
    int32_t st_skip;
    int32_t lv_iter;            /* 4     cfg.obj_lvis.*                         config.py:127-132 */
    float lv_offset, lv_relax, lv_near, lv_dist_th;
    float env_r;                /* 10: far of shadow rays                        config.py:113 */
    float bbox_margin;          /* 0.25, applied per pixel chunk                 sphere_tracing_renderer.py:1020-1022 */
    int32_t render_chunk;       /* 65536                                         base.yaml:170 */
    int32_t n_samples;          /* 3 surface samples                             base.yaml:169 */
    float surf_sample_range;    /* 0.005                                         config.py:76 */
    float fresnel_f0;           /* 0.02 */
    float albedo_slope, albedo_bias, rough_slope, rough_bias, albedo_multiplier, shading_albedo;
    int32_t env_h, env_w;       /* 16, 32 */
    int32_t vol_samples;        /* 128: base_renderer uniform samples            base.yaml:78 */
    float clip_near, clip_far;  /* 0.02, 10                                      config.py:79-80 */
    int32_t visibility_mode;    /* 0: DFSS soft-shadow tracing; 1: cfg.local_visibility (lvis = n.l > 0); 2: cfg.no_visibility (lvis = 1)
                                   sphere_tracing_renderer.py:296-301 */
    int32_t brdf_mode;          /* 0: glossy + Lambert; 1: cfg.lambert_only; 2: cfg.glossy_only     relight_utils.py:563-568 */
    int32_t tonemapping;        /* 1: the main pass's rgb passes linear2srgb (cfg.tonemapping_rendering, config.py:417); 0 for
                                   .exr/.hdr output (config.py:446-448)   sphere_tracing_renderer.py:523,731.  The novel-light
                                   re-shade maps unconditionally, as the reference does (novel_light_sphere_tracing.py:48,94) */
} ra_config;

/* Network tensors in torch layout (out_features x in_features), weight-norm already folded
 * (w = g * v / ||v||_row) by the caller.  "host or device" pointers; the library copies them.
 * State-dict keys: SURVEY.md 8b. */
typedef struct ra_weights {
    const float* resd_w[9]; const float* resd_b[9];     /* residual_deformation_network.mlp.linears.{l}: layer 0 is
                                                           (256, 63 + 3 n_bones), layer 4 (256, 256 + 63 + 3 n_bones) */
    const float* sdf_w[9];  const float* sdf_b[9];      /* signed_distance_network.mlp.lin{l} (folded)  */
    float sdf_beta;                                     /* clamp(_beta, 1e-9, 1e6) */
    const float* render_w[5]; const float* render_b[5]; /* render_network.l{l} (folded); NULL if absent; l3 is (256, 256 + 3 n_bones) */
    const float* albedo_w[3]; const float* albedo_b[3]; /* albedo_network.linears.{l}; NULL for AniSDF   */
    const float* rough_w[3];  const float* rough_b[3];  /* roughness_network.linears.{l}                 */
    const float* env_main;  int32_t env_main_h, env_main_w;  /* softplus(global_env_map_) expanded to 3 ch */
    const float* light_xyz;   /* (env_h, env_w, 3) */
    const float* light_area;  /* (env_h, env_w)    */
    const float* light_sharp; /* (env_h, env_w)    */
} ra_weights;

/* Per-frame tensors of `batch` (lib/datasets/pose_dataset.py:45-113, base_dataset.py:337-397).
 * Device pointers, borrowed until the next ra_set_frame. */
typedef struct ra_frame {
    const float* R;        /* (3,3)    */
    const float* Th;       /* (3)      */
    const float* poses;    /* (3 n_bones) = (156) */
    const float* A;        /* (n_bones,4,4) */
    const float* big_A;    /* (n_bones,4,4) */
    const float* weights;  /* (N,n_bones) */
    const float* pverts;   /* (N,3)    */
    const float* pnorm;    /* (N,3)    */
    const float* tverts;   /* (N,3) big-pose vertices */
    const float* wbounds;  /* (2,3)  world AABB of the posed body (+-0.05); read, not modified */
    const float* mat_cond; /* (3 n_bones) train_motion.poses[fix_material]; may be NULL for relight */
} ra_frame;

/* Per-ray output maps, each (P, C) fp32, premultiplied by acc like the reference's alpha_output_
 * (sphere_tracing_renderer.py:454-460,1113).  NULL pointers are skipped. */
typedef struct ra_outputs {
    float* rgb_map;        /* (P,3) */
    float* acc_map;        /* (P)   */
    float* depth_map;      /* (P)   */
    float* surf_map;       /* (P,3) */
    float* norm_map;       /* (P,3) */
    float* cpts_map;       /* (P,3) */
    float* bpts_map;       /* (P,3) */
    float* resd_map;       /* (P,3) */
    float* albedo_map;     /* (P,3) relight only */
    float* roughness_map;  /* (P)   relight only */
    float* shade_map;      /* (P,3) relight only */
    float* lvis_map;       /* (P,512) optional */
    float* ldot_map;       /* (P,512) optional */
    float* spec_map;       /* (P,3) relight only, optional: the main pass's specular view (cfg.vis_specular_map, sphere_tracing_renderer.py:739-748) */
} ra_outputs;

/* Work counters of the last render call (device-side counts read back on request; forces a sync). */
typedef struct ra_stats {
    int64_t n_rays;            /* P */
    int64_t n_fg;              /* S_fg: pixels with acc > 0 */
    int64_t n_shadow_rays;     /* traced (pixel, light) pairs */
    int64_t n_queries;         /* HDQ distance queries issued (all iterations) */
    int64_t n_queries_in_shell;/* ... of which went through the MLPs */
    int64_t n_attr_samples;    /* in-shell surface/volume samples (fwd + input-gradient) */
    int64_t n_dropped_shadow_rays; /* shadow rays beyond the workspace (256 per ray of max_rays): 0 for the reference's 16x32 light grid,
                                      whose antipodal symmetry puts at most L/2 lights in front of a pixel; non-zero = lvis incomplete */
    int64_t n_shadow_slots;    /* entries of the shadow-ray list incl. the padding of partly filled 32-ray packets (human pass + last floor batch) */
} ra_stats;

int  ra_create(ra_handle** out, const ra_config* cfg);
void ra_destroy(ra_handle* h);
const char* ra_last_error(ra_handle* h);

int ra_upload_weights(ra_handle* h, const ra_weights* w, void* stream);
int ra_set_frame(ra_handle* h, const ra_frame* f, void* stream);

/* sphere_tracing_renderer.Renderer.render for the relight network (a1-a18, a21). */
int ra_render_relight(ra_handle* h, const float* ray_o, const float* ray_d, const float* near, const float* far,
                      int64_t P, const ra_outputs* out, void* stream);
/* Tile sharding (BASELINE config 4): this handle renders only the rays of rank `rank` of `world`, dealt in interleaved blocks of
 * `block` rays out of the frame's `global_P` in-box rays.  The one place the reference's result depends on a ray's position in
 * the frame is the per-chunk wbounds growth (sphere_tracing_renderer.py:1020-1022, chunks of cfg.render_chunk_size rays): with the
 * layout set, the shadow-ray box uses the ray's GLOBAL index, so sharded and unsharded frames agree bit for bit.
 * world == 1 restores the default. */
int ra_set_ray_layout(ra_handle* h, int64_t global_P, int32_t block, int32_t world, int32_t rank);
/* The single collective of a sharded step: all-gather of every rank's finished, padded pixel block over
 * NCCL / NVLink.  comm is an `ncclComm_t` (passed as void* so that this header needs no nccl.h); send: n_floats fp32, recv:
 * world * n_floats.  libnccl.so.2 is resolved at the first call (dlopen), the library has no link-time NCCL dependency; the
 * Python mirror uses torch.distributed.all_gather_into_tensor for the same exchange (relightableavatar_b200/parallel.py). */
int ra_allgather(ra_handle* h, void* comm, const float* send, int64_t n_floats, float* recv, void* stream);
/* cfg.replace_light (sphere_tracing_renderer.py:1068-1069): `probe` (ph,pw,3, device, copied) lights the main pass of the following
 * ra_render_relight calls instead of the learned env-map; NULL restores the learned one. */
int ra_set_main_light(ra_handle* h, const float* probe, int32_t ph, int32_t pw, void* stream);
/* novel_light_sphere_tracing per-env-map re-shade (a19): probes (n_env,16,32,3); rgb/shade/spec (n_env,P,3).
 * Uses the maps of the preceding ra_render_relight call (kept in the workspace). */
int ra_relight_envmaps(ra_handle* h, const float* probes, int32_t n_env, float* rgb, float* shade, float* spec,
                       void* stream);
/* Env-map rotation sweep (vis_rotate_light; relight_utils.py:55-103 rotate_envmap / shift_image, novel_light_sphere_tracing.py:163-171):
 * out (n_rot, env_h, env_w, 3) = probe (env_h, env_w, 3) shifted by j0 .. j0+n_rot-1 steps of 1/repeat texel along the longitude;
 * feed the result to ra_relight_envmaps. */
int ra_rotate_probes(ra_handle* h, const float* probe, int32_t repeat, int32_t j0, int32_t n_rot, float* out, void* stream);
/* Image assembly (lib/visualizers/base_visualizer.py:182-202): img = bg; img[mask_at_box] = rgb_map; alpha[mask_at_box] = acc_map.
 * mask_at_box: H*W bytes; out_f (H,W,4) fp32 and/or out_u8 (H,W,4) = clip(.,0,1)*255; either may be NULL. */
int ra_assemble_image(ra_handle* h, const float* rgb_map, const float* acc_map, const unsigned char* mask_at_box, int32_t H, int32_t W,
                      float bg_brightness, float* out_f, unsigned char* out_u8, void* stream);
/* ---- Visualizer.generate_image on the device ---------------------------------------------------------
 * Reference: lib/visualizers/base_visualizer.py:55-231 (generate_image: the per-type map transforms, the ray -> image scatter,
 * the light-probe overlay, the alpha channel), lib/utils/relight_utils.py:38-52 (add_light_probe), lib/utils/data_utils.py:689-709
 * (save_image: BGR order, 16-bit png / 8-bit jpg quantisation).  A frame leaves the device as finished pixels. */
enum { RA_VIS_RENDERING = 0, RA_VIS_NORMAL = 1, RA_VIS_ALPHA = 2, RA_VIS_DEPTH = 3, RA_VIS_SHADING = 4, RA_VIS_ALBEDO = 5,
       RA_VIS_ROUGHNESS = 6, RA_VIS_SURFACE = 7, RA_VIS_RESIDUAL = 8, RA_VIS_SPECULAR = 9 };     /* Output (lib/config/config.py:364-378) */
typedef struct ra_visual_inputs {   /* one light's maps in ray order, (n,3) or (n); only what the requested type reads must be non-NULL */
    const float *rgb_map, *acc_map, *norm_map, *depth_map, *shade_map, *albedo_map, *roughness_map, *cpts_map, *bpts_map, *surf_map, *spec_map;
    const float* cam_R;      /* (3,3) batch.cam_R, device (Normal) */
    const float* tbounds;    /* (2,3) batch.tbounds, device (Surface) */
} ra_visual_inputs;
typedef struct ra_visual_config {
    float min_clip;              /* cfg.min_clip 1.0              config.py:46 */
    int32_t normalize_shading;   /* cfg.normalize_shading False   config.py:41 */
    int32_t normalize_specular;  /* cfg.normalize_specular True   config.py:42 */
    int32_t tonemapping_albedo;  /* cfg.tonemapping_albedo True   config.py:416 */
} ra_visual_config;
/* generate_image's rgb_map for one Output type: out (n,3) in ray order.  The percentiles of Depth / Residual / normalised
 * Shading / Specular are exact k-th order statistics (the reference's topk(k)[0].max() / .min()) selected on the device. */
int ra_visual_map(ra_handle* h, int32_t type, const ra_visual_inputs* in, int64_t n, const ra_visual_config* vc, float* out, void* stream);
typedef struct ra_image_config {
    float bg_brightness;         /* cfg.bg_brightness */
    int32_t channels;            /* 4 with cfg.store_alpha_channel, else 3 */
    int32_t bgr;                 /* save_image's RGB -> BGR swap for cv2 */
    const float* probe;          /* (eh,ew,3) env-map shown in the top-left corner (cfg.probe_size_ratio > 0), or NULL */
    int32_t eh, ew;
    const float* probe_dirs;     /* (uH,uW,3) world-space directions of the overlay pixels = gen_light_dir(uH, uW, cam_R)  relight_utils.py:9-35 */
    int32_t uH, uW;              /* uW = int(W * probe_size_ratio), uH = int(uW * eh / ew) */
} ra_image_config;
/* img = bg; img[mask_at_box] = map; overlay; alpha[mask_at_box] = acc_map -> (H,W,channels) as fp32 and / or
 * (v*255).clip(0,255) uint8 and / or (v*65535).clip(0,65535) uint16 (any of the three outputs may be NULL). */
int ra_assemble_visual(ra_handle* h, const float* map, const float* acc_map, const unsigned char* mask_at_box, int32_t H, int32_t W,
                       const ra_image_config* ic, float* out_f, unsigned char* out_u8, unsigned short* out_u16, void* stream);
/* rotate_envmap's shift_image for an image of any size (the env-map image attached to the floor rotates with the probe,
 * relight_utils.py:74-75,103): out (n_rot,H,W,3) = image shifted by step * (j0 .. j0+n_rot-1) texels, step = iW / (env_w * repeat). */
int ra_rotate_image(ra_handle* h, const float* image, int32_t H, int32_t W, double step, int32_t j0, int32_t n_rot, float* out, void* stream);

/* ---- ground-plane shading (cfg.vis_ground_shading) -------------------------------------------------
 * Reference: sphere_tracing_renderer.py:463-548 (render_ground), :1079-1111 (ground branch of Renderer.render), :395-451
 * (blend_output_), novel_light_sphere_tracing.py:69-98,191-212 (per-env-map floor re-shade + blend), cfg.env_lvis
 * (config.py:135-141), cfg.ground_* (config.py:45,104-107,353).  All maps here are image-sized: F = H*W pixels. */
typedef struct ra_ground_config {
    float normal[3];            /* cfg.ground_normal  (0,0,1) */
    float origin[3];            /* cfg.ground_origin  (0,0,0) */
    float albedo[3];            /* cfg.ground_albedo  (.05,.05,.05); used when attach_envmap == 0 */
    int32_t attach_envmap;      /* cfg.ground_attach_envmap (1): floor albedo = env-map colour along the view ray */
    float shading_multiplier;   /* cfg.ground_shading_multiplier (1.0) */
    int32_t iter;               /* cfg.env_lvis.iter 16 */
    float offset, relax, near_offset, dist_th;   /* .01, 0, .02, .005 */
} ra_ground_config;

typedef struct ra_ground_outputs {  /* every pointer (F, C) fp32; rgb/surf/albedo/lvis/ldot are required, the rest optional */
    float* rgb_map;        /* (F,3) */
    float* surf_map;       /* (F,3) */
    float* albedo_map;     /* (F,3) */
    float* roughness_map;  /* (F)   all ones */
    float* spec_map;       /* (F,3) shade / 20 */
    float* norm_map;       /* (F,3) */
    float* shade_map;      /* (F,3) */
    float* depth_map;      /* (F)   */
    float* lvis_map;       /* (F,512) far-field-blended visibility, kept by the caller for ra_relight_ground */
    float* ldot_map;       /* (F,512) */
} ra_ground_outputs;

/* `inds` / ground acc of Renderer.render (:1083-1088): mask_at_box (H*W bytes) + acc_map (P, from ra_render_relight) ->
 * acc_g (F) = 1 - acc scattered to the mask pixels (1 elsewhere).  Also records the pixel -> ray map used by ra_blend_ground.
 * The ray order is mask.nonzero() -- what batch_aware_indexing's topk(sorted=False) is relied upon to return. */
int ra_ground_begin(ra_handle* h, const unsigned char* mask_at_box, int32_t H, int32_t W, const float* acc_map, float* acc_g, void* stream);
/* render_ground over all F pixels after a ra_render_relight call (continues its per-chunk wbounds growth): ray_o, ray_d (F,3)
 * from get_rays(H, W, K, R, T) (net_utils.py:403-420); probe (ph,pw,3) lights the floor; albedo image (ih,iw,3) = envmap.image
 * if present else the probe (ignored when attach_envmap == 0). */
int ra_render_ground(ra_handle* h, const ra_ground_config* g, const float* ray_o, const float* ray_d, const float* acc_g, int64_t F,
                     const float* probe, int32_t ph, int32_t pw, const float* albedo_image, int32_t ih, int32_t iw,
                     const ra_ground_outputs* out, void* stream);
/* novel_light_sphere_tracing.render_ground for one env-map: re-shade the floor from its stored maps.
 * albedo_out receives the re-sampled albedo when attach_envmap, else a copy of albedo_in. */
int ra_relight_ground(ra_handle* h, const ra_ground_config* g, const float* probe, int32_t ph, int32_t pw, const float* albedo_image,
                      int32_t ih, int32_t iw, const float* ray_d, const float* albedo_in, float* lvis_map, float* ldot_map, int64_t F,
                      float* rgb, float* albedo_out, float* shade, float* spec, void* stream);
/* blend_output_ for one key of C channels: out (F,C) = ground * acc_g + scatter(human) * (1 - acc_g).
 * ground NULL: the acc_map rule (target 0); human NULL: alpha_times.  human_premul != 0: `human` (P,C) is already
 * multiplied by acc (the maps ra_render_relight returns), so it is added as is. */
int ra_blend_ground(ra_handle* h, const float* acc_g, const float* ground, const float* human, int32_t human_premul, int32_t C,
                    int64_t F, float* out, void* stream);
/* ra_relight_envmaps with raw (not acc-premultiplied) inputs -- what the reference feeds render_human when ground shading
 * is on (alpha_output_ is skipped, :1109-1113); outputs are multiplied by acc so that ra_blend_ground(human_premul=1) applies. */
int ra_relight_envmaps_raw(ra_handle* h, const float* probes, int32_t n_env, float* rgb, float* shade, float* spec, void* stream);

/* ---- per-frame batch preparation on the GPU ------------------------------------------------------
 * Replaces the CPU work of pose_dataset.__getitem__ (lib/datasets/pose_dataset.py:45-113): get_lbs_params / get_blend
 * (base_dataset.py:308-397) and get_rays_within_bounds (lib/utils/data_utils.py:925-938).  The body is uploaded once;
 * per frame only poses / Rh / Th / camera cross PCIe. */
typedef struct ra_body {        /* "host or device" pointers, copied by the library */
    const float* tjoints;       /* (J,3) rest joints                      base_dataset.py:205,219 */
    const int32_t* parents;     /* (J)   kinematic tree, -1 for the root  base_dataset.py:213 */
    const float* rverts;        /* (N,3) rest-pose vertices */
    const float* rnorm;         /* (N,3) rest-pose normals, or NULL when `faces` is given */
    const int32_t* faces;       /* (n_faces,3) or NULL: vertex normals of the posed mesh (pytorch3d verts_normals, base_dataset.py:380-381) */
    int32_t n_faces;
    const float* weights;       /* (N,J) skinning weights */
} ra_body;
int ra_upload_body(ra_handle* h, const ra_body* body, void* stream);

typedef struct ra_pose_outputs {   /* device pointers; NULL entries are skipped except A, R, pverts, pnorm (needed by ra_set_frame) */
    float* A;        /* (J,4,4)  get_rigid_transform (net_utils.py:1163-1172) */
    float* R;        /* (3,3)    cv2.Rodrigues(Rh) */
    float* pverts;   /* (N,3)    tpose_points_to_pose_points(rverts, weights, A)  blend_utils.py:303-313 */
    float* pnorm;    /* (N,3) */
    float* wverts;   /* (N,3)    pose_points_to_world_points                       blend_utils.py:264-273 */
    float* wnorm;    /* (N,3) */
    float* pbounds;  /* (2,3)    get_bounds(pverts), +-bounds_pad                  data_utils.py:1241-1248 */
    float* wbounds;  /* (2,3) */
} ra_pose_outputs;
/* poses (J*3 axis-angle), Rh (3), Th (3): DEVICE pointers (the only per-frame upload, < 1 KB). */
int ra_prepare_pose(ra_handle* h, const float* poses, const float* Rh, const float* Th, float bounds_pad, const ra_pose_outputs* out, void* stream);
/* get_rays_within_bounds: K, R (3,3), T (3) are HOST pointers (camera, 21 floats); wbounds (2,3) device.  Outputs have capacity
 * H*W rays; the rays of the pixels whose ray hits the box are written compacted in row-major pixel order, mask_at_box is
 * H*W bytes, *n_rays (device int) receives P. */
int ra_prepare_rays(ra_handle* h, const float* K, const float* R, const float* T, int32_t H, int32_t W, const float* wbounds,
                    float* ray_o, float* ray_d, float* near, float* far, unsigned char* mask_at_box, int32_t* n_rays, void* stream);

/* sphere_tracing_renderer.Renderer.render for the AniSDF network (config 1; raw 16-ch branch :634-635). */
int ra_render_anisdf_trace(ra_handle* h, const float* ray_o, const float* ray_d, const float* near, const float* far,
                           int64_t P, const ra_outputs* out, void* stream);
/* base_renderer.Renderer.render (config 2): 128 uniform samples per ray. */
int ra_render_anisdf_volume(ra_handle* h, const float* ray_o, const float* ray_d, const float* near, const float* far,
                            int64_t P, const ra_outputs* out, void* stream);

/* Building blocks exposed for parity tests (same kernels the render calls use). */
/* net.inference_world_distance_field(x, batch, smooth_transition, dist_th): x (n,3) -> sdf (n) */
int ra_query_sdf(ra_handle* h, const float* x, int64_t n, float dist_th, int32_t smooth, float* sdf, void* stream);
/* net(x, v, d, batch).raw (eval): x,v (n,3) -> raw (n, 17 | 16), zero rows out of shell */
int ra_query_raw(ra_handle* h, const float* x, const float* v, int64_t n, float* raw, void* stream);

/* exact K=3 nearest posed vertices (pytorch3d.ops.knn_points at sample_utils.py:122): x (n,3) world -> ids (n,3) vertex indices,
 * nearest first, d2 (n,3) squared distances in pose space */
int ra_query_knn(ra_handle* h, const float* x, int64_t n, int32_t* ids, float* d2, void* stream);
/* the same search as the shadow tracer runs it: every 32 consecutive points form a packet (parallel rays of neighbouring pixels)
 * whose far-field lanes walk the box hierarchy together.  Exact for any input; the grouping only decides which search runs. */
int ra_query_knn_packets(ra_handle* h, const float* x, int64_t n, int32_t* ids, float* d2, void* stream);

int ra_get_stats(ra_handle* h, ra_stats* out);   /* synchronises the device */
/* Device-side timing for bench.py's roofline line: when enabled, every launch of the fused MLP kernel and every
 * stage boundary of a render call is bracketed with CUDA events on the launching stream.  ra_profile_read
 * synchronises, returns the sums accumulated since the previous read and resets them.
 * stage_ms[4] = {surface tracing, surface attributes, light visibility, shading}. */
int ra_profile_enable(ra_handle* h, int32_t on);
int ra_profile_read(ra_handle* h, double* mlp_ms, int64_t* mlp_launches, double* stage_ms);
/* number of kernels the library launched since creation (bench.py's gpu_launches) */
int64_t ra_launch_count(ra_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* RA_B200_H */

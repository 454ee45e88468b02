"""ctypes binding of include/ra_b200.h.  Fails loudly when the CUDA library is missing or a call
returns an error -- there is no CPU fallback anywhere in the product path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RA_LIB_PATH') or os.path.join(HERE, 'libra_b200.so')      # RA_LIB_PATH: instrumented debug builds (tools/)

EXPORTS = ['ra_create', 'ra_destroy', 'ra_last_error', 'ra_upload_weights', 'ra_set_frame', 'ra_render_relight',
           'ra_relight_envmaps', 'ra_render_anisdf_trace', 'ra_render_anisdf_volume', 'ra_query_sdf', 'ra_query_raw',
           'ra_get_stats', 'ra_launch_count', 'ra_profile_enable', 'ra_profile_read', 'ra_rotate_probes', 'ra_assemble_image',
           'ra_ground_begin', 'ra_render_ground', 'ra_relight_ground', 'ra_blend_ground', 'ra_relight_envmaps_raw',
           'ra_upload_body', 'ra_prepare_pose', 'ra_prepare_rays', 'ra_set_ray_layout', 'ra_allgather', 'ra_visual_map', 'ra_assemble_visual',
           'ra_rotate_image', 'ra_set_main_light', 'ra_query_knn', 'ra_query_knn_packets']

fp = C.POINTER(C.c_float)


class ra_config(C.Structure):
    _fields_ = [('relight', C.c_int32), ('precision', C.c_int32), ('max_rays', C.c_int32), ('n_verts', C.c_int32),
                ('n_bones', C.c_int32), ('dist_th', C.c_float), ('blend_radius', C.c_float), ('resd_limit', C.c_float),
                ('st_iter', C.c_int32), ('st_tan_i', C.c_float), ('st_relax', C.c_float), ('st_offset', C.c_float),
                ('st_eps', C.c_float), ('st_skip', C.c_int32), ('lv_iter', C.c_int32), ('lv_offset', C.c_float),
                ('lv_relax', C.c_float), ('lv_near', C.c_float), ('lv_dist_th', C.c_float), ('env_r', C.c_float),
                ('bbox_margin', C.c_float), ('render_chunk', C.c_int32), ('n_samples', C.c_int32),
                ('surf_sample_range', C.c_float), ('fresnel_f0', C.c_float), ('albedo_slope', C.c_float),
                ('albedo_bias', C.c_float), ('rough_slope', C.c_float), ('rough_bias', C.c_float),
                ('albedo_multiplier', C.c_float), ('shading_albedo', C.c_float), ('env_h', C.c_int32), ('env_w', C.c_int32),
                ('vol_samples', C.c_int32), ('clip_near', C.c_float), ('clip_far', C.c_float), ('visibility_mode', C.c_int32), ('brdf_mode', C.c_int32),
                ('tonemapping', C.c_int32)]


class ra_weights(C.Structure):
    _fields_ = [('resd_w', fp * 9), ('resd_b', fp * 9), ('sdf_w', fp * 9), ('sdf_b', fp * 9), ('sdf_beta', C.c_float),
                ('render_w', fp * 5), ('render_b', fp * 5), ('albedo_w', fp * 3), ('albedo_b', fp * 3),
                ('rough_w', fp * 3), ('rough_b', fp * 3), ('env_main', fp), ('env_main_h', C.c_int32),
                ('env_main_w', C.c_int32), ('light_xyz', fp), ('light_area', fp), ('light_sharp', fp)]


class ra_frame(C.Structure):
    _fields_ = [(n, fp) for n in ('R', 'Th', 'poses', 'A', 'big_A', 'weights', 'pverts', 'pnorm', 'tverts', 'wbounds', 'mat_cond')]


OUTPUT_MAPS = ('rgb_map', 'acc_map', 'depth_map', 'surf_map', 'norm_map', 'cpts_map', 'bpts_map', 'resd_map', 'albedo_map',
               'roughness_map', 'shade_map', 'lvis_map', 'ldot_map', 'spec_map')


class ra_outputs(C.Structure):
    _fields_ = [(n, fp) for n in OUTPUT_MAPS]


class ra_body(C.Structure):
    _fields_ = [('tjoints', fp), ('parents', C.POINTER(C.c_int32)), ('rverts', fp), ('rnorm', fp), ('faces', C.POINTER(C.c_int32)),
                ('n_faces', C.c_int32), ('weights', fp)]


POSE_OUTPUTS = ('A', 'R', 'pverts', 'pnorm', 'wverts', 'wnorm', 'pbounds', 'wbounds')


class ra_pose_outputs(C.Structure):
    _fields_ = [(n, fp) for n in POSE_OUTPUTS]


GROUND_MAPS = ('rgb_map', 'surf_map', 'albedo_map', 'roughness_map', 'spec_map', 'norm_map', 'shade_map', 'depth_map', 'lvis_map', 'ldot_map')


class ra_ground_config(C.Structure):
    _fields_ = [('normal', C.c_float * 3), ('origin', C.c_float * 3), ('albedo', C.c_float * 3), ('attach_envmap', C.c_int32),
                ('shading_multiplier', C.c_float), ('iter', C.c_int32), ('offset', C.c_float), ('relax', C.c_float),
                ('near_offset', C.c_float), ('dist_th', C.c_float)]


class ra_ground_outputs(C.Structure):
    _fields_ = [(n, fp) for n in GROUND_MAPS]


VISUAL_INPUTS = ('rgb_map', 'acc_map', 'norm_map', 'depth_map', 'shade_map', 'albedo_map', 'roughness_map', 'cpts_map', 'bpts_map', 'surf_map',
                 'spec_map', 'cam_R', 'tbounds')
VIS_TYPES = {'rendering': 0, 'normal': 1, 'alpha': 2, 'depth': 3, 'shading': 4, 'albedo': 5, 'roughness': 6, 'surface': 7, 'residual': 8, 'specular': 9}


class ra_visual_inputs(C.Structure):
    _fields_ = [(n, fp) for n in VISUAL_INPUTS]


class ra_visual_config(C.Structure):
    _fields_ = [('min_clip', C.c_float), ('normalize_shading', C.c_int32), ('normalize_specular', C.c_int32), ('tonemapping_albedo', C.c_int32)]


class ra_image_config(C.Structure):
    _fields_ = [('bg_brightness', C.c_float), ('channels', C.c_int32), ('bgr', C.c_int32), ('probe', fp), ('eh', C.c_int32), ('ew', C.c_int32),
                ('probe_dirs', fp), ('uH', C.c_int32), ('uW', C.c_int32)]


class ra_stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ('n_rays', 'n_fg', 'n_shadow_rays', 'n_queries', 'n_queries_in_shell', 'n_attr_samples', 'n_dropped_shadow_rays', 'n_shadow_slots')]


_lib = None


def load():
    """Load libra_b200.so; raises (never falls back) if it is absent or lacks a declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} not built: run `python -m relightableavatar_b200.build` (no CPU fallback exists)')
    lib = C.CDLL(LIB_PATH)
    for s in EXPORTS:
        if not hasattr(lib, s):
            raise RuntimeError(f'{LIB_PATH} does not export {s}')
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
    lib.ra_create.argtypes = [C.POINTER(vp), C.POINTER(ra_config)]
    lib.ra_destroy.argtypes = [vp]; lib.ra_destroy.restype = None
    lib.ra_last_error.argtypes = [vp]; lib.ra_last_error.restype = C.c_char_p
    lib.ra_upload_weights.argtypes = [vp, C.POINTER(ra_weights), vp]
    lib.ra_set_frame.argtypes = [vp, C.POINTER(ra_frame), vp]
    for n in ('ra_render_relight', 'ra_render_anisdf_trace', 'ra_render_anisdf_volume'):
        getattr(lib, n).argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(ra_outputs), vp]
    lib.ra_relight_envmaps.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    lib.ra_query_sdf.argtypes = [vp, vp, i64, f32, i32, vp, vp]
    lib.ra_query_raw.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.ra_get_stats.argtypes = [vp, C.POINTER(ra_stats)]
    lib.ra_launch_count.argtypes = [vp]; lib.ra_launch_count.restype = i64
    lib.ra_assemble_image.argtypes = [vp, vp, vp, vp, i32, i32, f32, vp, vp, vp]
    lib.ra_rotate_probes.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.ra_relight_envmaps_raw.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    lib.ra_ground_begin.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.ra_render_ground.argtypes = [vp, C.POINTER(ra_ground_config), vp, vp, vp, i64, vp, i32, i32, vp, i32, i32, C.POINTER(ra_ground_outputs), vp]
    lib.ra_relight_ground.argtypes = [vp, C.POINTER(ra_ground_config), vp, i32, i32, vp, i32, i32, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.ra_blend_ground.argtypes = [vp, vp, vp, vp, i32, i32, i64, vp, vp]
    lib.ra_set_ray_layout.argtypes = [vp, i64, i32, i32, i32]
    lib.ra_allgather.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.ra_upload_body.argtypes = [vp, C.POINTER(ra_body), vp]
    lib.ra_prepare_pose.argtypes = [vp, vp, vp, vp, f32, C.POINTER(ra_pose_outputs), vp]
    lib.ra_prepare_rays.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ra_visual_map.argtypes = [vp, i32, C.POINTER(ra_visual_inputs), i64, C.POINTER(ra_visual_config), vp, vp]
    lib.ra_assemble_visual.argtypes = [vp, vp, vp, vp, i32, i32, C.POINTER(ra_image_config), vp, vp, vp, vp]
    lib.ra_rotate_image.argtypes = [vp, vp, i32, i32, C.c_double, i32, i32, vp, vp]
    lib.ra_set_main_light.argtypes = [vp, vp, i32, i32, vp]
    lib.ra_query_knn.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.ra_query_knn_packets.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.ra_profile_enable.argtypes = [vp, i32]
    lib.ra_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double)]
    _lib = lib
    return lib

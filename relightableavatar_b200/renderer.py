"""Host-side mirror of the reference renderer plugin surface.

The reference builds its renderer with
    importlib.import_module(cfg.renderer_module).Renderer(network)     lib/networks/renderer/make_renderer.py:5-8
and calls `renderer.render(batch) -> dotdict` under no_grad with `net.eval()` (run.py:68-85).  This
module offers the same surface -- `Renderer(net).render(batch)` -- and forwards the whole per-pixel
path to the CUDA library through the C-ABI of include/ra_b200.h:

    relightableavatar_b200.renderer            <->  reference module
    Renderer(net, mode='relight')               lib.networks.renderer.novel_light_sphere_tracing.Renderer
    Renderer(net, mode='anisdf_trace')          lib.networks.renderer.sphere_tracing_renderer.Renderer (AniSDF net)
    Renderer(net, mode='anisdf_volume')         lib.networks.renderer.base_renderer.Renderer

PyTorch is used for device memory, streams and (in parallel.py) torch.distributed only.  There is no
CPU fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import time
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import GROUND_MAPS, OUTPUT_MAPS, ra_config, ra_frame, ra_ground_config, ra_ground_outputs, ra_outputs, ra_stats, ra_weights

PRECISION = {'fp32': 0, 'tc': 1}


class dotdict(dict):
    """Attribute-access dict, like the reference's lib.utils.base_utils.dotdict."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


def default_config(relight: bool = True, **over) -> Dict:
    """Effective cfg values of xuzhen_12v_geo(_fix_mat) (notation paragraph)."""
    c = dict(relight=int(relight), precision=1, max_rays=1 << 17, n_verts=6890, n_bones=52,
             dist_th=0.125 if relight else 0.1, blend_radius=0.075, resd_limit=0.05,
             st_iter=16, st_tan_i=1000.0, st_relax=0.0, st_offset=0.02, st_eps=1e-8, st_skip=1,
             lv_iter=4, lv_offset=0.01, lv_relax=0.0, lv_near=0.02, lv_dist_th=0.125 if relight else 0.05,      # (unused without relighting)
             env_r=10.0, bbox_margin=0.25, render_chunk=65536, n_samples=3, surf_sample_range=0.005,
             fresnel_f0=0.02, albedo_slope=1.0, albedo_bias=0.0, rough_slope=0.9, rough_bias=0.09,
             albedo_multiplier=1.0, shading_albedo=0.8, env_h=16, env_w=32, vol_samples=128, clip_near=0.02, clip_far=10.0,
             visibility_mode=0, brdf_mode=0, tonemapping=1)
    c.update(over)
    return c


# Switches the reference reads on this path (ablations, debug views, alternative encodings) that the library implements at the
# reference's DEFAULT value only.  A cfg that sets one of them differently is refused instead of rendered differently.
# (Defaults: lib/config/config.py; pinned by tests/golden/cfg_values.json.)
FIXED_SWITCHES = dict(
    smpl_distance=False, ablate_hdq_mode='hdq', use_geodesic_filter=True, sample_vert_cnt=3, sdf_finite_diff=0,      # HDQ (base_network.py)
    xyz_res=10, view_res=4, sdf_res=8, feat_dim=256, relight_network_width=128, relight_network_depth=2, lambertian=False,   # network shapes
    no_dfss=False, no_claybook=False, only_visibility=False,            # shadow tracing (no_visibility / local_visibility: ra_config.visibility_mode)
    geometry_visibility=False, geometry_normal=False, bruteforce_st=False, check_termination_sdf=False, check_bound_sdf=False,
    zero_roughness=False, rgb_as_albedo=False,              # shading (lambert_only / glossy_only: ra_config.brdf_mode; replace_light: ra_set_main_light)
    vis_lvis_map=False, vis_ldot_map=False,             # debug views that overwrite shade_map with a visibility / cosine mean (:756-757)
    bg_brightness=0.0)
FIXED_ST_SWITCHES = dict(tan_i_multiplier=1)          # cfg.sphere_tracing.*


def check_fixed_switches(cfg) -> None:
    """Raise NotImplementedError when the reference cfg asks for a non-default value of a switch in FIXED_SWITCHES."""
    def get(node, k):
        try:
            return node[k] if k in node else None
        except TypeError:
            return getattr(node, k, None)
    bad = {k: get(cfg, k) for k, d in FIXED_SWITCHES.items() if get(cfg, k) is not None and get(cfg, k) != d}
    st = get(cfg, 'sphere_tracing')
    if st is not None:
        bad.update({'sphere_tracing.' + k: get(st, k) for k, d in FIXED_ST_SWITCHES.items() if get(st, k) is not None and get(st, k) != d})
    if bad:
        raise NotImplementedError('relightableavatar_b200 implements these reference switches at their default value only: '
                                  + ', '.join(f'{k}={v!r} (default {FIXED_SWITCHES.get(k, FIXED_ST_SWITCHES.get(k.split(".")[-1]))!r})' for k, v in bad.items()))


def config_from_reference_cfg(cfg, relight: bool, mode: Optional[str] = None) -> Dict:
    """Read the same keys the reference renderer reads from its global `cfg` (once, at construction).
    `cfg.n_samples` is the number of surface samples in the traced modes (3) and of ray samples in the volume renderer (128,
    base_renderer.py): it lands in `n_samples` or `vol_samples` accordingly."""
    check_fixed_switches(cfg)
    st, lv = cfg.sphere_tracing, cfg.obj_lvis
    volume = mode == 'anisdf_volume'
    return default_config(
        relight, dist_th=cfg.dist_th, blend_radius=cfg.blend_radius, resd_limit=cfg.resd_limit,
        st_iter=st.iter, st_tan_i=st.tan_i, st_relax=st.relax, st_offset=st.offset, st_eps=st.eps, st_skip=st.shadow_skip_iter,
        lv_iter=lv.iter, lv_offset=lv.offset, lv_relax=lv.relax, lv_near=lv.near_offset, lv_dist_th=lv.dist_th,
        env_r=cfg.env_r, bbox_margin=cfg.env_lvis.bbox_margin, render_chunk=cfg.render_chunk_size,
        **({'vol_samples': cfg.n_samples} if volume else {'n_samples': cfg.n_samples}),
        surf_sample_range=cfg.surf_sample_range, fresnel_f0=cfg.fresnel_f0, albedo_slope=cfg.albedo_slope,
        albedo_bias=cfg.albedo_bias, rough_slope=cfg.roughness_slope, rough_bias=cfg.roughness_bias,
        albedo_multiplier=cfg.albedo_multiplier, shading_albedo=cfg.shading_albedo, env_h=cfg.env_h, env_w=cfg.env_w,
        clip_near=cfg.clip_near, clip_far=cfg.clip_far, tonemapping=int(bool(cfg.tonemapping_rendering)),
        # ablation switches of light_visibility (sphere_tracing_renderer.py:296-301) and Microfacet (relight_utils.py:563-568)
        visibility_mode=2 if _flag(cfg, 'no_visibility') else (1 if _flag(cfg, 'local_visibility') else 0),
        brdf_mode=1 if _flag(cfg, 'lambert_only') else (2 if _flag(cfg, 'glossy_only') else 0))


def _flag(cfg, name) -> bool:
    try:
        return bool(cfg[name]) if name in cfg else False
    except TypeError:
        return bool(getattr(cfg, name, False))


def default_ground_config(**over) -> Dict:
    """cfg.ground_* (config.py:45,104-107,353) and cfg.env_lvis (config.py:135-141)."""
    g = dict(normal=(0.0, 0.0, 1.0), origin=(0.0, 0.0, 0.0), albedo=(0.05, 0.05, 0.05), attach_envmap=1, shading_multiplier=1.0,
             iter=16, offset=0.01, relax=0.0, near_offset=0.02, dist_th=0.005)
    g.update(over)
    return g


def ground_config_from_reference_cfg(cfg) -> Dict:
    e = cfg.env_lvis
    return default_ground_config(normal=tuple(cfg.ground_normal), origin=tuple(cfg.ground_origin), albedo=tuple(cfg.ground_albedo),
                                 attach_envmap=int(cfg.ground_attach_envmap), shading_multiplier=cfg.ground_shading_multiplier,
                                 iter=e.iter, offset=e.offset, relax=e.relax, near_offset=e.near_offset, dist_th=e.dist_th)


def get_rays(H: int, W: int, K: torch.Tensor, R: torch.Tensor, T: torch.Tensor):
    """Every pixel's ray, like net_utils.get_rays (:403-420): K, R (3,3), T (3,1) -> ray_o, ray_d (H*W, 3)."""
    K, R, T = K.reshape(3, 3), R.reshape(3, 3), T.reshape(3, 1)
    ray_o = -(R.mT @ T).ravel()
    i, j = torch.meshgrid(torch.arange(H, dtype=R.dtype, device=R.device), torch.arange(W, dtype=R.dtype, device=R.device), indexing='ij')
    xy1 = torch.stack([j, i, torch.ones_like(i)], dim=2)
    pixel_world = (xy1 @ torch.inverse(K).mT - T.ravel()) @ R
    d = pixel_world - ray_o
    ray_d = d / (d.norm(dim=-1, keepdim=True) + 1e-8)
    return ray_o[None, None].expand(pixel_world.shape).reshape(-1, 3).contiguous(), ray_d.reshape(-1, 3).contiguous()


def material_condition(batch: Dict, fix_material: int = 0, always_fix_material: bool = True) -> Optional[torch.Tensor]:
    """The colour network's condition vector (base_network.py:501-503): `train_motion.poses[:, fix_material]` (python indexing, so -1
    is the LAST training pose) when `fix_material >= 0 or always_fix_material`, otherwise this frame's own poses.  None when the
    batch carries no training motion (fine for relighting, which never evaluates the colour network)."""
    if fix_material >= 0 or always_fix_material:
        tm = batch.get('train_motion') if hasattr(batch, 'get') else None
        poses = tm['poses'] if tm is not None else batch.get('train_poses')
        return None if poses is None else torch.as_tensor(poses)[0, int(fix_material)].reshape(-1)
    return torch.as_tensor(batch['poses']).reshape(-1)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _fptr(t: Optional[torch.Tensor]):
    return C.cast(C.c_void_p(t.data_ptr() if t is not None else 0), _lib.fp)


class Engine:
    """Thin RAII wrapper over one `ra_handle` (one per device)."""

    def __init__(self, config: Dict, device='cuda:0'):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('relightableavatar_b200 runs on CUDA (sm_100a) only; there is no CPU path')
        self.config = dict(config)
        cfg = ra_config(**config)
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.ra_create(C.byref(self.h), C.byref(cfg))
        if rc:
            msg = self.lib.ra_last_error(self.h).decode() if self.h else 'ra_create failed'
            if self.h:
                self.lib.ra_destroy(self.h)          # releases the partially built handle and every device block it allocated
                self.h = None
            raise RuntimeError(f'ra_create: {msg}')
        self._keep = []          # tensors borrowed by the library until the next set_frame
        self._wkeep = []
        self.L = config['env_h'] * config['env_w']

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f'{what}: {self.lib.ra_last_error(self.h).decode()}')

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, 'h', None):
            torch.cuda.synchronize(self.device)
            self.lib.ra_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def upload_weights(self, sd: Dict[str, torch.Tensor]):
        """Extract the tensors named in SURVEY.md 8b from a state-dict; fold weight-norm (w = g v/||v||)."""
        dev = self.device
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        keep = []
        w = ra_weights()
        cond = 3 * self.config['n_bones']
        k0 = sd['residual_deformation_network.mlp.linears.0.weight'].shape[1]
        if k0 != 63 + cond:          # the C side reads (256, 63 + 3 n_bones) rows: a mismatch would silently shear every row
            raise ValueError(f'residual MLP input width {k0} != 63 + 3 * n_bones ({63 + cond}): create the Engine with n_bones={(k0 - 63) // 3}')

        def put(arr, i, t):
            t = f(t); keep.append(t); arr[i] = _fptr(t)

        for l in range(9):
            put(w.resd_w, l, sd[f'residual_deformation_network.mlp.linears.{l}.weight'])
            put(w.resd_b, l, sd[f'residual_deformation_network.mlp.linears.{l}.bias'])
            v, g = f(sd[f'signed_distance_network.mlp.lin{l}.weight_v']), f(sd[f'signed_distance_network.mlp.lin{l}.weight_g'])
            put(w.sdf_w, l, g * v / v.norm(dim=1, keepdim=True))
            put(w.sdf_b, l, sd[f'signed_distance_network.mlp.lin{l}.bias'])
        w.sdf_beta = float(sd['signed_distance_network._beta'].clamp(1e-9, 1e6))
        if 'render_network.l0.weight_v' in sd:
            for l in range(5):
                v, g = f(sd[f'render_network.l{l}.weight_v']), f(sd[f'render_network.l{l}.weight_g'])
                put(w.render_w, l, g * v / v.norm(dim=1, keepdim=True))
                put(w.render_b, l, sd[f'render_network.l{l}.bias'])
        if self.config['relight']:
            for l in range(3):
                put(w.albedo_w, l, sd[f'albedo_network.linears.{l}.weight']); put(w.albedo_b, l, sd[f'albedo_network.linears.{l}.bias'])
                put(w.rough_w, l, sd[f'roughness_network.linears.{l}.weight']); put(w.rough_b, l, sd[f'roughness_network.linears.{l}.bias'])
            env = f(sd['global_env_map_'])
            env = F.softplus(env.expand(*env.shape[:2], 3)).contiguous()        # relight_network.py:86-89
            keep.append(env)
            w.env_main, w.env_main_h, w.env_main_w = _fptr(env), env.shape[0], env.shape[1]
            self.env_main = env
            for name, key in (('light_xyz', 'light_xyz_'), ('light_area', 'light_area'), ('light_sharp', 'light_sharp')):
                t = f(sd[key]); keep.append(t); setattr(w, name, _fptr(t))
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_upload_weights(self.h, C.byref(w), self._stream()), 'ra_upload_weights')
        self._wkeep = keep

    # ------------------------------------------------------------------ frame
    def set_frame(self, batch: Dict, fix_material: int = 0, always_fix_material: bool = True):
        """Per-frame state.  The colour network's condition follows base_network.py:501-503: train_motion.poses[fix_material]
        (python indexing: -1 is the LAST training pose) when `fix_material >= 0 or always_fix_material`, else this frame's poses."""
        dev = self.device

        def g(key, shape=None):
            t = batch[key]
            t = torch.as_tensor(t)
            if t.device != dev or t.dtype != torch.float32:
                t = t.to(device=dev, dtype=torch.float32)
            t = t[0] if (t.ndim > 1 and t.shape[0] == 1) else t
            return t.contiguous()

        fr = ra_frame()
        keep = []
        J, N = self.config['n_bones'], self.config['n_verts']
        want = {'poses': 3 * J, 'A': J * 16, 'big_A': J * 16, 'weights': N * J, 'pverts': N * 3, 'pnorm': N * 3, 'tverts': N * 3,
                'R': 9, 'Th': 3, 'wbounds': 6}
        for name in ('R', 'Th', 'poses', 'A', 'big_A', 'weights', 'pverts', 'pnorm', 'tverts', 'wbounds'):
            t = g(name)
            if t.numel() != want[name]:
                raise ValueError(f'batch.{name} has {t.numel()} elements, expected {want[name]} (n_verts={N}, n_bones={J})')
            keep.append(t); setattr(fr, name, _fptr(t))
        mc = material_condition(batch, fix_material, always_fix_material)
        if mc is not None:
            mc = mc.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if mc is None and not self.config['relight']:
            raise KeyError('batch.train_motion.poses (the colour network\'s material condition) is missing')
        if mc is not None:
            if mc.numel() != 3 * J:
                raise ValueError(f'train_motion.poses rows have {mc.numel()} elements, expected {3 * J}')
            keep.append(mc)
        fr.mat_cond = _fptr(mc)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_set_frame(self.h, C.byref(fr), self._stream()), 'ra_set_frame')
        self._keep = keep

    # ------------------------------------------------------------------ render
    def _rays(self, batch):
        dev = self.device
        c = lambda k: torch.as_tensor(batch[k]).to(device=dev, dtype=torch.float32)[0].contiguous()
        return c('ray_o'), c('ray_d'), c('near'), c('far')

    def _alloc_outputs(self, P, keys):
        out = {}
        for k in keys:
            if k in ('acc_map', 'depth_map', 'roughness_map'):
                out[k] = torch.empty(P, device=self.device, dtype=torch.float32)
            elif k in ('lvis_map', 'ldot_map'):
                out[k] = torch.empty(P, self.L, device=self.device, dtype=torch.float32)
            else:
                out[k] = torch.empty(P, 3, device=self.device, dtype=torch.float32)
        o = ra_outputs()
        for k in OUTPUT_MAPS:
            setattr(o, k, _fptr(out.get(k)))
        return out, o

    def render(self, mode: str, ray_o, ray_d, near, far, keys):
        P = ray_o.shape[0]
        out, o = self._alloc_outputs(P, keys)
        fn = {'relight': self.lib.ra_render_relight, 'anisdf_trace': self.lib.ra_render_anisdf_trace,
              'anisdf_volume': self.lib.ra_render_anisdf_volume}[mode]
        self._rays_keep = (ray_o, ray_d, near, far)
        with torch.cuda.device(self.device):
            self._check(fn(self.h, _ptr(ray_o), _ptr(ray_d), _ptr(near), _ptr(far), P, C.byref(o), self._stream()), mode)
        return out

    def set_main_light(self, probe: Optional[torch.Tensor]):
        """cfg.replace_light (sphere_tracing_renderer.py:1068-1069): env-map (ph,pw,3) that lights the main pass instead of the learned
        one; None restores the learned light."""
        if probe is None:
            self._check(self.lib.ra_set_main_light(self.h, _ptr(None), 0, 0, self._stream()), 'ra_set_main_light')
            return
        probe = probe.to(device=self.device, dtype=torch.float32)
        probe = (probe[0] if probe.ndim == 4 else probe).contiguous()
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_set_main_light(self.h, _ptr(probe), probe.shape[0], probe.shape[1], self._stream()), 'ra_set_main_light')

    def set_ray_layout(self, global_P: int, block: int = 32, world: int = 1, rank: int = 0):
        """Tile sharding: tell the library which rays of the whole frame this handle renders (see ra_set_ray_layout)."""
        self._check(self.lib.ra_set_ray_layout(self.h, int(global_P), int(block), int(world), int(rank)), 'ra_set_ray_layout')

    def relight_envmaps(self, probes: torch.Tensor, P: int, want_spec=True):
        n = probes.shape[0]
        rgb = torch.empty(n, P, 3, device=self.device); shade = torch.empty_like(rgb)
        spec = torch.empty_like(rgb) if want_spec else None
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_relight_envmaps(self.h, _ptr(probes), n, _ptr(rgb), _ptr(shade), _ptr(spec), self._stream()),
                        'ra_relight_envmaps')
        return rgb, shade, spec

    # ------------------------------------------------------------------ ground plane
    def _gcfg(self, g: Dict) -> ra_ground_config:
        c = ra_ground_config()
        for k, v in g.items():
            if k in ('normal', 'origin', 'albedo'):
                setattr(c, k, (C.c_float * 3)(*[float(x) for x in v]))
            else:
                setattr(c, k, v)
        return c

    def relight_envmaps_raw(self, probes: torch.Tensor, P: int):
        """Per-env-map human re-shade from the RAW maps (ground shading on), outputs multiplied by acc."""
        n = probes.shape[0]
        rgb = torch.empty(n, P, 3, device=self.device); shade = torch.empty_like(rgb); spec = torch.empty_like(rgb)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_relight_envmaps_raw(self.h, _ptr(probes), n, _ptr(rgb), _ptr(shade), _ptr(spec), self._stream()),
                        'ra_relight_envmaps_raw')
        return rgb, shade, spec

    def ground_begin(self, mask_at_box: torch.Tensor, acc_map: torch.Tensor) -> torch.Tensor:
        mask = torch.as_tensor(mask_at_box).to(self.device).reshape(mask_at_box.shape[-2], mask_at_box.shape[-1]).to(torch.uint8).contiguous()
        H, W = mask.shape
        acc_g = torch.empty(H * W, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_ground_begin(self.h, _ptr(mask), H, W, _ptr(acc_map.contiguous()), _ptr(acc_g), self._stream()), 'ra_ground_begin')
        return acc_g

    def render_ground(self, g: Dict, ray_o: torch.Tensor, ray_d: torch.Tensor, acc_g: torch.Tensor, probe: torch.Tensor,
                      image: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        Fp = ray_o.shape[0]
        out = {}
        for k in GROUND_MAPS:
            shape = (Fp,) if k in ('roughness_map', 'depth_map') else ((Fp, self.L) if k in ('lvis_map', 'ldot_map') else (Fp, 3))
            out[k] = torch.empty(*shape, device=self.device)
        o = ra_ground_outputs()
        for k in GROUND_MAPS:
            setattr(o, k, _fptr(out[k]))
        gc = self._gcfg(g)
        probe = probe.contiguous()
        ih, iw = (image.shape[0], image.shape[1]) if image is not None else (0, 0)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_render_ground(self.h, C.byref(gc), _ptr(ray_o), _ptr(ray_d), _ptr(acc_g), Fp, _ptr(probe), probe.shape[0], probe.shape[1],
                                                  _ptr(image), ih, iw, C.byref(o), self._stream()), 'ra_render_ground')
        return out

    def relight_ground(self, g: Dict, probe: torch.Tensor, ray_d: torch.Tensor, ground: Dict[str, torch.Tensor], image: Optional[torch.Tensor] = None):
        Fp = ray_d.shape[0]
        rgb = torch.empty(Fp, 3, device=self.device); albedo = torch.empty_like(rgb); shade = torch.empty_like(rgb); spec = torch.empty_like(rgb)
        gc = self._gcfg(g)
        probe = probe.contiguous()
        ih, iw = (image.shape[0], image.shape[1]) if image is not None else (0, 0)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_relight_ground(self.h, C.byref(gc), _ptr(probe), probe.shape[0], probe.shape[1], _ptr(image), ih, iw, _ptr(ray_d),
                                                   _ptr(ground['albedo_map']), _ptr(ground['lvis_map']), _ptr(ground['ldot_map']), Fp,
                                                   _ptr(rgb), _ptr(albedo), _ptr(shade), _ptr(spec), self._stream()), 'ra_relight_ground')
        return rgb, albedo, shade, spec

    def blend_ground(self, acc_g: torch.Tensor, ground: Optional[torch.Tensor], human: Optional[torch.Tensor], human_premul: bool = True) -> torch.Tensor:
        """blend_output_ for one key: (F,C) or (F,) ground / (P,C) or (P,) human -> image-sized map."""
        ref = ground if ground is not None else human
        Cn = 1 if ref.ndim == 1 else ref.shape[-1]
        Fp = acc_g.shape[0]
        out = torch.empty((Fp,) if ref.ndim == 1 else (Fp, Cn), device=self.device)
        gt = ground.contiguous() if ground is not None else None
        ht = human.contiguous() if human is not None else None
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_blend_ground(self.h, _ptr(acc_g), _ptr(gt), _ptr(ht), int(human_premul), Cn, Fp, _ptr(out), self._stream()), 'ra_blend_ground')
        return out

    def rotate_probes(self, probe: torch.Tensor, repeat: int, j0: int, n_rot: int) -> torch.Tensor:
        """rotate_envmap / shift_image (relight_utils.py:55-103): (eh, ew, 3) -> (n_rot, eh, ew, 3)."""
        probe = probe.to(device=self.device, dtype=torch.float32).reshape(self.config['env_h'], self.config['env_w'], 3).contiguous()
        out = torch.empty(n_rot, self.config['env_h'], self.config['env_w'], 3, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_rotate_probes(self.h, _ptr(probe), int(repeat), int(j0), int(n_rot), _ptr(out), self._stream()), 'ra_rotate_probes')
        return out

    def rotate_image(self, image: torch.Tensor, repeat: int, j0: int, n_rot: int) -> torch.Tensor:
        """rotate_envmap's shift_image on the env-map IMAGE attached to the floor (relight_utils.py:74-75,103): (iH,iW,3) ->
        (n_rot,iH,iW,3), rotation j shifts by iW / (env_w * repeat) * j texels."""
        image = image.to(device=self.device, dtype=torch.float32).contiguous()
        iH, iW = int(image.shape[0]), int(image.shape[1])
        out = torch.empty(n_rot, iH, iW, 3, device=self.device)
        step = iW / (self.config['env_w'] * repeat)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_rotate_image(self.h, _ptr(image), iH, iW, C.c_double(step), int(j0), int(n_rot), _ptr(out), self._stream()), 'ra_rotate_image')
        return out

    def assemble_image(self, rgb_map: torch.Tensor, acc_map: torch.Tensor, mask_at_box: torch.Tensor, bg_brightness: float = 0.0):
        """Visualizer.generate_image's scatter (base_visualizer.py:182-202): -> (H,W,4) float RGBA and (H,W,4) uint8."""
        mask = torch.as_tensor(mask_at_box).to(self.device).reshape(mask_at_box.shape[-2], mask_at_box.shape[-1]).to(torch.uint8).contiguous()
        H, W = mask.shape
        rgb = rgb_map.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous()
        acc = acc_map.to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        out_f = torch.empty(H, W, 4, device=self.device)
        out_u8 = torch.empty(H, W, 4, device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_assemble_image(self.h, _ptr(rgb), _ptr(acc), _ptr(mask), H, W, float(bg_brightness), _ptr(out_f), _ptr(out_u8),
                                                   self._stream()), 'ra_assemble_image')
        return out_f, out_u8

    def query_sdf(self, x: torch.Tensor, dist_th: Optional[float] = None, smooth: bool = True) -> torch.Tensor:
        x = x.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous()
        out = torch.empty(x.shape[0], device=self.device)
        th = float(dist_th if dist_th is not None else self.config['dist_th'])
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_query_sdf(self.h, _ptr(x), x.shape[0], th, int(smooth), _ptr(out), self._stream()), 'ra_query_sdf')
        return out

    def query_knn(self, x: torch.Tensor, packets: bool = False):
        """Exact 3 nearest posed vertices of world points (sample_utils.py:122): -> ids (n,3) int32 vertex indices, d2 (n,3).
        `packets`: run the search the way the shadow tracer does (every 32 consecutive points = one packet of nearby points)."""
        x = x.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous()
        ids = torch.empty(x.shape[0], 3, device=self.device, dtype=torch.int32)
        d2 = torch.empty(x.shape[0], 3, device=self.device)
        with torch.cuda.device(self.device):
            fn = self.lib.ra_query_knn_packets if packets else self.lib.ra_query_knn
            self._check(fn(self.h, _ptr(x), x.shape[0], _ptr(ids), _ptr(d2), self._stream()), 'ra_query_knn')
        return ids, d2

    def query_raw(self, x: torch.Tensor, v: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = x.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous()
        v = v.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous() if v is not None else None
        Cn = 17 if self.config['relight'] else 16
        out = torch.empty(x.shape[0], Cn, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ra_query_raw(self.h, _ptr(x), _ptr(v), x.shape[0], _ptr(out), self._stream()), 'ra_query_raw')
        return out

    def stats(self) -> Dict[str, int]:
        s = ra_stats()
        self._check(self.lib.ra_get_stats(self.h, C.byref(s)), 'ra_get_stats')
        return {n: getattr(s, n) for n, _ in s._fields_}

    def launch_count(self) -> int:
        return int(self.lib.ra_launch_count(self.h))

    def profile_enable(self, on: bool = True):
        self._check(self.lib.ra_profile_enable(self.h, int(on)), 'ra_profile_enable')

    def profile_read(self):
        ms, n = C.c_double(0), C.c_int64(0)
        st = (C.c_double * 4)()
        self._check(self.lib.ra_profile_read(self.h, C.byref(ms), C.byref(n), st), 'ra_profile_read')
        return dict(mlp_ms=ms.value, mlp_launches=n.value, stage_ms=dict(zip(('surface', 'attributes', 'visibility', 'shading'), list(st))))


_MAIN_KEYS = {
    'relight': ('rgb_map', 'acc_map', 'depth_map', 'surf_map', 'norm_map', 'cpts_map', 'bpts_map', 'albedo_map', 'roughness_map', 'shade_map'),
    'anisdf_trace': ('rgb_map', 'acc_map', 'depth_map', 'surf_map', 'norm_map', 'cpts_map', 'bpts_map', 'resd_map'),
    'anisdf_volume': ('rgb_map', 'acc_map', 'depth_map', 'norm_map', 'cpts_map', 'bpts_map', 'resd_map'),
}


class Renderer(torch.nn.Module):
    """Drop-in for the reference's `Renderer(net)`; `render(batch)` returns the same dotdict layout.

    `net` is the loaded reference network (any object with `state_dict()`); its tensors are read once.
    `mode` selects which reference renderer is mirrored (see module docstring).  With a reference `cfg`
    object pass `cfg=` to read the same keys the reference reads; otherwise the xuzhen_12v_geo values apply.
    """

    def __init__(self, net, mode: str = 'relight', cfg=None, device='cuda:0', precision: str = 'tc', max_rays: int = 1 << 17,
                 test_light=None, return_lvis: bool = False, to_cpu: bool = False, sync_timing: bool = True,
                 ground_shading: Optional[bool] = None, ground: Optional[Dict] = None, rotate_ratio: Optional[int] = None,
                 engine=None, **overrides):
        """`test_light` (default: cfg.test_light under a reference cfg, else ('main',)): 'main' in it keeps the learned-light
        rendering (novel_light_sphere_tracing.py:155).  With a
        reference `cfg` every env-map of `batch.novel_lights` is rendered, as the reference does (its dataset already filtered them
        by cfg.test_light, base_dataset.py:141,159); without one, the env-maps named in `test_light` (or all of them for 'all').
        `rotate_ratio` (cfg.rotate_ratio when cfg.vis_rotate_light): every env-map is rendered at rotate_ratio * env_w rotations.
        `engine`: an already constructed Engine-like object (tests); otherwise one is created on `device`."""
        super().__init__()
        self.net = net
        self.mode = mode
        relight = mode == 'relight'
        conf = config_from_reference_cfg(cfg, relight, mode) if cfg is not None else default_config(relight)
        # cfg.n_bones / cfg.cond_dim are derived from the body model by the reference (config.py:441,465-466); the network's own
        # first layer is the authority here: 63 + 3 * n_bones input columns (52 SMPL-H, 24 SMPL)
        k0 = net.state_dict()['residual_deformation_network.mlp.linears.0.weight'].shape[1]
        conf.update(n_bones=(k0 - 63) // 3)
        conf.update(precision=PRECISION[precision], max_rays=int(max_rays))
        # cfg.vis_specular_map: the main pass also returns its specular view (sphere_tracing_renderer.py:739-748)
        self.main_spec = bool(overrides.pop('vis_specular_map', getattr(cfg, 'vis_specular_map', False) if cfg is not None else False))
        # cfg.replace_light: name of the batch env-map that lights the main pass instead of the learned one (:1068-1069)
        self.replace_light = overrides.pop('replace_light', None) or (getattr(cfg, 'replace_light', '') if cfg is not None else '') or ''
        conf.update(overrides)
        self.engine = Engine(conf, device) if engine is None else engine
        if engine is None:
            self.engine.upload_weights(net.state_dict())
        self.all_lights = cfg is not None
        if rotate_ratio is None:
            rotate_ratio = int(cfg.rotate_ratio) if (cfg is not None and getattr(cfg, 'vis_rotate_light', False)) else -1
        self.rotate_ratio = int(rotate_ratio)
        self.fix_material = getattr(cfg, 'fix_material', 0) if cfg is not None else 0
        afm = getattr(cfg, 'always_fix_material', None) if cfg is not None else None
        self.always_fix_material = True if afm is None else bool(afm)
        if test_light is None:
            test_light = tuple(getattr(cfg, 'test_light', None) or ()) if cfg is not None else ('main',)
        self.test_light = tuple(test_light)
        self.return_lvis = return_lvis
        self.to_cpu = to_cpu
        self.sync_timing = sync_timing     # reference behaviour: cuda.synchronize + perf_counter around the main pass
        # cfg.vis_ground_shading (readme.md:64): floor pass over all H*W pixels + blend_output_; outputs become image-sized
        self.ground_shading = bool(getattr(cfg, 'vis_ground_shading', False)) if ground_shading is None else bool(ground_shading)
        self.ground = ground_config_from_reference_cfg(cfg) if (cfg is not None and ground is None) else default_ground_config(**(ground or {}))

    def _light_names(self, lights) -> list:
        return [n for n in lights if self.all_lights or n in self.test_light or 'all' in self.test_light]

    def _probe_of(self, light, key='probe'):
        return light.get(key) if isinstance(light, dict) else (light if key == 'probe' else None)

    def _light_sweep(self, lights, names, with_image: bool = False):
        """rotate_envmap's enumeration (relight_utils.py:55-103): yields (name, probe (eh,ew,3) on the device) in the reference's
        order -- every env-map once, or at rotate_ratio * env_w rotations named f'{key}-{j:04d}' (cfg.vis_rotate_light).
        `with_image`: yields (name, probe, image) -- the env-map image attached to the floor rotates with the probe (:74-75,103);
        None when the light carries no image (the floor then samples the probe)."""
        eng = self.engine
        eh, ew = eng.config['env_h'], eng.config['env_w']
        for n in names:
            probe = torch.as_tensor(self._probe_of(lights[n])).to(device=eng.device, dtype=torch.float32).reshape(eh, ew, 3).contiguous()
            image = self._probe_of(lights[n], 'image') if with_image else None
            if image is not None:
                image = torch.as_tensor(image).to(device=eng.device, dtype=torch.float32)
                image = (image[0] if image.ndim == 4 else image).contiguous()
            if self.rotate_ratio <= 0:
                yield (n, probe, image) if with_image else (n, probe)
                continue
            n_rot = ew * self.rotate_ratio
            # rotations per call: bounded device memory of the sweep (an 8k env-map image is 400 MB per rotation)
            step = 16 if image is None else max(1, min(16, (1 << 28) // max(image.numel(), 1)))
            for j0 in range(0, n_rot, step):
                rot = eng.rotate_probes(probe, self.rotate_ratio, j0, min(step, n_rot - j0))
                rimg = eng.rotate_image(image, self.rotate_ratio, j0, rot.shape[0]) if image is not None else None
                for k in range(rot.shape[0]):
                    name = f'{n}-{j0 + k:04d}'
                    yield (name, rot[k], rimg[k] if rimg is not None else None) if with_image else (name, rot[k])

    def _ensure_capacity(self, P: int) -> None:
        """A frame with more rays than the handle was created for (e.g. the first 1024^2 frame under a cfg that carries no image
        size) gets a larger handle: workspaces are sized by ra_config.max_rays at ra_create."""
        eng = self.engine
        if P <= eng.config['max_rays'] or not isinstance(eng, Engine):
            return
        conf = dict(eng.config, max_rays=int(P * 1.25) + 1024)
        dev = eng.device
        eng.close()
        self.engine = Engine(conf, dev)
        self.engine.upload_weights(self.net.state_dict())

    @torch.no_grad()
    def render(self, batch) -> dotdict:
        self._ensure_capacity(int(torch.as_tensor(batch['ray_o']).shape[-2]))
        eng = self.engine
        eng.set_frame(batch, self.fix_material, self.always_fix_material)
        ray_o, ray_d, near, far = eng._rays(batch)
        P = ray_o.shape[0]
        keys = list(_MAIN_KEYS[self.mode])
        if self.mode == 'relight' and self.return_lvis:
            keys += ['lvis_map', 'ldot_map']
        if self.mode == 'relight' and self.main_spec:
            keys += ['spec_map']
        if self.mode != 'relight':
            out = eng.render(self.mode, ray_o, ray_d, near, far, keys)
            return dotdict({k: v[None] for k, v in out.items()})
        # novel_light_sphere_tracing.Renderer.render (:101-221): main pass, then one cheap re-shade per env-map
        main_probe = eng.env_main
        if self.replace_light:
            main_probe = torch.as_tensor(self._probe_of((batch.get('novel_lights') or {})[self.replace_light])).to(device=eng.device, dtype=torch.float32)
            main_probe = (main_probe[0] if main_probe.ndim == 4 else main_probe).contiguous()
            eng.set_main_light(main_probe)
        if self.sync_timing:
            torch.cuda.synchronize(eng.device)
        tick = time.perf_counter()
        main = eng.render('relight', ray_o, ray_d, near, far, keys)
        if self.sync_timing:
            torch.cuda.synchronize(eng.device)
        diff = time.perf_counter() - tick
        relight = dotdict()
        conv = (lambda t: t.cpu()) if self.to_cpu else (lambda t: t)
        if self.ground_shading:
            if eng.config.get('visibility_mode') or self.replace_light:
                raise NotImplementedError('vis_ground_shading together with no_visibility / local_visibility / replace_light is not implemented')
            return self._render_with_ground(batch, main, P, diff, conv)
        # `main` of the reference (:126-158): the learned-light entry keeps the `visual` keys and stays on the device; every novel light
        # gets `{**main, **human}` (so also ray_o and, when kept, the (P,512) lvis / ldot maps) moved to the host by to_cpu (:216).
        # The main maps are moved once and shared by all lights instead of once per light.
        if 'main' in self.test_light:
            relight.main = dotdict({k: v[None] for k, v in main.items() if k not in ('lvis_map', 'ldot_map')})
            relight.main.envmap = dotdict(probe=main_probe[None])
        lights = batch.get('novel_lights') or {}
        sweep = list(self._light_sweep(lights, self._light_names(lights)))
        shared = None
        for c0 in range(0, len(sweep), 16):                     # re-shade in groups: the stored visibility is read once per 4 probes
            group = sweep[c0:c0 + 16]
            probes = torch.stack([p for _, p in group]).contiguous()
            rgb, shade, spec = eng.relight_envmaps(probes, P)
            if shared is None:
                shared = dotdict({k: conv(v[None]) for k, v in main.items()})
                shared.ray_o = conv(ray_o[None])
            for i, (n, _) in enumerate(group):
                human = dotdict(shared)
                human.update(rgb_map=conv(rgb[i][None]), shade_map=conv(shade[i][None]), spec_map=conv(spec[i][None]))
                human.envmap = dotdict(probe=conv(probes[i][None]))
                relight[n] = human
        relight.diff = diff
        return relight

    _VISUAL = ('rgb_map', 'acc_map', 'norm_map', 'surf_map', 'bpts_map', 'cpts_map', 'spec_map', 'shade_map', 'depth_map', 'albedo_map', 'roughness_map')

    def _render_with_ground(self, batch, main, P, diff, conv) -> dotdict:
        """The vis_ground_shading + vis_novel_light path (sphere_tracing_renderer.py:1079-1107, novel_light_sphere_tracing.py:157-216):
        floor pass over every pixel, then per light blend_output_(ground, human) into image-sized maps (mask_at_box becomes all-True)."""
        eng = self.engine
        dev = eng.device
        t = lambda k: torch.as_tensor(batch[k]).to(device=dev, dtype=torch.float32)
        meta = batch.get('meta') or {}
        H = int(torch.as_tensor(meta['H'] if 'H' in meta else batch['H']).reshape(-1)[0])
        W = int(torch.as_tensor(meta['W'] if 'W' in meta else batch['W']).reshape(-1)[0])
        acc_g = eng.ground_begin(torch.as_tensor(batch['mask_at_box'])[0], main['acc_map'])
        ray_o, ray_d = get_rays(H, W, t('cam_K'), t('cam_R'), t('cam_T'))
        ground = eng.render_ground(self.ground, ray_o, ray_d, acc_g, eng.env_main)

        def blend(grd: Dict, human: Dict, conv=conv) -> dotdict:
            out = dotdict()
            for k in self._VISUAL + ('lvis_map', 'ldot_map'):
                if k == 'acc_map' or (k not in grd and k not in human):
                    continue
                if k in ('lvis_map', 'ldot_map') and not self.return_lvis:
                    continue
                out[k] = conv(eng.blend_ground(acc_g, grd.get(k), human.get(k) if k in self._VISUAL else None, True)[None])
            out.acc_map = conv(eng.blend_ground(acc_g, None, main['acc_map'], False)[None])
            return out

        relight = dotdict()
        human_main = {k: main[k] for k in self._VISUAL if k in main}
        if 'main' in self.test_light:
            relight.main = blend(ground, human_main, conv=lambda t: t)          # the learned-light entry stays on the device (:155-158)
            relight.main.envmap = dotdict(probe=eng.env_main[None])
        lights = batch.get('novel_lights') or {}
        # every env-map once, or (cfg.vis_rotate_light) at rotate_ratio * env_w rotations of probe AND attached image
        group = []

        def flush():
            if not group:
                return
            probes = torch.stack([p for _, p, _ in group]).contiguous()
            rgb, shade, spec = eng.relight_envmaps_raw(probes, P)
            for i, (n, _, img) in enumerate(group):
                human = dict(human_main)
                human.update(rgb_map=rgb[i], shade_map=shade[i], spec_map=spec[i])
                g_rgb, g_alb, g_shade, g_spec = eng.relight_ground(self.ground, probes[i], ray_d, ground, img)
                grd = {k: ground[k] for k in self._VISUAL if k in ground}
                grd.update(rgb_map=g_rgb, albedo_map=g_alb, shade_map=g_shade, spec_map=g_spec)
                relight[n] = blend(grd, human)
                relight[n].envmap = dotdict(probe=conv(probes[i][None]))
            group.clear()

        for item in self._light_sweep(lights, self._light_names(lights), with_image=True):
            group.append(item)
            if len(group) == 4:              # the human re-shade reads the stored visibility once per 4 probes
                flush()
        flush()
        if isinstance(batch.get('mask_at_box'), torch.Tensor):
            batch['mask_at_box'][:] = True           # later used for visualization (:1101)
        relight.diff = diff
        return relight

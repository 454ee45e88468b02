"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference renderer from the read-only tree under
import shims on a synthetic scene and dumps its outputs as golden vectors.

The reference cannot be imported as-is here (easymocap / pytorch3d / smplx / termcolor / pdbr ...
are not installed, no network).  A `sys.meta_path` finder hands out permissive stub modules for
those names; the only third-party *arithmetic* on the path, `pytorch3d.ops.knn_points`, is
restated as an exact squared-L2 K-NN.  Nothing is copied from the reference: it is imported from
where it lies (`$RA_REFERENCE`, /root/reference or baseline/_ref).

One process per config (hot-path parameters are bound as default args at import time):

    python oracle/ref_harness.py --mode relight|anisdf_trace|anisdf_volume --H 48 --out x.npz

Never imported by the product; used by tests/golden/make_golden.py only.
"""
from __future__ import annotations

import argparse
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

STUB_ROOTS = ('easymocap', 'pytorch3d', 'smplx', 'termcolor', 'pdbr', 'trimesh', 'mcubes', 'imageio', 'skimage',
              'lpips', 'torch_scatter', 'h5py', 'ujson', 'matplotlib', 'cupy', 'cupy_knn', 'open3d', 'sympy_stub',
              'bvh_distance_queries', 'nvdiffrast', 'largesteps', 'spconv', 'rich_stub', 'tensorboardX', 'plyfile',
              'pymeshlab', 'pyrender', 'OpenGL', 'glfw')


class _Anything:
    """Callable, attribute-able, subclass-able placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def find_reference() -> str:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for c in (os.environ.get('RA_REFERENCE'), '/root/reference', os.path.join(here, 'baseline', '_ref')):
        if c and os.path.exists(os.path.join(c, 'lib', 'networks', 'renderer', 'sphere_tracing_renderer.py')):
            return c
    raise FileNotFoundError('reference tree not found')


def install_shims():
    import torch
    sys.meta_path.insert(0, _StubFinder())
    import termcolor
    termcolor.colored = lambda s, *a, **k: s
    import pytorch3d.ops as p3o

    def knn_points(p1, p2, K=1, return_nn=False, return_sorted=True, **kw):
        outs_d, outs_i = [], []
        for s in range(0, p1.shape[1], 8192):
            q = p1[:, s:s + 8192]
            d = (q[:, :, None, 0] - p2[:, None, :, 0]) ** 2
            d = d + (q[:, :, None, 1] - p2[:, None, :, 1]) ** 2
            d = d + (q[:, :, None, 2] - p2[:, None, :, 2]) ** 2
            v, i = torch.topk(d, K, dim=-1, largest=False, sorted=True)
            outs_d.append(v); outs_i.append(i)
        if not outs_d:
            return p1.new_zeros(p1.shape[0], 0, K), torch.zeros(p1.shape[0], 0, K, dtype=torch.long), None
        return torch.cat(outs_d, 1), torch.cat(outs_i, 1), None

    p3o.knn_points = knn_points
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None


def setup_reference(mode: str, n_bones: int = 52):
    """Import the reference config and replay the cascade of lib/config/config.py:498-517 by hand."""
    root = find_reference()
    install_shims()
    sys.path.insert(0, root)
    sys.argv = ['oracle']
    os.chdir(root)
    from lib.config import cfg, yacs
    cfg.merge_strain(yacs.load_cfg(open('configs/mobile_stage/xuzhen_12v_geo.yaml', 'r')))
    if mode in ('relight', 'relight_ground'):
        cfg.relighting = True; cfg.vis_novel_light = True; cfg.vis_pose_sequence = True
        if mode == 'relight_ground':            # the README showcase option (readme.md:64): ground-plane shading, row f2
            cfg.vis_ground_shading = True
            cfg.store_alpha_channel = False     # what parse_cfg would set (config.py:451-452)
        cfg.merge_from_other_cfg(cfg.relighting_cfg)
        cfg.merge_from_other_cfg(cfg.pose_seq_cfg)
        cfg.merge_from_other_cfg(cfg.novel_light_cfg)
    elif mode == 'anisdf_trace':
        cfg.vis_pose_sequence = True; cfg.vis_sphere_tracing = True
        cfg.merge_from_other_cfg(cfg.pose_seq_cfg)
        cfg.merge_from_other_cfg(cfg.sphere_tracing_cfg)
    elif mode == 'anisdf_volume':
        cfg.vis_pose_sequence = True
        cfg.merge_from_other_cfg(cfg.pose_seq_cfg)
    else:
        raise ValueError(mode)
    cfg.n_bones = n_bones              # what parse_cfg derives from the body model (config.py:441,465-466): 52 SMPL-H, 24 SMPL
    cfg.cond_dim = 3 * n_bones
    cfg.vis_rendering_map = True
    cfg.probe_size_ratio = 0.0
    cfg.geometry_pretrain = '/nonexistent'
    return cfg


def to_ref_batch(b: dict, device='cpu'):
    import torch
    from lib.utils.base_utils import dotdict
    out = dotdict()
    for k, v in b.items():
        if k == 'novel_lights':
            out.novel_lights = dotdict({n: dotdict(probe=torch.from_numpy(p).to(device)) for n, p in v.items()})
        elif k == 'train_poses':
            out.train_motion = dotdict(poses=torch.from_numpy(v).to(device))
        elif isinstance(v, np.ndarray) and v.ndim > 0:
            out[k] = torch.from_numpy(v.copy()).to(device)
    out.meta = dotdict(H=torch.tensor([int(b['H'])]), W=torch.tensor([int(b['W'])]))
    if 'novel_lights' not in out:
        out.novel_lights = dotdict()          # the novel-light renderer loops over it unconditionally (novel_light_sphere_tracing.py:165)
    return out


def run(mode: str, H: int, seed: int, n_env: int, fitted: bool, threads: int = 0, frame: int = 0, azim_deg: float = 20.0,
        cam_dist: float = 3.0, n_bones: int = 52, tonemapping: bool = True, fix_material: int = 0, always_fix_material: bool = True):
    import torch
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    cfg = setup_reference(mode, n_bones)
    cfg.fix_material, cfg.always_fix_material = int(fix_material), bool(always_fix_material)      # base_network.py:501-503
    cfg.tonemapping_rendering = bool(tonemapping)      # False = what parse_cfg sets for vis_ext .exr / .hdr (config.py:446-448)
    from relightableavatar_b200 import scene
    if threads:
        torch.set_num_threads(threads)
    if mode.startswith('relight'):
        cfg.test_light = ['main'] + list(scene.make_envmaps(n_env, 10 + seed).keys()) if n_env else ['main']
    from lib.networks.make_network import make_network
    from lib.networks.renderer.make_renderer import make_renderer
    net = make_network(cfg)
    sd = scene.make_state_dict(seed, relight=mode.startswith('relight'), fitted=fitted, n_bones=n_bones)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    missing = [m for m in missing if 'freq_bands' not in m and 'embedder' not in m]
    assert not unexpected, unexpected
    assert not missing, missing
    net.eval()
    renderer = make_renderer(cfg, net)
    b = scene.make_batch(H, H, frame=frame, n_frames=frame + 1, seed=seed, n_env=n_env if mode.startswith('relight') else 0,
                         cam_dist=cam_dist, azim_deg=azim_deg, n_bones=n_bones)
    batch = to_ref_batch(b)
    torch.manual_seed(0)        # compute_ground_tris draws a random tangent (net_utils.py:392-396)
    with torch.no_grad():
        out = renderer.render(batch)
    flat = {}

    def put(prefix, d):
        for k, v in d.items():
            if isinstance(v, torch.Tensor):
                flat[prefix + k] = v.detach().cpu().numpy()
            elif isinstance(v, dict):
                put(prefix + k + '.', v)

    put('', out)
    flat['wbounds_after'] = batch.wbounds.cpu().numpy()
    return flat, b


def run_prep(H: int, seed: int, frame: int = 2):
    """Row f1 fixtures: the reference's own in-repo dataset-side functions (numpy / torch, importable under the shims) on the
    synthetic scene: get_rays_within_bounds, get_bounds (lib/utils/data_utils.py), tpose_points_to_pose_points,
    pose_points_to_world_points (lib/utils/blend_utils.py).  get_rigid_transform needs smplx (absent): A comes from the scene."""
    import torch
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    root = find_reference()
    install_shims()
    sys.path.insert(0, root); sys.argv = ['oracle']; os.chdir(root)
    from lib.utils import data_utils as DU, blend_utils as BU
    from relightableavatar_b200 import scene
    b = scene.make_batch(H, H, frame=frame, n_frames=frame + 1, seed=seed, n_env=0)
    body = scene.make_body(seed)
    ro, rd, near, far, mask = DU.get_rays_within_bounds(H, H, b['cam_K'][0], b['cam_R'][0], b['cam_T'][0], b['wbounds'][0])
    rv = torch.as_tensor(body.rverts, dtype=torch.float32)[None]
    pp = BU.tpose_points_to_pose_points(rv, torch.as_tensor(b['weights']), torch.as_tensor(b['A']))
    ww = BU.pose_points_to_world_points(pp, torch.as_tensor(b['R']), torch.as_tensor(b['Th']))
    return dict(ray_o=ro, ray_d=rd, near=near, far=far, mask_at_box=mask, pverts=pp[0].numpy(), wverts=ww[0].numpy(),
                wbounds=DU.get_bounds(ww[0].numpy()), pbounds=DU.get_bounds(pp[0].numpy()), _frame=np.int64(frame))


def run_rotate(seed: int = 0, repeat: int = 4):
    """Row f4 fixture: the reference's own rotate_envmap (lib/utils/relight_utils.py:55-103) on two synthetic probes."""
    import torch
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    setup_reference('relight')
    from lib.utils.base_utils import dotdict
    from lib.utils.relight_utils import rotate_envmap
    from relightableavatar_b200 import scene
    maps = scene.make_envmaps(2, 10 + seed)
    novel = dotdict({k: dotdict(probe=torch.from_numpy(v)[None], image=torch.from_numpy(v)[None]) for k, v in maps.items()})
    eW = next(iter(maps.values())).shape[1]
    idx = [0, 1, 5, 63, eW * repeat - 1, eW * repeat, eW * repeat + 7, 2 * eW * repeat - 1]
    out = {'_repeat': np.int64(repeat), '_index': np.asarray(idx, np.int64)}
    for i in idx:
        name, env = rotate_envmap(novel, i, repeat, eW, eW)
        out[f'probe_{i}'] = env.probe[0].numpy()
        out[f'name_{i}'] = np.asarray(name)
    return out


VISUAL_CASES = {'main': ('rendering', 'normal', 'alpha', 'depth', 'shading', 'albedo', 'roughness', 'surface', 'residual', 'envmap'),
                'sky00': ('rendering', 'specular', 'shading')}


def run_visual(fixture: str = 'relight_48'):
    """Row f3 fixture: the reference's own Visualizer.generate_image (lib/visualizers/base_visualizer.py:55-231, incl. add_light_probe and
    the alpha channel) applied to the maps the reference rendered for tests/golden/<fixture>.npz, for every Output type the path can
    produce; plus save_image's quantisation (lib/utils/data_utils.py:689-709) restated on two of the images through cv2-free numpy
    (save_image itself only adds the cv2.imwrite call)."""
    import torch
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    g = dict(np.load(os.path.join(here, 'tests', 'golden', fixture + '.npz')))
    cfg = setup_reference('relight')
    cfg.probe_size_ratio = 0.2          # config.py:354 (setup_reference switches the overlay off for the rendering fixtures)
    cfg.store_alpha_channel = True
    from lib.config.config import Output
    from lib.utils.base_utils import dotdict
    from lib.visualizers.base_visualizer import Visualizer
    from lib.utils import relight_utils as RU
    from relightableavatar_b200 import scene
    if not torch.cuda.is_available():      # gen_light_xyz defaults to device='cuda' (relight_utils.py:423); same arithmetic on the CPU
        _glx = RU.gen_light_xyz
        RU.gen_light_xyz = lambda h, w, envmap_r=1e2, device='cpu': _glx(h, w, envmap_r, device='cpu')
    H = int(g['_H'])
    b = scene.make_batch(H, H, seed=int(g['_seed']), n_env=int(g['_n_env']))
    batch = to_ref_batch(b)
    main = dotdict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('main.') and k != 'main.envmap.probe'})
    main.envmap = dotdict(probe=torch.from_numpy(g['main.envmap.probe']))
    outs = {'main': main}
    for n in b['novel_lights']:
        o = dotdict(main)
        o.update({k[len(n) + 1:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(n + '.')})
        o.envmap = dotdict(probe=torch.from_numpy(b['novel_lights'][n]))
        outs[n] = o
    flat = {'_fixture': np.asarray(fixture)}
    for n, types in VISUAL_CASES.items():
        for t in types:
            ty = next(o for o in Output if o.name.lower() == t)
            img = Visualizer.generate_image(dotdict(outs[n]), batch, ty)
            flat[f'{n}.{t}'] = np.asarray(img, np.float32)
    return flat


CFG_KEYS = ('dist_th', 'blend_radius', 'resd_limit', 'env_r', 'render_chunk_size', 'n_samples', 'surf_sample_range', 'fresnel_f0',
            'albedo_slope', 'albedo_bias', 'roughness_slope', 'roughness_bias', 'albedo_multiplier', 'shading_albedo', 'env_h', 'env_w',
            'clip_near', 'clip_far', 'tonemapping_rendering', 'ground_normal', 'ground_origin', 'ground_albedo', 'ground_attach_envmap', 'ground_shading_multiplier')
CFG_GROUPS = {'sphere_tracing': ('iter', 'tan_i', 'relax', 'offset', 'eps', 'shadow_skip_iter'),
              'obj_lvis': ('iter', 'offset', 'relax', 'near_offset', 'dist_th'),
              'env_lvis': ('bbox_margin', 'iter', 'offset', 'relax', 'near_offset', 'dist_th')}


def dump_cfg(mode: str) -> dict:
    """The values the reference's renderers read from the global cfg for this path, after its own config cascade
    (lib/config/config.py + configs/mobile_stage/xuzhen_12v_geo.yaml): pins relightableavatar_b200.renderer.default_config."""
    cfg = setup_reference(mode)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    plain = lambda v: list(v) if isinstance(v, (list, tuple)) else (v.item() if hasattr(v, 'item') else v)
    out = {k: plain(cfg[k]) for k in CFG_KEYS}
    for g, keys in CFG_GROUPS.items():
        out[g] = {k: plain(cfg[g][k]) for k in keys}
    # the switches the drop-in supports at their default only (renderer.FIXED_SWITCHES): record what the reference has
    from relightableavatar_b200.renderer import FIXED_ST_SWITCHES, FIXED_SWITCHES
    out.update({k: plain(cfg[k]) for k in FIXED_SWITCHES})
    out['sphere_tracing'].update({k: plain(cfg.sphere_tracing[k]) for k in FIXED_ST_SWITCHES})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', required=True)
    ap.add_argument('--H', type=int, default=48)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--n_env', type=int, default=2)
    ap.add_argument('--raw_init', action='store_true', help='geometric-init SDF instead of the fitted one')
    ap.add_argument('--frame', type=int, default=0, help='pose frame of the synthetic motion')
    ap.add_argument('--azim', type=float, default=20.0, help='camera azimuth (deg)')
    ap.add_argument('--cam_dist', type=float, default=3.0)
    ap.add_argument('--n_bones', type=int, default=52, help='52 = SMPL-H (xuzhen), 24 = SMPL (ZJU-MoCap / synthetic-human configs)')
    ap.add_argument('--fix_material', type=int, default=0)
    ap.add_argument('--no_always_fix_material', action='store_true')
    ap.add_argument('--slim', action='store_true', help='store only the rgb / acc maps (full-size runs: the (P,512) visibility maps are 141 MB each)')
    ap.add_argument('--linear', action='store_true', help='cfg.tonemapping_rendering False (HDR output)')
    ap.add_argument('--out', required=True)
    a = ap.parse_args()
    out_path = os.path.abspath(a.out)
    if a.mode.startswith('cfg_'):
        import json
        json.dump(dump_cfg(a.mode[4:]), open(out_path, 'w'), indent=1, sort_keys=True)
        print('wrote', out_path)
        return
    if a.mode == 'rotate':
        flat = run_rotate(a.seed)
    elif a.mode == 'prep':
        flat = run_prep(a.H, a.seed)
    elif a.mode == 'visual':
        flat = run_visual()
    else:
        flat, _ = run(a.mode, a.H, a.seed, a.n_env, not a.raw_init, frame=a.frame, azim_deg=a.azim, cam_dist=a.cam_dist, n_bones=a.n_bones, tonemapping=not a.linear,
                      fix_material=a.fix_material, always_fix_material=not a.no_always_fix_material)
    if a.slim:
        flat = {k: v for k, v in flat.items() if k.endswith('rgb_map') or k.endswith('acc_map')}
    np.savez_compressed(out_path, **flat)
    print('wrote', out_path, {k: v.shape for k, v in flat.items()})


if __name__ == '__main__':
    main()

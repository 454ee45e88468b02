"""Walks every kernel of the path once on small inputs, without the oracle: meant to run under
`compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py` (gpurun; not a test).
Prints one line per stage so a report can be attributed to the stage that launched it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer
from relightableavatar_b200.prepare import FramePreparer

DEV = 'cuda:0'
H = int(sys.argv[1]) if len(sys.argv) > 1 else 24


def stage(name):
    torch.cuda.synchronize()
    print('STAGE_OK', name, flush=True)


b = scene.make_batch(H, H, seed=0, n_env=2)
sd = scene.make_state_dict(0, relight=True, fitted=True)
for prec in ('tc', 'fp32'):
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision=prec, max_rays=2048, test_light=('main', 'all'),
                 return_lvis=True, sync_timing=False)
    out = r.render(b)
    assert torch.isfinite(out['main']['rgb_map']).all()
    stage(f'relight {prec}')
    if prec == 'fp32':
        name, probe = next(iter(b['novel_lights'].items()))
        rot = r.engine.rotate_probes(torch.as_tensor(probe[0]), 4, 0, 5)
        r.engine.relight_envmaps(rot, b['ray_o'].shape[1])
        stage('rotation sweep')
        r.engine.assemble_image(out['main']['rgb_map'][0], out['main']['acc_map'][0], torch.as_tensor(b['mask_at_box'][0]))
        stage('image assembly')
        x = torch.as_tensor(b['wverts'][0][::7]).float() + 0.02
        v = torch.nn.functional.normalize(torch.randn(x.shape[0], 3), dim=-1)
        r.engine.query_sdf(x, 0.125, True); r.engine.query_raw(x, v)
        stage('query_sdf / query_raw')
        body = scene.make_body(0)
        poses, Rh, _ = scene.make_motion(1, 1)
        g = torch.Generator().manual_seed(0)
        faces = torch.randint(0, body.rverts.shape[0], (3000, 3), generator=g).numpy().astype(np.int32)
        for kw in ({'rnorm': body.rnorm, 'tnorm': body.tnorm}, {'faces': faces}):
            prep = FramePreparer(r.engine, body.joints, body.parents, body.rverts, body.weights, body.big_A, body.tverts, **kw)
            gb = prep.make_batch(poses[0], Rh[0], b['Th'][0, 0], b['cam_K'][0], b['cam_R'][0], b['cam_T'][0], H, H, extra={'train_poses': b['train_poses']})
        r.render(gb)
        stage('batch preparation + render')
    r.engine.close()

# fewer rays than one tile, and none at all
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision='tc', max_rays=2048, test_light=('main', 'all'), sync_timing=False)
for n in (5, 0):
    bb = {k: (v[:, :n] if k in ('ray_o', 'ray_d', 'near', 'far') else v) for k, v in b.items()}
    r.render(bb)
    stage(f'{n} rays')
r.engine.close()

r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision='fp32', max_rays=2048, test_light=('main', 'all'),
             ground_shading=True, sync_timing=False)
r.render(b)
stage('ground shading')
r.engine.close()

sd_a = scene.make_state_dict(0, relight=False, fitted=True)
for mode in ('anisdf_trace', 'anisdf_volume'):
    try:
        r = Renderer(scene.SyntheticNet(sd_a, False), mode=mode, device=DEV, precision='fp32', max_rays=2048, sync_timing=False)
    except Exception as e:                                   # mode names are the renderer's; report, don't hide
        print('STAGE_SKIPPED', mode, repr(e)); continue
    r.render(b)
    stage(mode)
    r.engine.close()
# round-2 entry points: image side (every Output type, overlay, 16-bit), rotation sweep with the floor, ablation switches,
# replaced main light, the 3-NN query, the tensor-memory MLP variant
from relightableavatar_b200.visualizer import Visualizer
from relightableavatar_b200.renderer import Engine, default_config
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision='tc', max_rays=2048, test_light=('main', 'all'), sync_timing=False,
             vis_specular_map=True)
out = r.render(b)
vis = Visualizer(r.engine)
bb = dict(b); bb['tbounds'] = b['tbounds']
for t in ('rendering', 'normal', 'alpha', 'depth', 'shading', 'albedo', 'roughness', 'surface', 'residual', 'specular', 'envmap'):
    vis.generate_image(out['main'], bb, t)
vis.encode_frame(out, bb, types=('rendering', 'normal'), ext='.png')
vis.encode_frame(out, bb, types=('rendering',), ext='.jpg')
stage('visualizer')
ids, d2 = r.engine.query_knn(torch.as_tensor(b['wverts'][0][::5]).float() + 0.3)
stage('query_knn')
pts = (torch.as_tensor(b['wverts'][0][::40]).float()[:, None] + 0.4 + torch.rand(1, 32, 3) * 0.05).reshape(-1, 3)
r.engine.query_knn(pts, packets=True)
stage('query_knn packets')
r.engine.close()
name = next(iter(b['novel_lights']))
for over in (dict(visibility_mode=1), dict(visibility_mode=2), dict(brdf_mode=1), dict(brdf_mode=2), dict(replace_light=name)):
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision='tc', max_rays=2048, test_light=('main', 'all'), sync_timing=False, **over)
    r.render(b)
    r.engine.close()
stage('ablation switches')
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision='tc', max_rays=2048, test_light=('all',), ground_shading=True,
             sync_timing=False, rotate_ratio=1)
bl = dict(b); bl['novel_lights'] = {name: {'probe': b['novel_lights'][name], 'image': np.repeat(b['novel_lights'][name], 2, axis=2)}}
r.render(bl)
stage('rotation sweep with ground shading')
r.engine.close()
for variant in ('7', '8'):
    os.environ['RA_TC_VARIANT'] = variant
    eng = Engine(default_config(True, precision=1, max_rays=2048), DEV)
    eng.upload_weights(sd); eng.set_frame(b)
    eng.query_sdf(torch.as_tensor(b['wverts'][0][::3]).float() + 0.01, 0.125, True)
    eng.query_sdf((torch.as_tensor(b['wverts'][0]).float()[:, None] + torch.randn(1, 6, 3) * 0.02).reshape(-1, 3), 0.125, True)      # > 148 tiles: two slots
    stage(f'k_mlp_tc{variant}')
    eng.close()
os.environ.pop('RA_TC_VARIANT')
from relightableavatar_b200 import parallel
pool = parallel.FramesInFlight(lambda: Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision='tc', max_rays=2048,
                                                test_light=('main',), sync_timing=False, ground_shading=True), 2)
for o in pool.render_sequence([b, b, b]):
    pass
pool.close()
stage('frames in flight with ground shading')
print('SANITIZE_DONE')

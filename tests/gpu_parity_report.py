"""Prints the measured values behind the parity assertions of tests/test_gpu_parity.py (run under gpurun; a checker, not a
collected test): PSNR against the full-size reference fixtures in both precisions, error quantiles / maxima of every map at
64x64 against the oracle, visibility-map flip statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import ra_oracle as O
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer

DEV = 'cuda:0'
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def pixels(fixture, precision):
    g = dict(np.load(os.path.join(GOLD, fixture + '.npz')))
    Hh, n_env, frame = int(g['_H']), int(g['_n_env']), int(g.get('_frame', 0))
    b = scene.make_batch(Hh, Hh, frame=frame, n_frames=frame + 1, seed=0, n_env=n_env)
    sd = scene.make_state_dict(0, relight=True, fitted=True)
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision=precision, max_rays=b['ray_o'].shape[1] + 8,
                 test_light=('main', 'all'), sync_timing=False)
    out = r.render(b)
    ref_acc = torch.from_numpy(g['main.acc_map'][0].astype(np.float32))
    flips = int(((ref_acc > 0) != (out['main']['acc_map'][0].cpu() > 0)).sum())
    for name in ['main'] + list(b.get('novel_lights', {})):
        ref = torch.from_numpy(g[f'{name}.rgb_map'][0].astype(np.float32))
        got = out[name]['rgb_map'][0].cpu()
        e = (got - ref).abs()
        print(f'{fixture} [{precision}] {name}: PSNR {O.psnr(O.assemble_image(b, got), O.assemble_image(b, ref)):.2f} dB, flips {flips} of {int((ref_acc > 0).sum())}, '
              f'rgb q99 {torch.quantile(e.flatten()[::3], 0.99):.3e} max {e.max():.3e}', flush=True)
    r.engine.close()


for fx, precs in (('relight_512_pixels', ('tc', 'fp32')), ('relight_1024_f5_pixels', ('tc',))):
    for p in precs:
        pixels(fx, p)

sd_a = scene.make_state_dict(0, relight=False, fitted=True)
for fixture, mode, Hh in (('anisdf_trace_128', 'anisdf_trace', 128), ('anisdf_volume_512_pixels', 'anisdf_volume', 512)):
    g = dict(np.load(os.path.join(GOLD, fixture + '.npz')))
    b = scene.make_batch(Hh, Hh, seed=0, n_env=0)
    r = Renderer(scene.SyntheticNet(sd_a, False), mode=mode, device=DEV, precision='fp32', max_rays=b['ray_o'].shape[1] + 8)
    out = r.render(b)
    ref = torch.from_numpy(g['rgb_map'][0].astype(np.float32))
    e = (out['rgb_map'][0].cpu() - ref).abs()
    print(f'{fixture}: PSNR {O.psnr(O.assemble_image(b, out["rgb_map"][0].cpu()), O.assemble_image(b, ref)):.2f} dB, rgb q98 {torch.quantile(e.flatten()[::3], 0.98):.3e} max {e.max():.3e}', flush=True)
    r.engine.close()

b = scene.make_batch(64, 64, seed=0, n_env=2)
sd = scene.make_state_dict(0, relight=True, fitted=True)
probes = {k: v[0] for k, v in b['novel_lights'].items()}
ref = O.render_novel_light(b, sd, O.Cfg(), probes, torch.float32, DEV)
for prec in ('fp32', 'tc'):
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision=prec, max_rays=8192, test_light=('main', 'all'), return_lvis=True)
    out = r.render(b)
    main = out['main']
    both = ((main['acc_map'][0] > 0) & (ref['main']['acc_map'] > 0))
    for k in ('rgb_map', 'acc_map', 'surf_map', 'albedo_map', 'roughness_map', 'shade_map', 'norm_map', 'depth_map', 'cpts_map', 'bpts_map'):
        e = (main[k][0] - ref['main'][k]).abs()
        ef = e[both]
        print(f'[{prec}] main.{k:14s} q98 {torch.quantile(e.flatten(), .98):.3e} q99.8 {torch.quantile(e.flatten(), .998):.3e} max {e.max():.3e} | fg-agree max {ef.max():.3e}')
    any_light = next(iter(probes))
    for k in ('lvis_map', 'ldot_map'):
        got = out[any_light][k][0] if k in out[any_light] else None
        full = ref['_main_full'][k]          # already acc-premultiplied (a blend key)
        e = (got - full).abs()
        print(f'[{prec}] {k}: mean {e.mean():.3e} q98 {torch.quantile(e.flatten()[::5], .98):.3e} q99.9 {torch.quantile(e.flatten()[::5], .999):.3e} max {e.max():.3e} frac>0.05 {(e > 0.05).float().mean():.3e}')
    for n in probes:
        print(f'[{prec}] PSNR {n} {O.psnr(O.assemble_image(b, out[n]["rgb_map"][0].cpu()), O.assemble_image(b, ref[n]["rgb_map"].cpu())):.2f}')
    print(f'[{prec}] PSNR main {O.psnr(O.assemble_image(b, main["rgb_map"][0].cpu()), O.assemble_image(b, ref["main"]["rgb_map"].cpu())):.2f}', flush=True)
    r.engine.close()
# smoke-sized frame (32x32, tc)
b = scene.make_batch(32, 32, seed=0, n_env=1)
probes = {k: v[0] for k, v in b['novel_lights'].items()}
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision='tc', max_rays=4096, test_light=('main', 'all'))
out = r.render(b)
ref = O.render_novel_light(b, sd, O.Cfg(), probes, torch.float32, DEV)
n = next(iter(probes))
print(f'smoke 32x32 tc PSNR {O.psnr(O.assemble_image(b, out[n]["rgb_map"][0].cpu()), O.assemble_image(b, ref[n]["rgb_map"].cpu())):.2f}')

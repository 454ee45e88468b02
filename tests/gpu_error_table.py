"""Prints an error table CUDA-vs-oracle (run under gpurun: python tests/gpu_error_table.py [fp32|tc] [H]; a checker, not a collected test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ra_oracle as O
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Engine, Renderer, default_config

DEV = 'cuda:0'
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
b = scene.make_batch(H, H, seed=0, n_env=2)
sd = scene.make_state_dict(0, relight=True, fitted=True)
cfg = O.Cfg()
eng = Engine(default_config(True, precision={'fp32': 0, 'tc': 1}[prec], max_rays=max(8192, b['ray_o'].shape[1])), DEV)
eng.upload_weights(sd); eng.set_frame(b)
g = torch.Generator().manual_seed(0)
wv = torch.as_tensor(b['wverts'][0])
idx = torch.randint(0, wv.shape[0], (20000,), generator=g)
x = torch.cat([wv[idx] + torch.randn(20000, 3, generator=g) * 0.06, wv[idx[:2000]] + torch.randn(2000, 3, generator=g) * 0.5]).float()
got = eng.query_sdf(x, 0.125, True)
W = O.Weights(sd, torch.float32, DEV); fr = O.Frame.from_batch(b, cfg, torch.float32, DEV)
with torch.no_grad():
    ref = O.hdq_distance(x.to(DEV), fr, W, cfg, 0.125, True)[:, 0]
e = (got - ref).abs()
print(f'[{prec}] query_sdf: q50 {e.median():.3e} q99 {torch.quantile(e, .99):.3e} max {e.max():.3e} frac>1e-3 {(e > 1e-3).float().mean():.3e}  stats {eng.stats()}')
if prec == 'fp32':
    v = torch.nn.functional.normalize(torch.randn(x.shape[0], 3, generator=g), dim=-1)
    graw = eng.query_raw(x[:6000], v[:6000])
    rraw = O.network_forward(x[:6000].to(DEV), v[:6000].to(DEV), fr, W, cfg)
    nzg, nzr = graw.abs().sum(-1) > 0, rraw.abs().sum(-1) > 0
    print('query_raw: in-shell mismatch', (nzg != nzr).sum().item(), 'of', nzr.sum().item())
    both = nzg & nzr
    er = (graw - rraw).abs()[both]
    names = ['cpts'] * 3 + ['bpts'] * 3 + ['resd'] * 3 + ['albedo'] * 3 + ['rough'] + ['norm'] * 3 + ['occ']
    for c0, c1, nm in ((0, 3, 'cpts'), (3, 6, 'bpts'), (6, 9, 'resd'), (9, 12, 'albedo'), (12, 13, 'rough'), (13, 16, 'norm'), (16, 17, 'occ')):
        ee = er[:, c0:c1].flatten()
        print(f'   {nm:7s} q50 {ee.median():.3e} q99 {torch.quantile(ee, .99):.3e} max {ee.max():.3e}')
eng.close()
probes = {k: v[0] for k, v in b['novel_lights'].items()}
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=DEV, precision=prec, max_rays=max(8192, b['ray_o'].shape[1]), test_light=('main', 'all'), return_lvis=True)
t0 = time.time(); out = r.render(b); torch.cuda.synchronize(); t1 = time.time()
out = r.render(b); torch.cuda.synchronize(); t2 = time.time()
print(f'render {H}x{H}: first {t1 - t0:.3f}s second {t2 - t1:.3f}s stats {r.engine.stats()} launches {r.engine.launch_count()}')
t0 = time.time()
ref = O.render_novel_light(b, sd, cfg, probes, torch.float32, DEV)
torch.cuda.synchronize(); print(f'oracle on GPU: {time.time() - t0:.2f}s')
fg_g, fg_r = out['main']['acc_map'][0] > 0, ref['main']['acc_map'] > 0
print('fg', fg_g.sum().item(), fg_r.sum().item(), 'mismatch', (fg_g != fg_r).sum().item())
for k in ('rgb_map', 'acc_map', 'surf_map', 'albedo_map', 'roughness_map', 'shade_map', 'norm_map', 'depth_map', 'cpts_map'):
    ee = (out['main'][k][0] - ref['main'][k]).abs().flatten()
    print(f'   main.{k:14s} q50 {ee.median():.3e} q98 {torch.quantile(ee, .98):.3e} max {ee.max():.3e}')
ee = (out[next(iter(probes))]['lvis_map'][0] - ref['_main_full']['lvis_map']).abs().flatten()      # per-light dicts carry the maps (:189)
print(f'   lvis q98 {torch.quantile(ee[:2000000], .98):.3e} max {ee.max():.3e}')
for n in probes:
    for k in ('rgb_map', 'shade_map', 'spec_map'):
        ee = (out[n][k][0] - ref[n][k]).abs().flatten()
        print(f'   {n}.{k:10s} q98 {torch.quantile(ee, .98):.3e} max {ee.max():.3e}')
    print('   PSNR', n, O.psnr(O.assemble_image(b, out[n]['rgb_map'][0].cpu()), O.assemble_image(b, ref[n]['rgb_map'].cpu())))
print('   PSNR main', O.psnr(O.assemble_image(b, out['main']['rgb_map'][0].cpu()), O.assemble_image(b, ref['main']['rgb_map'].cpu())))

"""The five BASELINE.json configs on ONE GPU (the bench line is config 3; this is the per-config timing table of DESIGN.md §6).
Each config: 3 warm-up calls, then the mean of n timed Renderer.render calls (CUDA events).  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer


def timed(r, b, n):
    for _ in range(3):
        r.render(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r.render(b)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
sd_a = scene.make_state_dict(0, relight=False, fitted=True)
sd_r = scene.make_state_dict(0, relight=True, fitted=True)
b = scene.make_batch(128, 128, seed=0, n_env=0)
r = Renderer(scene.SyntheticNet(sd_a, False), mode='anisdf_trace', precision='tc', max_rays=b['ray_o'].shape[1] + 8, sync_timing=False)
res['1: AniSDF sphere-trace 128x128'] = dict(ms=timed(r, b, 20), rays=int(b['ray_o'].shape[1])); r.engine.close()
b = scene.make_batch(512, 512, seed=0, n_env=0)
r = Renderer(scene.SyntheticNet(sd_a, False), mode='anisdf_volume', precision='tc', max_rays=b['ray_o'].shape[1] + 8, sync_timing=False)
res['2: AniSDF volume render 512x512 (128 samples/ray)'] = dict(ms=timed(r, b, 3), rays=int(b['ray_o'].shape[1])); r.engine.close()
b = scene.make_batch(512, 512, seed=0, n_env=8)
r = Renderer(scene.SyntheticNet(sd_r, True), mode='relight', precision='tc', max_rays=b['ray_o'].shape[1] + 8, test_light=('main',), sync_timing=False)
res['3: relight 512x512, main env-map'] = dict(ms=timed(r, b, 10), rays=int(b['ray_o'].shape[1])); r.engine.close()
r = Renderer(scene.SyntheticNet(sd_r, True), mode='relight', precision='tc', max_rays=b['ray_o'].shape[1] + 8, test_light=('main', 'all'), sync_timing=False)
res['4: relight 512x512, main + 8 env-maps (one GPU; tile sharding splits the rays)'] = dict(ms=timed(r, b, 10), rays=int(b['ray_o'].shape[1])); r.engine.close()
b = scene.make_batch(1024, 1024, seed=0, n_env=0)
r = Renderer(scene.SyntheticNet(sd_r, True), mode='relight', precision='tc', max_rays=b['ray_o'].shape[1] + 8, test_light=('main',), sync_timing=False)
res['5: relight 1024x1024 frame (one GPU; the 100-frame sequence is frame-sharded)'] = dict(ms=timed(r, b, 5), rays=int(b['ray_o'].shape[1])); r.engine.close()
print(json.dumps(res))

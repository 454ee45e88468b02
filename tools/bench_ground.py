"""Row f2 measurement: one 512x512 relit frame WITH ground-plane shading (floor pass over all 262 144 pixels, 16-iteration
env_lvis soft shadows, blend) + n novel env-maps, through Renderer.render.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer

H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b = scene.make_batch(H, H, seed=0, n_env=n_env)
sd = scene.make_state_dict(0, True, True)
for k in ('mask_at_box',):
    b[k] = torch.as_tensor(b[k]).cuda()
P = b['ray_o'].shape[1]
res = {}
for ground in (False, True):
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device='cuda:0', precision='tc', max_rays=P + 8, test_light=('main', 'all'),
                 sync_timing=False, ground_shading=ground)
    for _ in range(2):
        bb = dict(b); bb['mask_at_box'] = b['mask_at_box'].clone()
        r.render(bb)
    torch.cuda.synchronize()
    l0 = r.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        bb = dict(b); bb['mask_at_box'] = b['mask_at_box'].clone()
        out = r.render(bb)
    e1.record(); torch.cuda.synchronize()
    res['ground' if ground else 'plain'] = dict(ms_per_frame=e0.elapsed_time(e1) / n, launches_per_frame=(r.engine.launch_count() - l0) / n,
                                                stats=r.engine.stats())
    r.engine.close()
print(json.dumps({'what': f'{H}x{H} relit frame + {n_env} novel env-maps, with and without ground-plane shading (row f2)', **res}))

"""SURVEY.md 8 f4 measurement: env-map rotation sweep (vis_rotate_light) -- 128 rotations of one probe re-shaded from one
visibility pass at 512x512.  Prints one JSON line (env-maps/s, ms per env-map, HBM bytes moved per env-map)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer

b = scene.make_batch(512, 512, seed=0, n_env=1)
sd = scene.make_state_dict(0, True, True)
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', precision='tc', max_rays=b['ray_o'].shape[1] + 8, test_light=('main',), sync_timing=False)
r.render(b)
eng = r.engine
P = b['ray_o'].shape[1]
st = eng.stats()
probe = torch.as_tensor(next(iter(b['novel_lights'].values()))[0])
n_rot, repeat = 128, 4
for _ in range(3):
    rot = eng.rotate_probes(probe, repeat, 0, n_rot)
    out = eng.relight_envmaps(rot, P)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K):
    rot = eng.rotate_probes(probe, repeat, 0, n_rot)
    out = eng.relight_envmaps(rot, P)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
S = st['n_fg']
read = S * 512 * 8 / 4 + S * 40           # visibility + cosine maps once per 4 probes, per-pixel attributes
write = P * 3 * 4 * 3 * 2                 # zero fill + rgb/shade/spec
print(json.dumps({'metric': 'env-map rotation sweep (f4)', 'env_maps_per_s': n_rot / (ms / 1e3), 'ms_per_env_map': ms / n_rot,
                  'P_rays': P, 'S_fg': S, 'algorithmic_bytes_per_env_map': read + write,
                  'achieved_GBps': (read + write) / (ms / n_rot / 1e3) / 1e9}))

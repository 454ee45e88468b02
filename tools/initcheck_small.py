"""Debug: one small tensor-core-mode frame with floor pass, meant for `compute-sanitizer --tool initcheck`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer
H = int(sys.argv[1]) if len(sys.argv) > 1 else 24
b = scene.make_batch(H, H, seed=0, n_env=0)
sd = scene.make_state_dict(0, relight=True, fitted=True)
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device='cuda:0', precision='tc', max_rays=2048, test_light=('main',),
             return_lvis=True, ground_shading=len(sys.argv) > 2, sync_timing=False)
out = r.render(dict(b))
torch.cuda.synchronize()
print('ok', float(out['main']['rgb_map'].sum()))

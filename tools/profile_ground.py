"""One 512x512 relit frame with ground-plane shading (row f2) for profilers: 1 warm-up frame, then 1 frame.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ground_launches.csv python tools/profile_ground.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer
H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
b = scene.make_batch(H, H, seed=0, n_env=0)
sd = scene.make_state_dict(0, True, True)
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device='cuda:0', precision='tc', max_rays=b['ray_o'].shape[1] + 8, test_light=('main',),
             sync_timing=False, ground_shading=True)
for _ in range(2):
    r.render(dict(b))
torch.cuda.synchronize()
print(r.engine.stats())

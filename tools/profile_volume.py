"""One 512x512 AniSDF volume-rendered frame (BASELINE configs[1]: 128 samples per ray) for profilers: 1 warm-up frame, then 1 frame.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/volume_launches.csv python tools/profile_volume.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer
H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
b = scene.make_batch(H, H, seed=0, n_env=0)
sd = scene.make_state_dict(0, relight=False, fitted=True)
r = Renderer(scene.SyntheticNet(sd, False), mode='anisdf_volume', precision='tc', max_rays=b['ray_o'].shape[1] + 8, sync_timing=False)
for _ in range(2):
    r.render(b)
torch.cuda.synchronize()
print(r.engine.stats())

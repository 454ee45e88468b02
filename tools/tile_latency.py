"""Where a tile-sharded frame's time goes on ONE of W ranks (BASELINE configs[3] strong scaling), measured on a single GPU:
renders the rays rank 0 of W would own (interleaved 32-ray blocks of the 512x512 frame, main + 8 env-maps) and reports the device
time per call, the host time per call (calls issued back to back without synchronising) and the stage split."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import parallel, scene
from relightableavatar_b200.renderer import Renderer

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device('cuda:0')
sd = scene.make_state_dict(0, relight=True, fitted=True)
b = scene.make_batch(512, 512, frame=0, n_frames=8, seed=0, n_env=8)
bd = {k: torch.from_numpy(v).to(dev) for k, v in b.items() if hasattr(v, 'ndim') and getattr(v, 'ndim', 0) > 0 and k != 'novel_lights'}
bd['novel_lights'] = {n: torch.from_numpy(p).to(dev) for n, p in b['novel_lights'].items()}
P = b['ray_o'].shape[1]
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device=dev, precision='tc', max_rays=P + 1024, test_light=('main', 'all'), sync_timing=False)
eng = r.engine
res = {}
for world in (1, W):
    def step():
        if world == 1:
            return r.render(bd)
        local, own = parallel.shard_batch_rays(bd, 0, world)
        eng.set_ray_layout(P, parallel.BLOCK, world, 0)
        out = r.render(local)
        eng.set_ray_layout(0, parallel.BLOCK, 1, 0)
        return out
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    eng.profile_enable(True)
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        step()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    pr = eng.profile_read(); eng.profile_enable(False)
    res[f'world_{world}'] = dict(device_ms=e0.elapsed_time(e1) / n, host_issue_ms=1e3 * (t1 - t0) / n, mlp_ms=pr['mlp_ms'] / n, mlp_launches=pr['mlp_launches'] / n,
                                 stage_ms={k: v / n for k, v in pr['stage_ms'].items()}, rays=int(P if world == 1 else parallel.tile_partition(P, 0, world).numel()))
print(json.dumps(res))

"""Experiment: frames/s with ONE frame in flight vs TWO (two handles, two streams, frames alternate) on one GPU.
The latency-bound stages of a frame (surface tracing: 17 dependent iterations; the attribute GEMM chain) leave most SMs idle;
a second, independent frame can use them."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer

dev = torch.device('cuda:0')
sd = scene.make_state_dict(0, relight=True, fitted=True)
net = scene.SyntheticNet(sd, True)
frames = []
for f in range(4):
    b = scene.make_batch(512, 512, frame=f, n_frames=8, seed=0, n_env=0)
    frames.append({k: torch.from_numpy(v).to(dev) for k, v in b.items() if hasattr(v, 'ndim') and getattr(v, 'ndim', 0) > 0 and k != 'novel_lights'})
P = max(f['ray_o'].shape[1] for f in frames)
res = {}
for n_flight in (1, 2, 3):
    rs = [Renderer(net, mode='relight', device=dev, precision='tc', max_rays=P + 1024, test_light=('main',), sync_timing=False) for _ in range(n_flight)]
    streams = [torch.cuda.Stream(dev) for _ in range(n_flight)]
    def run(n):
        for i in range(n):
            k = i % n_flight
            with torch.cuda.stream(streams[k]):
                rs[k].render(frames[i % len(frames)])
    run(2 * n_flight + 2)
    torch.cuda.synchronize()
    n = 24
    t0 = time.perf_counter()
    run(n)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t1 = time.perf_counter() - t0
    res[f'in_flight_{n_flight}'] = dict(frames_per_s=n / t1, ms_per_frame=1e3 * t1 / n, host_issue_ms_per_frame=1e3 * t_issue / n)
    for r in rs:
        r.engine.close()
print(json.dumps(res))

"""Fused distance-MLP kernel alone: ms per launch on a 2 M-point work list for the given RA_TC_VARIANTs (CUDA events around the
kernel, ra_profile_*), plus the max / mean difference of the distances between the variants.
    python tools/mlp_microbench.py 6 7"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Engine, default_config

variants = sys.argv[1:] or ['6', '7']
b = scene.make_batch(64, 64, seed=0, n_env=0)
sd = scene.make_state_dict(0, True, True)
g = torch.Generator().manual_seed(0)
wv = torch.as_tensor(b['wverts'][0])
n = 2_000_000
x = (wv[torch.randint(0, wv.shape[0], (n,), generator=g)] + torch.randn(n, 3, generator=g) * 0.03).float().cuda()
outs = {}
for v in variants:
    os.environ['RA_TC_VARIANT'] = v.split(':')[0]
    os.environ['RA_TC_SKEW'] = v.split(':')[1] if ':' in v else '4'
    eng = Engine(default_config(True, precision=1, max_rays=16384), 'cuda:0')
    eng.upload_weights(sd); eng.set_frame(b)
    for _ in range(3):
        out = eng.query_sdf(x, 0.125, True)
    torch.cuda.synchronize()
    eng.profile_enable(True)
    for _ in range(10):
        out = eng.query_sdf(x, 0.125, True)
    p = eng.profile_read()
    st = eng.stats()
    outs[v] = out.clone()
    tf = st['n_queries_in_shell'] * 2192384 / (p['mlp_ms'] / p['mlp_launches'] * 1e-3) / 1e12
    print(f'variant {v}: {p["mlp_ms"] / p["mlp_launches"]:.3f} ms per launch over {st["n_queries_in_shell"]} in-shell rows = {tf:.0f} TFLOP/s (algorithmic)', flush=True)
    eng.close()
v0 = variants[0]
for v in variants[1:]:
    d = (outs[v] - outs[v0]).abs()
    print(f'variant {v} vs {v0}: max {float(d.max()):.3e} mean {float(d.mean()):.3e}')

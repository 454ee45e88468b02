"""Fit the synthetic scene's SDF MLP to the body's big-pose capsule-union SDF (CPU, minutes).

No trained checkpoint of the reference exists offline.  With the raw geometric
init the network's zero-set is a 0.5 m sphere, so rays would terminate on the SMPL proxy distance
and never exercise the learned-surface branch of the hierarchical distance query.  This script
trains ONLY `signed_distance_network.mlp.lin*` (weight_v / weight_g / bias, reference key names) for
a fixed number of Adam steps from the seeded geometric init and writes
`relightableavatar_b200/data/sdf_fit_seed0.npz` (float32).  The result is a fixture: it is
committed, and `scene.make_state_dict(fitted=True)` loads it.

    python tools/fit_synthetic_sdf.py [--steps 2500]
"""
import argparse
import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relightableavatar_b200 import scene  # noqa: E402


def pe(x, L):
    freqs = 2.0 ** torch.arange(L, dtype=torch.float32)
    xf = x[..., None, :] * freqs[:, None]
    enc = torch.stack([torch.sin(xf), torch.cos(xf)], dim=-2)
    return torch.cat([x, enc.reshape(*x.shape[:-1], L * 6)], dim=-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=2500)
    ap.add_argument('--batch', type=int, default=8192)
    ap.add_argument('--seed', type=int, default=0)
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    torch.set_num_threads(os.cpu_count())
    body = scene.make_body(args.seed)
    sd = scene.make_state_dict(args.seed, relight=False, fitted=False)
    params = {}
    for l in range(9):
        for s in ('weight_v', 'weight_g', 'bias'):
            k = f'signed_distance_network.mlp.lin{l}.{s}'
            params[k] = sd[k].clone().requires_grad_(True)

    segA = torch.tensor(np.array([s[1] for s in body.big_segs]), dtype=torch.float32)
    segB = torch.tensor(np.array([s[2] for s in body.big_segs]), dtype=torch.float32)
    segR = torch.tensor(np.array([s[3] for s in body.big_segs]), dtype=torch.float32)

    def target(x):
        ab = segB - segA
        t = (((x[:, None] - segA) * ab).sum(-1) / (ab * ab).sum(-1)).clamp(0, 1)
        cp = segA + t[..., None] * ab
        return ((x[:, None] - cp).norm(dim=-1) - segR).min(-1)[0]

    def net(x):
        inp = pe(x, 8)
        h = inp
        for l in range(9):
            v, g, b = (params[f'signed_distance_network.mlp.lin{l}.{s}'] for s in ('weight_v', 'weight_g', 'bias'))
            w = g * v / v.norm(dim=1, keepdim=True)
            if l == 4:
                h = torch.cat([h, inp], -1) / math.sqrt(2)
            h = F.linear(h, w, b)
            if l < 8:
                h = F.softplus(h, beta=100)
        return h[:, 0]

    tv = torch.tensor(body.tverts, dtype=torch.float32)
    lo, hi = tv.min(0)[0] - 0.3, tv.max(0)[0] + 0.3
    opt = torch.optim.Adam(list(params.values()), lr=5e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, args.steps, eta_min=2e-5)
    g = torch.Generator().manual_seed(args.seed + 1)
    t0 = time.time()
    for it in range(args.steps):
        n = args.batch
        idx = torch.randint(0, tv.shape[0], (n,), generator=g)
        near = tv[idx[: n // 2]] + torch.randn(n // 2, 3, generator=g) * 0.02
        mid = tv[idx[n // 2: n * 7 // 8]] + torch.randn(n * 7 // 8 - n // 2, 3, generator=g) * 0.10
        uni = lo + (hi - lo) * torch.rand(n - n * 7 // 8, 3, generator=g)
        x = torch.cat([near, mid, uni])
        y = target(x)
        p = net(x)
        wgt = 1.0 / (0.02 + y.abs())
        loss = (wgt * (p - y).abs()).sum() / wgt.sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step(); sched.step()
        if it % 100 == 0 or it == args.steps - 1:
            with torch.no_grad():
                e = (p - y).abs()
                print(f'it {it:5d} loss {loss.item():.5f} | near mean {e[: n // 2].mean():.5f} max {e[: n // 2].max():.4f} | '
                      f'all mean {e.mean():.5f} | {time.time() - t0:.0f}s', flush=True)
    out = os.path.join(os.path.dirname(os.path.abspath(scene.__file__)), 'data', f'sdf_fit_seed{args.seed}.npz')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez(out, **{k: v.detach().numpy().astype(np.float32) for k, v in params.items()})
    print('wrote', out, os.path.getsize(out) / 1e6, 'MB')


if __name__ == '__main__':
    main()

"""Shadow-ray packets: human frame and floor pass at 512^2 under combinations of RA_PKT_ORDER (rays generated as same-light /
32-neighbouring-pixel packets) and RA_PKT_SEARCH (far-field 3-NN per packet); each a bit mask (bit 0: floor pass, bit 1: human pass),
given as two digits per combination, e.g. `pkt_sweep.py 512 00 11 33`.  One JSON line per combination."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer

H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
combos = [tuple(int(c) for c in a) for a in sys.argv[2:]] or [(0, 0), (1, 1), (3, 3)]
b = scene.make_batch(H, H, seed=0, n_env=0)
sd = scene.make_state_dict(0, True, True)
b['mask_at_box'] = torch.as_tensor(b['mask_at_box']).cuda()
P = b['ray_o'].shape[1]
for order, search in combos:
    os.environ['RA_PKT_ORDER'] = str(order); os.environ['RA_PKT_SEARCH'] = str(search)
    res = dict(order=order, search=search)
    for ground in (False, True):
        r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device='cuda:0', precision='tc', max_rays=P + 8, test_light=('main',),
                     sync_timing=False, ground_shading=ground)
        for _ in range(2):
            bb = dict(b); bb['mask_at_box'] = b['mask_at_box'].clone()
            r.render(bb)
        torch.cuda.synchronize()
        r.engine.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5 if ground else 10
        e0.record()
        for _ in range(n):
            bb = dict(b); bb['mask_at_box'] = b['mask_at_box'].clone()
            r.render(bb)
        e1.record(); torch.cuda.synchronize()
        pr = r.engine.profile_read()
        mlp_ms, stage = pr['mlp_ms'], list(pr['stage_ms'].values())
        st = r.engine.stats()
        res['ground' if ground else 'plain'] = dict(ms=round(e0.elapsed_time(e1) / n, 3), mlp_ms=round(mlp_ms / n, 3), stages=[round(x / n, 3) for x in stage],
                                                    rays=st['n_shadow_rays'], slots=st['n_shadow_slots'], queries=st['n_queries'])
        r.engine.close()
    print(json.dumps(res), flush=True)

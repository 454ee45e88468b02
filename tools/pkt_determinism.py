"""Debug: is a frame bit-identical run to run / across shadow-ray orders?  Prints the differing entries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer
H = int(sys.argv[1]) if len(sys.argv) > 1 else 96
ground = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
b = scene.make_batch(H, H, seed=0, n_env=0)
sd = scene.make_state_dict(0, relight=True, fitted=True)
outs = []
for order, search in ((0, 0), (0, 0), (1, 0), (1, 0), (1, 1), (1, 1)):
    os.environ['RA_PKT_ORDER'] = str(order); os.environ['RA_PKT_SEARCH'] = str(search)
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', device='cuda:0', precision=os.environ.get('PREC', 'tc'), max_rays=max(16384, H * H // 3), test_light=('main',),
                 return_lvis=True, ground_shading=ground, sync_timing=False)
    out = r.render(dict(b))
    outs.append(((order, search), out['main']['lvis_map'].clone(), r.engine.stats()))
    r.engine.close()
ref = outs[0][1]
for key, lv, st in outs[1:]:
    d = (lv != ref).nonzero()
    print(key, 'differing entries', d.shape[0], 'queries', st['n_queries'], 'in-shell', st['n_queries_in_shell'])
    for row in d[:6].tolist():
        print('   at', row, float(lv[tuple(row)]), float(ref[tuple(row)]))

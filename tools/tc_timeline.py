"""Debug: per-layer clock timeline of CTA 0 (second tile) of k_mlp_tc on a large query batch."""
import ctypes, os, sys
os.environ['RA_TC_VARIANT'] = '1'      # this tool reads the single-CTA kernel's 18 x 8 timeline (tools/tc6_timeline.py: the pair kernel)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relightableavatar_b200 import scene, _lib
from relightableavatar_b200.renderer import Engine, default_config
b = scene.make_batch(64, 64, seed=0, n_env=0)
sd = scene.make_state_dict(0, True, True)
eng = Engine(default_config(True, precision=1, max_rays=16384), 'cuda:0')
eng.upload_weights(sd); eng.set_frame(b)
lib = _lib.load()
lib.ra_debug_tc_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
g = torch.Generator().manual_seed(0)
wv = torch.as_tensor(b['wverts'][0])
n = 2_000_000
x = (wv[torch.randint(0, wv.shape[0], (n,), generator=g)] + torch.randn(n, 3, generator=g) * 0.03).float().cuda()
eng.query_sdf(x, 0.125, True); torch.cuda.synchronize()
lib.ra_debug_tc_timeline(eng.h, None, 1)
eng.query_sdf(x, 0.125, True); torch.cuda.synchronize()
arr = (ctypes.c_ulonglong * (18 * 8))()
lib.ra_debug_tc_timeline(eng.h, arr, 0)
t0 = arr[0]
print('layer | mma: wait_act  issue(+wait weights) | epi: wait_acc  work  arrive | layer period')
prev = None
for l in range(18):
    a = [arr[l * 8 + k] for k in range(7)]
    per = (a[0] - prev) if prev else 0
    prev = a[0]
    print(f'{l:2d}   | {a[1]-a[0]:6d} {a[2]-a[1]:6d} | {a[4]-a[3]:6d} {a[5]-a[4]:6d} {a[6]-a[5]:6d} | {per:6d}   (mma_done->epi_start {a[4]-a[2]:6d})')

"""Debug: per-(layer, half, slot) clock timeline of cluster 0 / leader CTA (second quad) of k_mlp_tc8.
Needs the instrumented build: python -m relightableavatar_b200.build --timeline ; RA_LIB_PATH=.../libra_b200_tl.so"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['RA_TC_VARIANT'] = '8'
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
arr = (ctypes.c_ulonglong * 512)()
lib.ra_debug_tc_timeline(eng.h, arr, 0)
print('layer half slot | mma: wait_dep  issue | epi: wait_acc  work | mma period | acc period')
pm = pe = None
for l in range(18):
    for h in range(2):
        for p in range(2):
            ti = (l * 2 + h) * 2 + p
            m = [arr[256 + ti * 3 + k] for k in range(3)]
            e = [arr[ti * 3 + k] for k in range(3)]
            if not m[0] and not e[0]:
                continue
            print(f'{l:2d} {h} {p} | {m[1]-m[0]:6d} {m[2]-m[1]:6d} | {e[1]-e[0]:6d} {e[2]-e[1] if e[2] else 0:6d} | {(m[0]-pm) if pm else 0:6d} | {(e[1]-pe) if pe else 0:6d}')
            pm, pe = m[0], e[1]

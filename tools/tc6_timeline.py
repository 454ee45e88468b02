"""Debug: per-(layer, slot) clock timeline of cluster 0 / leader CTA (second quad) of k_mlp_tc6 on a large query batch.
Needs the instrumented build: python -m relightableavatar_b200.build --timeline ; RA_LIB_PATH=.../libra_b200_tl.so RA_TC_VARIANT=6"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('RA_TC_VARIANT', '6')
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.query_sdf(x, 0.125, True); e1.record(); torch.cuda.synchronize()
print(f'query_sdf({n}) {e0.elapsed_time(e1):.3f} ms (front-end + MLP kernel)')
lib.ra_debug_tc_timeline(eng.h, None, 1)
eng.query_sdf(x, 0.125, True); torch.cuda.synchronize()
arr = (ctypes.c_ulonglong * 512)()
lib.ra_debug_tc_timeline(eng.h, arr, 0)
print('layer slot | mma: wait_act  issue  (of which waiting on weights) | epi: wait_acc  work  arrive | period | acc_commit->epi_wake | epi_arrive->mma_wake(next)')
prev = None
for i in range(36):
    a = [arr[i * 8 + k] for k in range(8)]
    per = (a[0] - prev) if prev else 0
    prev = a[0]
    nxt = arr[(i + 2) * 8 + 1] - a[6] if i + 2 < 36 and a[6] else 0
    print(f'{i//2:2d} {i%2} | {a[1]-a[0]:6d} {a[2]-a[1]:6d} ({a[7]:6d}) | {a[4]-a[3]:6d} {a[5]-a[4]:6d} {a[6]-a[5] if a[6] else 0:6d} | {per:6d} | {a[4]-a[2]:6d} | {nxt:6d}')
print('layer 2, per chunk: wait_full | issue 2 MMAs | commit | gap to next chunk')
for p in range(2):
    for c in range(9):
        d = [arr[288 + (p * 9 + c) * 4 + k] for k in range(4)]
        nx = arr[288 + (p * 9 + c + 1) * 4] if c < 8 else d[3]
        print(f'  slot {p} chunk {c}: {d[1]-d[0]:5d} | {d[2]-d[1]:5d} | {d[3]-d[2]:5d} | {nx-d[3]:5d}')

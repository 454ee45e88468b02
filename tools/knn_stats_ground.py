"""Debug: KNN path statistics of the floor pass (needs the -DRA_KNN_STATS build: libra_b200_stats.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relightableavatar_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libra_b200_stats.so')
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer
b = scene.make_batch(512, 512, seed=0, n_env=0)
sd = scene.make_state_dict(0, True, True)
lib = _lib.load()
arr = (ctypes.c_ulonglong * 12)()
for ground in (False, True):
    r = Renderer(scene.SyntheticNet(sd, True), mode='relight', precision='tc', max_rays=80000, sync_timing=False, ground_shading=ground)
    r.render(dict(b)); lib.ra_debug_knn_stats(arr, 1)
    r.render(dict(b)); lib.ra_debug_knn_stats(arr, 1)
    n = arr[0] + arr[1]
    print('ground' if ground else 'plain', ': near-finished', arr[0], 'far-phase', arr[1], 'far cells/query', arr[2] / max(arr[1], 1), 'far verts/query', arr[3] / max(arr[1], 1),
          '| near phase per query: verts', arr[4] / max(n, 1), 'levels', arr[5] / max(n, 1), r.engine.stats()['n_queries'],
          '| packets', arr[6], 'refused', arr[7], 'lanes/packet', arr[8] / max(arr[6], 1), 'cells/packet', arr[9] / max(arr[6], 1), 'verts/packet', arr[10] / max(arr[6], 1))
    r.engine.close()

"""Debug: KNN path statistics per render stage (needs the -DRA_KNN_STATS build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relightableavatar_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libra_b200_stats.so')
import torch
from relightableavatar_b200 import scene
from relightableavatar_b200.renderer import Renderer
b = scene.make_batch(512, 512, seed=0, n_env=0)
sd = scene.make_state_dict(0, True, True)
r = Renderer(scene.SyntheticNet(sd, True), mode='relight', precision='tc', max_rays=80000, sync_timing=False)
lib = _lib.load()
arr = (ctypes.c_ulonglong * 12)()
r.render(b); lib.ra_debug_knn_stats(arr, 1)
r.render(b); lib.ra_debug_knn_stats(arr, 1)
n = arr[0] + arr[1]
print('frame: near-finished', arr[0], 'far-phase', arr[1], 'far cells/query', arr[2] / max(arr[1], 1), 'far verts/query', arr[3] / max(arr[1], 1),
      '| near phase per query: verts', arr[4] / max(n, 1), 'rings', arr[5] / max(n, 1), 'row scans', arr[6] / max(n, 1), r.engine.stats())

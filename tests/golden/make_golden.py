"""Generate the golden vectors in this directory by running the UNMODIFIED reference renderer
(under oracle/ref_harness.py's import shims) on the seeded synthetic scene.

Run in the dev container where /root/reference exists (the GPU box has no reference tree):

    python tests/golden/make_golden.py [case names]

Each fixture stores only the reference OUTPUTS plus the scene arguments; inputs are regenerated
from the seed by relightableavatar_b200/scene.py (deterministic numpy / torch CPU generators).
One subprocess per config: the reference binds hot-path parameters at import time.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = [
    # name,            mode,            H,  n_env, keys kept
    ('relight_48', 'relight', 48, 2, None),
    ('anisdf_trace_48', 'anisdf_trace', 48, 0, None),
    ('anisdf_volume_24', 'anisdf_volume', 24, 0, None),
    # row f2: vis_ground_shading on (image-sized maps; the CPU harness scrambles `inds`, see oracle.render_ground_pass)
    ('relight_ground_24', 'relight_ground', 24, 1, None),
    # row f1: the reference's dataset-side numpy / torch functions (rays, AABB, LBS, bounds)
    ('prep_24', 'prep', 24, 0, None),
]

DROP_DUP = ('surf_map', 'depth_map', 'acc_map', 'albedo_map', 'roughness_map', 'norm_map', 'ray_o', 'cpts_map',
            'bpts_map', 'envmap.probe')


def main():
    only = set(sys.argv[1:])
    for name, mode, H, n_env, _ in CASES:
        if only and name not in only:
            continue
        tmp = os.path.join('/tmp', f'golden_{name}.npz')
        subprocess.check_call([sys.executable, os.path.join(ROOT, 'oracle', 'ref_harness.py'), '--mode', mode,
                               '--H', str(H), '--n_env', str(n_env), '--out', tmp])
        d = dict(np.load(tmp))
        keep = {}
        seen_lvis = False
        for k, v in d.items():
            light, _, key = k.partition('.')
            if mode == 'relight_ground' and light not in ('main', 'wbounds_after') and key not in ('rgb_map', 'shade_map', 'spec_map', 'albedo_map'):
                continue                 # per-light copies of the blended main maps
            if mode == 'relight' and light not in ('main', 'wbounds_after'):
                if key in ('lvis_map', 'ldot_map'):
                    if seen_lvis and light != first_light:
                        continue
                    first_light = light
                    keep[key] = v.astype(np.float32)
                    continue
                if key in DROP_DUP:      # per-light copies of the main maps (novel_light_sphere_tracing.py:189)
                    continue
            keep[k] = v
            if 'lvis_map' in keep and 'ldot_map' in keep:
                seen_lvis = True
        keep['_H'] = np.int64(H); keep['_n_env'] = np.int64(n_env); keep['_seed'] = np.int64(0)
        out = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(out, **keep)
        print(name, os.path.getsize(out) / 1e3, 'kB', sorted(keep))


if __name__ == '__main__':
    main()

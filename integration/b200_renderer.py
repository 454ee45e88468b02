"""The one file a reference maintainer adds: `lib/networks/renderer/b200_renderer.py`.

The reference selects its renderer by module name,

    renderer = importlib.import_module(cfg.renderer_module).Renderer(network)        lib/networks/renderer/make_renderer.py:5-8

and only ever calls `renderer.render(batch)` (run.py:68-85).  This module is that `Renderer`: it reads the reference's global
`cfg` (the same keys the reference's own renderers bind at import time) and forwards to the CUDA library.  Selected like any
other renderer:

    python run.py -t visualize -c configs/mobile_stage/xuzhen_12v_geo.yaml relighting True vis_novel_light True \
           vis_pose_sequence True renderer_module lib.networks.renderer.b200_renderer

tests/test_dropin_reference.py loads this file under exactly that module name next to the unmodified reference tree and runs
both renderers on the same network object and batch.
"""
import os

from lib.config import cfg                                   # the reference's own config object
from relightableavatar_b200.renderer import Renderer as _B200Renderer


class Renderer(_B200Renderer):
    def __init__(self, net):
        mode = ('relight' if cfg.relighting else
                'anisdf_trace' if cfg.vis_sphere_tracing else 'anisdf_volume')
        dev = next(net.parameters()).device
        super().__init__(net, mode=mode, cfg=cfg, device=dev,      # cfg: incl. test_light, vis_ground_shading, ground_*, env_lvis, vis_rotate_light
                         precision=os.environ.get('RA_B200_PRECISION', 'tc'),
                         to_cpu=True,                               # novel_light_sphere_tracing.py:216 hands CPU maps to the visualizer
                         # the reference's per-light dicts also carry the two (P,512) visibility maps (141 MB each at 512^2) that
                         # nothing downstream reads; RA_B200_FULL_KEYS=1 reproduces that, the default leaves them out
                         return_lvis=os.environ.get('RA_B200_FULL_KEYS', '0') == '1',
                         max_rays=max(int(cfg.H), 512) * max(int(cfg.W), 512))      # grown on demand by the first larger frame

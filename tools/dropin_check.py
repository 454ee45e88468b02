"""Drop-in check under the reference's OWN objects (row b): reference `cfg`, reference `make_network` module, reference dotdict
batch, reference `make_renderer` selecting `cfg.renderer_module` -- once the stock novel-light renderer, once the binding
integration/b200_renderer.py registered as lib.networks.renderer.b200_renderer -- rendered on the same GPU and compared.

    python tools/dropin_check.py [--H 64] [--n_env 2] [--precision fp32] [--ground]

Needs a reference tree (RA_REFERENCE, /root/reference or baseline/_ref: oracle/install_reference.py) and a GPU.  Prints one JSON
line `DROPIN {...}`; tests/test_dropin_reference.py asserts on it.  Test infrastructure: never imported by the product.
"""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--H', type=int, default=64)
    ap.add_argument('--n_env', type=int, default=2)
    ap.add_argument('--precision', default='fp32')
    ap.add_argument('--ground', action='store_true')
    a = ap.parse_args()
    os.environ['RA_B200_PRECISION'] = a.precision
    os.environ['RA_B200_FULL_KEYS'] = '1'
    import numpy as np
    import torch
    from oracle import ref_harness as RH
    from oracle import ra_oracle as O
    from relightableavatar_b200 import scene
    cfg = RH.setup_reference('relight_ground' if a.ground else 'relight')          # chdir into the reference tree, import shims, cfg cascade
    names = list(scene.make_envmaps(a.n_env, 10).keys())
    cfg.test_light = ['main'] + names
    from lib.networks.make_network import make_network
    from lib.networks.renderer.make_renderer import make_renderer
    from lib.utils.base_utils import dotdict
    dev = os.environ.get('RA_DROPIN_DEVICE', 'cuda')          # (a CPU dry run of the reference half only: the plugin itself has no CPU path)
    net = make_network(cfg)
    sd = scene.make_state_dict(0, relight=True, fitted=True)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    net = net.to(dev).eval()
    b = scene.make_batch(a.H, a.H, seed=0, n_env=a.n_env)

    def fresh_batch():           # the reference grows batch.wbounds in place: every renderer gets its own copy
        return RH.to_ref_batch(b, device=dev)

    stock_module = cfg.renderer_module
    torch.manual_seed(0)
    with torch.no_grad():
        ref = make_renderer(cfg, net).render(fresh_batch())
    # the binding, under the module name a maintainer would give it
    name = 'lib.networks.renderer.b200_renderer'
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'integration', 'b200_renderer.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    cfg.renderer_module = name
    renderer = make_renderer(cfg, net)
    assert type(renderer).__module__ == name
    batch = fresh_batch()
    with torch.no_grad():
        ours = renderer.render(batch)
    torch.cuda.synchronize()

    res = dict(stock_module=stock_module, H=a.H, precision=a.precision, ground=a.ground, lights=sorted(k for k in ref if k != 'diff'),
               so_loaded=any('libra_b200' in l for l in open('/proc/self/maps').read().splitlines()))
    res['same_lights'] = sorted(ours.keys()) == sorted(ref.keys())
    res['diff_is_float'] = isinstance(ours['diff'], float) and isinstance(ref['diff'], float)
    res['return_type'] = type(ours).__name__
    layout = {}
    psnr = {}
    maxerr = {}
    mask_b = {'mask_at_box': (np.ones_like(b['mask_at_box']) if a.ground else b['mask_at_box'])}
    for light in res['lights']:
        ro, oo = ref[light], ours[light]
        kr = sorted(k for k in ro.keys()); ko = sorted(k for k in oo.keys())
        lay = dict(same_keys=kr == ko, missing=[k for k in kr if k not in ko], extra=[k for k in ko if k not in kr])
        bad = []
        for k in kr:
            if k not in oo:
                continue
            r, o = ro[k], oo[k]
            if isinstance(r, dict):
                r, o = r['probe'], o['probe']
            if tuple(r.shape) != tuple(o.shape) or r.dtype != o.dtype or r.device.type != o.device.type:
                bad.append((k, tuple(r.shape), tuple(o.shape), str(r.dtype), str(o.dtype), r.device.type, o.device.type))
        lay['mismatched'] = bad
        layout[light] = lay
        img_r = O.assemble_image(mask_b, ro['rgb_map'][0].float().cpu())
        img_o = O.assemble_image(mask_b, oo['rgb_map'][0].float().cpu())
        psnr[light] = O.psnr(img_o, img_r)
        fg = (ro['acc_map'][0] > 0).cpu() & (oo['acc_map'][0] > 0).cpu()
        errs = {}
        for k in ('rgb_map', 'acc_map', 'surf_map', 'albedo_map', 'roughness_map', 'shade_map', 'depth_map', 'spec_map'):
            if k in ro and k in oo:
                e = (ro[k][0].float().cpu() - oo[k][0].float().cpu()).abs()
                errs[k] = float(torch.quantile(e.flatten()[:: max(1, e.numel() // 500000)], 0.99))
        maxerr[light] = errs
        res.setdefault('fg_flips', {})[light] = int(((ro['acc_map'][0] > 0).cpu() != (oo['acc_map'][0] > 0).cpu()).sum())
    res.update(layout=layout, psnr=psnr, q99=maxerr, ref_diff_s=ref['diff'], ours_diff_s=ours['diff'],
               probe_equal=bool(torch.equal(ref['main']['envmap']['probe'].cpu(), ours['main']['envmap']['probe'].cpu())) if 'main' in ref else None)
    print('DROPIN ' + json.dumps(res), flush=True)


if __name__ == '__main__':
    main()

"""-m gpu: the drop-in boundary for real (row b).  The reference's own `cfg`, `make_network` module, dotdict batch and
`make_renderer` (lib/networks/renderer/make_renderer.py:5-8) select the binding integration/b200_renderer.py through
`cfg.renderer_module`, exactly as `run.py -t visualize ... renderer_module lib.networks.renderer.b200_renderer` would, and the result
is compared with the reference's stock novel_light_sphere_tracing.Renderer run on the same GPU, network object and batch
(tools/dropin_check.py, one subprocess per case: the reference binds its hot-path parameters at import time).
Skipped when no reference tree travelled with the snapshot (oracle/install_reference.py -> baseline/_ref)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    sys.path.insert(0, ROOT)
    from oracle import ref_harness as RH
    try:
        ref = RH.find_reference()
    except FileNotFoundError:
        pytest.skip('no reference tree (RA_REFERENCE, /root/reference, baseline/_ref)')
    if os.path.basename(ref) == '_ref':
        from oracle import install_reference
        assert install_reference.verify(ref), 'baseline/_ref differs from the reference it was mirrored from'
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'dropin_check.py'), *args], capture_output=True, text=True, timeout=1500)
    line = [l for l in r.stdout.splitlines() if l.startswith('DROPIN ')]
    assert r.returncode == 0 and line, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(line[-1][7:])


def _check_layout(res):
    assert res['so_loaded'], 'libra_b200.so was not loaded by the plugin'
    assert res['return_type'] == 'dotdict' and res['same_lights'] and res['diff_is_float']
    for light, lay in res['layout'].items():
        assert lay['same_keys'], f'{light}: missing {lay["missing"]} extra {lay["extra"]}'
        assert not lay['mismatched'], f'{light}: {lay["mismatched"]}'       # shape, dtype and device (per-light maps on the CPU, main on the GPU)
    assert res['probe_equal']


def test_make_renderer_selects_the_plugin_fp32():
    """64x64, main + 2 novel env-maps, reference-precision mode: same dotdict layout (keys, shapes, dtypes, devices, `diff`,
    `envmap.probe`) and PSNR >= 75 dB against the stock renderer (measured 79-97 dB against the oracle at this size)."""
    res = _run('--H', '64', '--n_env', '2', '--precision', 'fp32')
    _check_layout(res)
    assert res['stock_module'].endswith('novel_light_sphere_tracing')
    for light, p in res['psnr'].items():
        assert p >= 75.0, f'{light}: PSNR {p:.1f} dB vs the stock renderer'
        assert res['fg_flips'][light] <= 2
        for k, e in res['q99'][light].items():
            assert e <= 2e-4, f'{light}.{k}: q99 {e:.3e}'


def test_make_renderer_selects_the_plugin_tc():
    """The product default (tensor-core distance MLPs): same layout, PSNR >= 55 dB (measured 61-67 dB against the oracle)."""
    res = _run('--H', '64', '--n_env', '1', '--precision', 'tc')
    _check_layout(res)
    for light, p in res['psnr'].items():
        assert p >= 55.0, f'{light}: PSNR {p:.1f} dB vs the stock renderer'


def test_make_renderer_selects_the_plugin_ground_shading():
    """The README showcase option (vis_ground_shading, readme.md:64) through the same door: image-sized maps."""
    res = _run('--H', '24', '--n_env', '1', '--precision', 'fp32', '--ground')
    assert res['so_loaded'] and res['same_lights']
    for light, lay in res['layout'].items():
        assert not lay['missing'], f'{light}: missing {lay["missing"]}'
        assert not lay['mismatched'], f'{light}: {lay["mismatched"]}'

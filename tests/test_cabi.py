"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/ra_b200.h declares (CPU only,
no compute calls)."""
import ctypes
import os
import re

from relightableavatar_b200 import _lib, build


def test_library_builds_and_exports_header_symbols():
    path = build.build(verbose=False)
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'ra_b200.h')).read()
    declared = set(re.findall(r'\b(ra_[a-z_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    for s in declared:
        assert hasattr(lib, s), f'{s} declared in ra_b200.h but not exported'
    assert set(_lib.EXPORTS) <= declared


def test_ctypes_struct_sizes_match_header():
    # 39 4-byte fields in ra_config; pointers are 8 bytes
    assert ctypes.sizeof(_lib.ra_config) == 39 * 4
    assert ctypes.sizeof(_lib.ra_frame) == 11 * 8
    assert ctypes.sizeof(_lib.ra_outputs) == 14 * 8
    assert ctypes.sizeof(_lib.ra_stats) == 8 * 8
    assert ctypes.sizeof(_lib.ra_ground_config) == 16 * 4
    assert ctypes.sizeof(_lib.ra_ground_outputs) == 10 * 8
    assert ctypes.sizeof(_lib.ra_body) == 7 * 8          # 6 pointers + int32 padded to 8
    assert ctypes.sizeof(_lib.ra_pose_outputs) == 8 * 8


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'relightableavatar_b200')
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f'{f} imports the oracle'


def _dot(d):
    from relightableavatar_b200.renderer import dotdict
    return dotdict({k: _dot(v) if isinstance(v, dict) else v for k, v in d.items()})


def test_default_config_equals_the_reference_cfg_cascade():
    """renderer.default_config / default_ground_config restate the EFFECTIVE cfg of xuzhen_12v_geo(_fix_mat); the fixture holds
    what the reference's own config cascade yields (tests/golden/make_golden.py cfg_values -> oracle/ref_harness.py dump_cfg)."""
    import json
    import pytest
    from relightableavatar_b200 import renderer as R
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cfg_values.json')
    if not os.path.exists(p):
        pytest.skip('cfg fixture missing')
    vals = json.load(open(p))
    for mode, relight in (('relight', True), ('anisdf_trace', False), ('anisdf_volume', False)):
        cfg = _dot(vals[mode])
        got, want = R.config_from_reference_cfg(cfg, relight, mode), R.default_config(relight)
        assert set(got) == set(want)
        if mode == 'anisdf_volume':      # the volume renderer's ray chunk (8192): no effect on results, ra_render_anisdf_volume
            assert got.pop('render_chunk') == 8192      # streams 8192-ray slabs on its own; the traced modes use 65536
            want.pop('render_chunk')
        for k in want:
            assert got[k] == pytest.approx(want[k], rel=1e-12), (mode, k, got[k], want[k])
        for k, d in R.FIXED_SWITCHES.items():          # the reference's own defaults are the values the library implements
            assert vals[mode][k] == d, (mode, k, vals[mode][k], d)
    cfg = _dot(vals['relight'])
    cfg['no_dfss'] = True                               # an ablation switch the library does not implement: refused, not ignored
    with pytest.raises(NotImplementedError, match='no_dfss'):
        R.config_from_reference_cfg(cfg, True, 'relight')
    cfg = _dot(vals['relight'])
    cfg['sphere_tracing']['tan_i_multiplier'] = 2
    with pytest.raises(NotImplementedError, match='tan_i_multiplier'):
        R.config_from_reference_cfg(cfg, True, 'relight')
    g, gw = R.ground_config_from_reference_cfg(_dot(vals['relight'])), R.default_ground_config()
    for k in gw:
        assert g[k] == pytest.approx(gw[k]), (k, g[k], gw[k])


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/ra_b200.h compiles as pedantic C99, a C host links against the library, and without a B200 ra_create refuses
    loudly (non-zero status + message) instead of falling back to anything."""
    import shutil
    import subprocess
    import pytest
    if shutil.which('gcc') is None:
        pytest.skip('gcc missing')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(build.build(verbose=False))
    exe = str(tmp_path / 'host_minimal')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-I' + os.path.join(root, 'include'),
                           os.path.join(root, 'examples', 'host_minimal.c'), '-L' + libdir, '-lra_b200', '-Wl,-rpath,' + libdir, '-o', exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode in (0, 2), (r.returncode, r.stderr)
    if r.returncode == 2:
        assert 'ra_create failed:' in r.stderr and len(r.stderr.strip()) > len('ra_create failed:')
    else:
        assert 'handle created' in r.stdout

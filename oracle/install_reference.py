"""TEST INFRASTRUCTURE -- places the UNMODIFIED reference where the GPU box can see it: `baseline/_ref/`.

    python oracle/install_reference.py            # in the dev container, where /root/reference exists

`/root/reference` does not exist on the GPU box; `baseline/_ref/` is git-ignored (never part of the history) but travels with the
repo snapshot.  The contract's recipe

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference

fails for this reference ("Neither 'setup.py' nor 'pyproject.toml' found": it is a research tree that is run in place, `python run.py ...`),
so the tree is mirrored file by file instead: `lib/` and `configs/` (1.1 MB, Python + YAML only), byte for byte, plus a
MANIFEST with the sha256 of every file so that a test can tell the copy is unmodified.  Nothing under baseline/_ref is imported by the
product; `oracle/ref_harness.py` (find_reference) is the only reader: the reference arm of bench.py and the drop-in GPU test.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('RA_REFERENCE_SRC', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')
PARTS = ('lib', 'configs', 'run.py')


def sha(path: str) -> str:
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def main() -> int:
    if not os.path.isdir(SRC):
        print(f'{SRC} not present: nothing to install (the GPU box uses the copy that travelled with the snapshot)')
        return 0
    pip = subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--find-links', '/opt/wheelhouse',
                          '--target', '/tmp/_ra_ref_pip_target', SRC], capture_output=True, text=True)
    pip_msg = (pip.stdout + pip.stderr).strip().splitlines()[-1] if (pip.stdout + pip.stderr).strip() else ''
    print('pip install:', 'ok' if pip.returncode == 0 else f'failed ({pip_msg})')
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for part in PARTS:
        s = os.path.join(SRC, part)
        if os.path.isdir(s):
            for dp, dn, fn in os.walk(s):
                dn[:] = [d for d in dn if d != '__pycache__']
                for f in fn:
                    if f.endswith('.pyc'):
                        continue
                    sp = os.path.join(dp, f)
                    rel = os.path.relpath(sp, SRC)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(sp, os.path.join(DST, rel))
                    manifest[rel] = sha(sp)
        elif os.path.isfile(s):
            shutil.copyfile(s, os.path.join(DST, part))
            manifest[part] = sha(s)
    json.dump({'source': SRC, 'pip_install': 'ok' if pip.returncode == 0 else pip_msg, 'files': manifest},
              open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=0, sort_keys=True)
    print(f'mirrored {len(manifest)} files of the unmodified reference into {DST}')
    return 0


def verify(dst: str = DST) -> bool:
    """True when every file under `dst` still has the sha256 recorded at install time (the copy is the unmodified reference)."""
    m = os.path.join(dst, 'MANIFEST.json')
    if not os.path.exists(m):
        return False
    files = json.load(open(m))['files']
    return all(os.path.exists(os.path.join(dst, rel)) and sha(os.path.join(dst, rel)) == h for rel, h in files.items())


if __name__ == '__main__':
    sys.exit(main())

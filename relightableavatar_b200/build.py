"""Builds the in-tree CUDA library `libra_b200.so` (sm_100a only) with nvcc.

    python -m relightableavatar_b200.build

The .so is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only dev container.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'ra_api.cu')
OUT = os.path.join(HERE, 'libra_b200.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))] + \
       [os.path.join(os.path.dirname(HERE), 'include', 'ra_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = True, extra=(), out: str = OUT) -> str:
    """`extra` / `out`: instrumented debug builds (e.g. extra=['-DRA_TC_TIMELINE'], out='.../libra_b200_tl.so')."""
    if not force and out == OUT and not needs_build():
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ['-lcuda', '-ldl', '-o', out, SRC]
    if verbose:
        print(' '.join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout)
    return out


if __name__ == '__main__':
    if '--timeline' in sys.argv:
        print(build(force=True, extra=['-DRA_TC_TIMELINE'], out=os.path.join(HERE, 'libra_b200_tl.so')))
    else:
        print(build(force='--force' in sys.argv))

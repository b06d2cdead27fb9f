"""Build `libcitylearn_b200.so` in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / 'csrc' / 'citylearn_b200.cu'
OUT = PKG / 'libcitylearn_b200.so'
DEPS = [SRC, PKG / 'csrc' / 'unit_physics.cuh', PKG.parent / 'include' / 'citylearn_b200.h']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '--fmad=false',          # the fp64 mode reproduces the reference's (non-fused) NumPy arithmetic bit for bit
    '-shared', '-Xcompiler', '-fPIC',
    # (no --split-compile: splitting the module for parallel optimisation changed the code of every instantiation - the plain fp64 step
    #  kernel flipped between 116 registers / no stack and 128 registers / 168 B of spills from one unrelated edit to the next, and the
    #  spill-free variant it produced was still 2.5 % (fp64) / 5.5 % (fp32) slower on B200 than the single-module build; 1.5 min vs 40 s)
]


def nvcc_path() -> str:
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.isfile(c):
            return c
    raise RuntimeError('nvcc not found')


def is_stale() -> bool:
    if not OUT.is_file():
        return True
    t = OUT.stat().st_mtime
    return any(p.stat().st_mtime > t for p in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return OUT
    cmd = [nvcc_path(), *NVCC_FLAGS, '-o', str(OUT), str(SRC)]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed:\n{res.stdout}\n{res.stderr}')
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force=True, verbose=True))

"""In-tree build of libdiffsbdd_b200.so with nvcc for sm_100a (no torch dependency in the library).

The built .so stays next to this file (git-ignored, but shipped to the GPU box by gpurun)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_NAME = 'libdiffsbdd_b200.so'
LIB_PATH = os.path.join(HERE, LIB_NAME)
INSTR_LIB_PATH = os.path.join(HERE, 'libdiffsbdd_b200_instr.so')   # same sources with -DDSB_TC_INSTRUMENT=1 (profiles/tc_ablate.py)
SOURCES = ['dsb_api.cu', 'dsb_node.cu', 'dsb_edge.cu', 'dsb_tc.cu']
HEADERS = [os.path.join(CSRC, 'dsb_internal.cuh'), os.path.join(CSRC, 'dsb_tc.cuh'), os.path.join(HERE, '..', 'include', 'diffsbdd_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++20',
              '-Xcompiler', '-fPIC', '-Wno-deprecated-gpu-targets']


def _nvcc() -> str:
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: cannot build libdiffsbdd_b200.so')


def _digest() -> str:
    h = hashlib.sha256()
    for p in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(p, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current(instrumented: bool = False) -> bool:
    """True when the library on disk was built from exactly the sources (and flags) next to it."""
    stamp = os.path.join(HERE, 'csrc', '.build_stamp_instr' if instrumented else '.build_stamp')
    lib_path = INSTR_LIB_PATH if instrumented else LIB_PATH
    if instrumented and os.environ.get('DSB_VARIANT_FLAGS'):
        return os.path.exists(lib_path)
    if not (os.path.exists(lib_path) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False, instrumented: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a and link the shared library. Returns its path.
    instrumented=True builds the ablation / cycle-accounting variant next to the product library."""
    stamp = os.path.join(HERE, 'csrc', '.build_stamp_instr' if instrumented else '.build_stamp')
    lib_path = INSTR_LIB_PATH if instrumented else LIB_PATH
    extra = ['-DDSB_TC_INSTRUMENT=1'] if instrumented else []
    if instrumented and os.environ.get('DSB_VARIANT_FLAGS'):      # tuning builds reuse the second library slot
        extra = os.environ['DSB_VARIANT_FLAGS'].split()
    dig = _digest()
    if not force and os.path.exists(lib_path) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == dig:
                return lib_path
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'csrc', 'build_instr' if instrumented else 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out, file=sys.stderr)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), out))
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Wno-deprecated-gpu-targets', '-o', lib_path] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed: %s\n%s' % (' '.join(cmd), r.stdout))
    with open(stamp, 'w') as f:
        f.write(dig)
    return lib_path


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
    if '--instrumented' in sys.argv:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, instrumented=True))

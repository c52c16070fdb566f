"""Builds tuning variants of the library (different unroll factors of the edge kernels' hot loops) into
profiles/variants/*.so; time one with  DSB_LIB_PATH=profiles/variants/<name>.so DSB_INSTRUMENT=0 python profiles/tc_ablate.py 0"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsbdd_b200 import _build  # noqa: E402

VARIANTS = {'p8': ['-DDSB_P_UNROLL=8'], 'p4': ['-DDSB_P_UNROLL=4'], 'e2': ['-DDSB_E1_UNROLL=2', '-DDSB_E2_UNROLL=2'],
            'p4e2': ['-DDSB_P_UNROLL=4', '-DDSB_E1_UNROLL=2', '-DDSB_E2_UNROLL=2'], 'p1': ['-DDSB_P_UNROLL=1']}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'variants')
os.makedirs(out, exist_ok=True)
nvcc = _build._nvcc()
for name, flags in VARIANTS.items():
    objs = []
    for src in _build.SOURCES:
        obj = os.path.join(out, f'{name}_{src[:-3]}.o')
        objs.append(obj)
        subprocess.check_call([nvcc] + _build.NVCC_FLAGS + flags + ['-c', os.path.join(_build.CSRC, src), '-o', obj])
    lib = os.path.join(out, f'{name}.so')
    subprocess.check_call([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Wno-deprecated-gpu-targets', '-o', lib] + objs)
    for o in objs:
        os.remove(o)
    print(lib)

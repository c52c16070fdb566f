"""Builds tuning variants of the library into profiles/variants/<name>.so: dsb_tc.cu is recompiled with the variant's
-D flags and linked with the product objects of the other translation units (diffsbdd_b200/csrc/build).

    python profiles/build_variants.py name1=-DFOO=1,-DBAR=2 name2=-DFOO=0 ...
    DSB_LIB_PATH=profiles/variants/<name>.so DSB_INSTRUMENT=0 python profiles/tc_ablate.py 0      (or profiles/time_variants.py)"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsbdd_b200 import _build  # noqa: E402

_build.build()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'variants')
os.makedirs(out, exist_ok=True)
nvcc = _build._nvcc()
objdir = os.path.join(_build.CSRC, 'build')
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition('=')
    flags = [f for f in flags.split(',') if f]
    obj = os.path.join(out, f'{name}_dsb_tc.o')
    procs.append((name, obj, subprocess.Popen([nvcc] + _build.NVCC_FLAGS + flags + ['-c', os.path.join(_build.CSRC, 'dsb_tc.cu'), '-o', obj])))
for name, obj, p in procs:
    if p.wait() != 0:
        raise SystemExit(f'variant {name} failed to compile')
    others = [os.path.join(objdir, s.replace('.cu', '.o')) for s in _build.SOURCES if s != 'dsb_tc.cu']
    lib = os.path.join(out, f'{name}.so')
    subprocess.check_call([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Wno-deprecated-gpu-targets', '-o', lib, obj] + others)
    os.remove(obj)
    print(lib)

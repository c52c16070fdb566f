"""Times the edge / node kernel classes of every library in profiles/variants/ (and the product library) on the
BASELINE configs[2] forward: one subprocess per library (a process loads the library once)."""
import glob
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
libs = [('product', '')] + [(os.path.basename(p)[:-3], p) for p in sorted(glob.glob(os.path.join(here, 'variants', '*.so')))]
for name, path in libs:
    env = dict(os.environ, DSB_INSTRUMENT='0')
    if path:
        env['DSB_LIB_PATH'] = path
    r = subprocess.run([sys.executable, os.path.join(here, 'tc_ablate.py'), '0'], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith('flags')]
    print(f'{name:24s} {line[0] if line else r.stderr[-400:]}', flush=True)

"""One golden forward through the native kernels, for compute-sanitizer:
    compute-sanitizer --tool racecheck|synccheck|memcheck python profiles/sanitize_forward.py <golden case> <math mode>"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import load_golden, assert_close  # noqa: E402
from diffsbdd_b200.dynamics import EGNNDynamics  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else 'fullatom_b2_n200_l6'
mode = sys.argv[2] if len(sys.argv) > 2 else '3xfp16'
cfg, sd, inp, want, _ = load_golden(case)
net = EGNNDynamics.from_config(cfg, device='cuda')
net.load_state_dict(sd)
net.eval()
net.math_mode = mode
with torch.no_grad():
    out = net(*[x.cuda() for x in inp])
torch.cuda.synchronize()
err = assert_close(out[0].cpu(), want[0], f'{case} {mode}')
print(f'{case} mode={mode}: max abs err {err:.2e} (E={net.last_num_edges})')

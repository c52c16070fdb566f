"""Bring-up helper for the tensor-core path: runs golden cases with DSB_MATH_MODE from the environment and prints the
error against the reference golden vectors (one process per mode, so a trapped kernel cannot poison the next mode)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from helpers import load_golden  # noqa: E402
from diffsbdd_b200.dynamics import EGNNDynamics  # noqa: E402

mode = os.environ.get('DSB_MATH_MODE', 'auto')
for case in sys.argv[1:] or ['config1_n64_l4', 'fullatom_b2_n200_l6', 'ca_b3_l6']:
    cfg, sd, inp, want, edges = load_golden(case)
    net = EGNNDynamics.from_config(cfg, device='cuda')
    net.load_state_dict(sd)
    net.eval()
    t0 = time.time()
    with torch.no_grad():
        out = net(*[x.cuda() for x in inp])
    torch.cuda.synchronize()
    ea = (out[0].cpu() - want[0]).abs().max().item()
    er = (out[1].cpu() - want[1]).abs().max().item()
    print(f'mode={mode} (bits {net.math_mode}) {case}: max|err| ligand={ea:.3e} pocket={er:.3e}  ({time.time() - t0:.2f}s)', flush=True)

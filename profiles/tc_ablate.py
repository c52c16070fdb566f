"""Ablation timing of the tensor-core edge kernel at BASELINE configs[2] size: which role bounds the tile time?
flags: 1 skip weight copies, 2 skip producer work, 4 skip epilogue work, 8 skip MMAs (results are garbage when set)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsbdd_b200 import _native, synthetic as syn  # noqa: E402
from diffsbdd_b200.config import FULLATOM_COND  # noqa: E402
from diffsbdd_b200.dynamics import EGNNDynamics  # noqa: E402

B = 64
cfg = FULLATOM_COND
net = EGNNDynamics.from_config(cfg, device='cuda')
net.load_state_dict(syn.synthetic_state_dict(cfg, 0))
net.eval()
net.defer_status_check = True
inp = [x.cuda() for x in syn.synthetic_denoiser_inputs(cfg, [25] * B, [175] * B, seed=3)]
lib = _native.load()
lib.dsb_debug_set_tc_flags.argtypes = [C.c_int]
with torch.no_grad():
    net(*inp)
    net.set_profiling(True)
    for flags in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 3, 5, 6, 7, 15]:
        lib.dsb_debug_set_tc_flags(flags)
        net(*inp); net(*inp)
        net.collect_profile(reset=True)
        for _ in range(5):
            net(*inp)
        p = net.collect_profile(reset=True)
        print(f'flags={flags:2d}  edge_gcl {p["edge_gcl"]["ms"] / 30 * 1e3:8.1f} us/launch   edge_coord {p["edge_coord"]["ms"] / 30 * 1e3:8.1f}'
              f'   node_gemm {p["node_gemm"]["ms"] / 120 * 1e3:7.1f} us/launch', flush=True)
    lib.dsb_debug_set_tc_flags(0)
print('edges', int(net._status[1]))

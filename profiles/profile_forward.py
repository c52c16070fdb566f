"""Profiling driver: a few eager denoiser forwards at BASELINE configs[2] size (B=64, N_L=25, N_P=175) for
ncu.  Run under `ncu --profile-from-start off ...`: only the region between cudaProfilerStart/Stop is captured.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python profiles/profile_forward.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsbdd_b200 import synthetic as syn  # noqa: E402
from diffsbdd_b200.config import FULLATOM_COND  # noqa: E402
from diffsbdd_b200.dynamics import EGNNDynamics  # noqa: E402

B = int(os.environ.get('PROFILE_BATCH', '64'))
CALLS = int(os.environ.get('PROFILE_CALLS', '1'))
cfg = FULLATOM_COND
net = EGNNDynamics.from_config(cfg, device='cuda')
net.load_state_dict(syn.synthetic_state_dict(cfg, 0))
net.eval()
inp = [x.cuda() for x in syn.synthetic_denoiser_inputs(cfg, [25] * B, [175] * B, seed=3)]
with torch.no_grad():
    for _ in range(3):
        net(*inp)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(CALLS):
        net(*inp)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('edges', net.last_num_edges, 'launches/forward', net.launches_per_forward)

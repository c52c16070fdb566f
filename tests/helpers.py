"""Shared helpers for the parity tests."""
import glob
import json
import os

import numpy as np
import torch

from diffsbdd_b200.config import DynamicsConfig
from diffsbdd_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
# stated fp32 parity tolerance for one denoiser forward (~20x the reference's own
# fp32-vs-fp64 noise floor of 3-4e-7, far below TF32's ~1e-3)
ATOL, RTOL = 1e-5, 1e-4


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    cfg = DynamicsConfig(**json.loads(str(z['cfg'])))
    sd = syn.synthetic_state_dict(cfg, int(z['weight_seed']))
    chk = syn.state_dict_checksum(sd)
    assert abs(chk - float(z['weight_checksum'])) <= 1e-9 * max(1.0, abs(chk)), 'weight recipe drifted'
    inp = (torch.from_numpy(z['xh_atoms']), torch.from_numpy(z['xh_residues']), torch.from_numpy(z['t']),
           torch.from_numpy(z['mask_atoms']), torch.from_numpy(z['mask_residues']))
    out = (torch.from_numpy(z['out_atoms']), torch.from_numpy(z['out_residues']))
    edges = torch.from_numpy(z['edges'].astype(np.int64))
    return cfg, sd, inp, out, edges


def assert_close(got, want, what, atol=ATOL, rtol=RTOL):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    worst = float((err - tol).max()) if err.numel() else -1.0
    assert worst <= 0, f'{what}: max abs err {float(err.max()):.3e} exceeds atol={atol} rtol={rtol}'
    return float(err.max()) if err.numel() else 0.0

"""Generates tests/golden/ddpm/*.npz: outputs of the UNMODIFIED reference ``ConditionalDDPM`` samplers
(conditional_model.py:479, :558, :364) driven by a CPU denoiser stand-in (the oracle restatement of
EGNNDynamics.forward, bit-identical to the reference module on CPU — tests/test_oracle_golden.py), with
fixed torch seeds.  They pin the DDPM wrapper of this repo (schedule, mu/sigma update, COM handling,
RePaint loop) independently of the CUDA kernels.  Build-container only."""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from diffsbdd_b200 import synthetic as syn  # noqa: E402
from oracle import ref_shim  # noqa: E402
from ddpm_cases import DDPM_CFG, HIST, OracleDynamics, make_pocket, make_ligand, SAMPLER_CASES  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ddpm')


def main():
    ref = ref_shim.load_reference()
    sd = syn.synthetic_state_dict(DDPM_CFG, 5)
    for name, spec in SAMPLER_CASES.items():
        dyn = OracleDynamics(DDPM_CFG, sd)
        ddpm = ref.ConditionalDDPM(dynamics=dyn, atom_nf=DDPM_CFG.atom_nf, residue_nf=DDPM_CFG.residue_nf,
                                   n_dims=3, timesteps=spec['T'], noise_schedule='polynomial_2',
                                   noise_precision=5e-4, loss_type='l2', norm_values=(1, 4), size_histogram=HIST)
        ddpm.eval()
        pocket = make_pocket()
        torch.manual_seed(spec['seed'])
        if spec['kind'] == 'sample':
            out = ddpm.sample_given_pocket(pocket, torch.tensor(spec['n_lig']), return_frames=spec['frames'],
                                           timesteps=spec['timesteps'])
        elif spec['kind'] == 'inpaint':
            lig, fixed = make_ligand(spec['n_lig'], spec['n_fixed'])
            out = ddpm.inpaint(lig, pocket, fixed, resamplings=spec['resamplings'], timesteps=spec['timesteps'],
                               center=spec['center'])
        else:
            lig, _ = make_ligand(spec['n_lig'], 0)
            out = ddpm.diversify(lig, pocket, noising_steps=spec['noising_steps'])
        np.savez_compressed(os.path.join(OUT, name + '.npz'), xh_lig=out[0].numpy(), xh_pocket=out[1].numpy(),
                            lig_mask=out[2].numpy(), pocket_mask=out[3].numpy(),
                            gamma=ddpm.gamma.gamma.detach().numpy())
        print(name, tuple(out[0].shape), float(out[0].abs().max()))


if __name__ == '__main__':
    main()

"""CPU: this repo's DDPM wrapper (diffsbdd_b200.conditional_model / en_diffusion, eager engine) must
reproduce what the UNMODIFIED reference samplers produced (tests/golden/ddpm/*.npz,
tests/golden/make_golden_ddpm.py) given the same denoiser stand-in and the same torch seed."""
import os

import numpy as np
import pytest
import torch

from ddpm_cases import (DDPM_CFG, HIST, OracleDynamics, make_pocket, make_ligand, SAMPLER_CASES, JOINT_CFG,
                        JOINT_CASES, make_pocket_fixed)
from diffsbdd_b200 import synthetic as syn
from diffsbdd_b200.conditional_model import ConditionalDDPM
from diffsbdd_b200.en_diffusion import EnVariationalDiffusion, PredefinedNoiseSchedule, scatter_mean

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ddpm')


def build(spec):
    sd = syn.synthetic_state_dict(DDPM_CFG, 5)
    ddpm = ConditionalDDPM(dynamics=OracleDynamics(DDPM_CFG, sd), atom_nf=DDPM_CFG.atom_nf,
                           residue_nf=DDPM_CFG.residue_nf, n_dims=3, timesteps=spec['T'],
                           noise_schedule='polynomial_2', noise_precision=5e-4, loss_type='l2',
                           norm_values=(1, 4), size_histogram=HIST)
    ddpm.eval()
    return ddpm


@pytest.mark.parametrize('name', sorted(SAMPLER_CASES))
def test_sampler_matches_reference_golden(name):
    spec = SAMPLER_CASES[name]
    z = np.load(os.path.join(GOLD, name + '.npz'))
    ddpm = build(spec)
    assert np.array_equal(ddpm.gamma.gamma.numpy(), z['gamma']), 'noise schedule table differs'
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
    for got, key in zip(out, ('xh_lig', 'xh_pocket', 'lig_mask', 'pocket_mask')):
        want = torch.from_numpy(z[key])
        assert got.shape == want.shape
        if got.dtype.is_floating_point:
            assert torch.allclose(got, want, atol=1e-5, rtol=1e-5), (key, float((got - want).abs().max()))
        else:
            assert torch.equal(got, want)


def _compare(out, z):
    for got, key in zip(out, ('xh_lig', 'xh_pocket', 'lig_mask', 'pocket_mask')):
        want = torch.from_numpy(z[key])
        assert got.shape == want.shape
        if got.dtype.is_floating_point:
            assert torch.allclose(got, want, atol=1e-5, rtol=1e-5), (key, float((got - want).abs().max()))
        else:
            assert torch.equal(got, want)


@pytest.mark.parametrize('name', sorted(JOINT_CASES))
def test_joint_model_matches_reference_golden(name):
    """EnVariationalDiffusion.sample / .inpaint (en_diffusion.py:839, :677) incl. RePaint jumps and frames."""
    spec = JOINT_CASES[name]
    z = np.load(os.path.join(GOLD, name + '.npz'))
    sd = syn.synthetic_state_dict(JOINT_CFG, 6)
    ddpm = EnVariationalDiffusion(dynamics=OracleDynamics(JOINT_CFG, sd), atom_nf=JOINT_CFG.atom_nf,
                                  residue_nf=JOINT_CFG.residue_nf, n_dims=3, timesteps=spec['T'],
                                  noise_schedule='polynomial_2', noise_precision=5e-4, loss_type='l2',
                                  norm_values=(1, 4), size_histogram=HIST)
    ddpm.eval()
    assert np.array_equal(ddpm.gamma.gamma.numpy(), z['gamma'])
    pocket = make_pocket()
    torch.manual_seed(spec['seed'])
    if spec['kind'] == 'sample':
        out = ddpm.sample(len(spec['n_lig']), torch.tensor(spec['n_lig']), pocket['size'],
                          return_frames=spec['frames'], timesteps=spec['timesteps'])
    else:
        lig, fixed = make_ligand(spec['n_lig'], spec['n_fixed'])
        out = ddpm.inpaint(lig, pocket, fixed, make_pocket_fixed(spec, pocket), resamplings=spec['resamplings'],
                           jump_length=spec['jump_length'], return_frames=spec['frames'],
                           timesteps=spec['timesteps'])
    _compare(out, z)


def test_repaint_schedule_properties():
    """en_diffusion.py:653-674: total denoising steps = timesteps + jump_length * (#jumps)."""
    f = EnVariationalDiffusion.get_repaint_schedule
    assert f(1, 1, 5) == [5]
    assert f(2, 1, 3) == [1, 2, 2, 1] or sum(f(2, 1, 3)) >= 3
    for T in (4, 10, 50):
        for r in (1, 3):
            for j in (1, 2, 5):
                sched = f(r, j, T)
                assert sum(sched) - j * (len(sched) - 1) == T


def test_cosine_schedule_is_monotone():
    g = PredefinedNoiseSchedule('cosine', timesteps=100, precision=1e-4).gamma
    assert torch.all(g[1:] >= g[:-1])


def test_sample_keeps_ligand_com_free():
    spec = SAMPLER_CASES['sample_T6']
    ddpm = build(spec)
    torch.manual_seed(0)
    xh_lig, xh_pocket, lig_mask, _ = ddpm.sample_given_pocket(make_pocket(), torch.tensor(spec['n_lig']))
    com = scatter_mean(xh_lig[:, :3], lig_mask)
    assert com.abs().max() < 5e-2
    assert torch.all(xh_lig[:, 3:].sum(1) == 1)     # one-hot atom types (conditional_model.py:132)

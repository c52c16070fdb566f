"""Shared definitions of the DDPM-wrapper parity cases (used by the golden generator and the tests)."""
import torch

from diffsbdd_b200.config import DynamicsConfig
from diffsbdd_b200 import synthetic as syn
from oracle.cpu_denoiser import OracleDynamics  # noqa: F401

# small conditional denoiser (kernel-supported dims: H=64) so CPU loops stay fast
DDPM_CFG = DynamicsConfig(joint_nf=16, hidden_nf=64, n_layers=2)
HIST = [[0.0, 1.0, 2.0], [1.0, 3.0, 1.0], [2.0, 1.0, 0.5]]
N_POCKET = [22, 17]

SAMPLER_CASES = {
    'sample_T6': dict(kind='sample', T=6, timesteps=None, frames=1, n_lig=[7, 5], seed=101),
    'sample_T12_frames3_sub6': dict(kind='sample', T=12, timesteps=6, frames=3, n_lig=[4, 9], seed=102),
    'inpaint_T4_r2_ligand': dict(kind='inpaint', T=8, timesteps=4, resamplings=2, n_lig=[8, 6], n_fixed=3,
                                 center='ligand', seed=103),
    'inpaint_T3_r1_pocket': dict(kind='inpaint', T=3, timesteps=None, resamplings=1, n_lig=[5, 7], n_fixed=2,
                                 center='pocket', seed=104),
    'diversify_3of10': dict(kind='diversify', T=10, noising_steps=3, n_lig=[6, 6], seed=105),
}


# joint model (update_pocket_coords=True): EnVariationalDiffusion.sample / .inpaint (en_diffusion.py:839, :677)
JOINT_CFG = DynamicsConfig(joint_nf=16, hidden_nf=64, n_layers=2, update_pocket_coords=True)
JOINT_CASES = {
    'joint_sample_T5': dict(kind='sample', T=5, timesteps=None, frames=1, n_lig=[6, 4], seed=201),
    'joint_inpaint_T6_r2_j2': dict(kind='inpaint', T=6, timesteps=None, resamplings=2, jump_length=2, frames=1,
                                   n_lig=[7, 5], n_fixed=2, pocket_fixed=True, seed=202),
    'joint_inpaint_T8_sub4_frames2': dict(kind='inpaint', T=8, timesteps=4, resamplings=1, jump_length=1, frames=2,
                                          n_lig=[5, 6], n_fixed=0, pocket_fixed=True, seed=203),
    'joint_inpaint_T4_r3_partial_pocket': dict(kind='inpaint', T=4, timesteps=None, resamplings=3, jump_length=1,
                                               frames=1, n_lig=[6, 6], n_fixed=3, pocket_fixed=False, seed=204),
}


def make_pocket_fixed(spec, pocket):
    """0/1 per pocket node: all fixed, or every third node free (exercises the pocket blend of en_diffusion.py:775)."""
    f = torch.ones(len(pocket['mask']))
    if not spec['pocket_fixed']:
        f[::3] = 0
    return f


def make_pocket(device='cpu'):
    p = syn.synthetic_pocket(DDPM_CFG, N_POCKET, seed=31, spread=3.0)
    return {k: v.to(device) for k, v in p.items()}


def make_ligand(n_lig, n_fixed, device='cpu'):
    """Reference ``ligand`` dict + 0/1 ``lig_fixed`` (inpaint.py:117-141: first n_fixed atoms of every sample)."""
    g = torch.Generator().manual_seed(77)
    n = sum(n_lig)
    mask = torch.repeat_interleave(torch.arange(len(n_lig)), torch.tensor(n_lig))
    x = torch.randn((n, 3), generator=g) * 1.5
    types = torch.randint(0, DDPM_CFG.atom_nf, (n,), generator=g)
    fixed = torch.zeros(n)
    start = 0
    for k in n_lig:
        fixed[start:start + n_fixed] = 1
        start += k
    lig = {'x': x.to(device), 'one_hot': torch.nn.functional.one_hot(types, DDPM_CFG.atom_nf).float().to(device),
           'size': torch.tensor(n_lig, device=device), 'mask': mask.to(device)}
    return lig, fixed.to(device)

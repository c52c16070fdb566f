"""GPU: DDPM samplers around the native denoiser — the fused update kernel (dsb_ddpm_ligand_update), the
eager (reference-order) loop against the CPU oracle-driven wrapper with injected noise, and the CUDA-graphed
loop against the eager loop."""
import ctypes as C

import pytest
import torch

from ddpm_cases import DDPM_CFG, HIST, make_pocket, make_ligand
from oracle.cpu_denoiser import OracleDynamics
from diffsbdd_b200 import _native, synthetic as syn
from diffsbdd_b200.conditional_model import ConditionalDDPM
from diffsbdd_b200.dynamics import EGNNDynamics
from diffsbdd_b200.en_diffusion import scatter_mean

pytestmark = pytest.mark.gpu


def build(T, device, native=True, engine='auto'):
    sd = syn.synthetic_state_dict(DDPM_CFG, 5)
    if native:
        dyn = EGNNDynamics.from_config(DDPM_CFG, device=device)
        dyn.load_state_dict(sd)
    else:
        dyn = OracleDynamics(DDPM_CFG, sd)
    ddpm = ConditionalDDPM(dynamics=dyn, atom_nf=DDPM_CFG.atom_nf, residue_nf=DDPM_CFG.residue_nf, n_dims=3,
                           timesteps=T, noise_schedule='polynomial_2', noise_precision=5e-4, loss_type='l2',
                           norm_values=(1, 4), size_histogram=HIST)
    ddpm.loop_engine = engine
    return ddpm.to(device).eval()


def test_fused_ddpm_update_kernel_matches_torch():
    g = torch.Generator().manual_seed(0)
    n_lig, n_poc = [5, 1, 9], [11, 7, 3]
    A, R = 10, 10
    lm = torch.repeat_interleave(torch.arange(3), torch.tensor(n_lig)).cuda()
    pm = torch.repeat_interleave(torch.arange(3), torch.tensor(n_poc)).cuda()
    z = torch.randn((15, 3 + A), generator=g).cuda()
    eps = torch.randn((15, 3 + A), generator=g).cuda()
    noise = torch.randn((15, 3 + A), generator=g).cuda()
    pocket = torch.randn((21, 3 + R), generator=g).cuda()
    coef = (torch.rand((3, 3), generator=g) + 0.5).cuda()
    mu = z / coef[lm, 0:1] - coef[lm, 1:2] * eps
    want = mu + coef[lm, 2:3] * noise
    com = scatter_mean(want[:, :3], lm)
    want[:, :3] -= com[lm]
    want_p = pocket.clone()
    want_p[:, :3] -= com[pm]
    z_out, p_out = torch.empty_like(z), torch.empty_like(pocket)
    lib = _native.load()
    _native.check(lib.dsb_ddpm_ligand_update(z.data_ptr(), eps.data_ptr(), noise.data_ptr(), coef.data_ptr(),
                                             lm.data_ptr(), pm.data_ptr(), pocket.data_ptr(), 15, 21, 3, A, R,
                                             z_out.data_ptr(), p_out.data_ptr(),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.allclose(z_out, want, atol=2e-6, rtol=1e-6)
    assert torch.allclose(p_out, want_p, atol=2e-6, rtol=1e-6)
    # in place
    z2, p2 = z.clone(), pocket.clone()
    _native.check(lib.dsb_ddpm_ligand_update(z2.data_ptr(), eps.data_ptr(), noise.data_ptr(), coef.data_ptr(),
                                             lm.data_ptr(), pm.data_ptr(), p2.data_ptr(), 15, 21, 3, A, R,
                                             z2.data_ptr(), p2.data_ptr(),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(z2, z_out) and torch.equal(p2, p_out)


class _NoiseTape:
    """Replaces ``sample_gaussian`` so a CPU run and a GPU run consume the same noise."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)

    def __call__(self, size, device):
        return torch.randn(size, generator=self.g).to(device)


def test_eager_loop_matches_cpu_wrapper_with_injected_noise():
    T = 6
    n_lig = torch.tensor([7, 5])
    cpu = build(T, 'cpu', native=False)
    cpu.sample_gaussian = _NoiseTape(9)
    want = cpu.sample_given_pocket(make_pocket(), n_lig)
    gpu = build(T, 'cuda', native=True, engine='eager')
    gpu.sample_gaussian = _NoiseTape(9)
    got = gpu.sample_given_pocket(make_pocket('cuda'), n_lig.cuda())
    assert torch.equal(got[2].cpu(), want[2])
    scale = float(want[0][:, :3].abs().max())
    assert torch.allclose(got[0][:, :3].cpu(), want[0][:, :3], atol=1e-4 * scale)
    assert torch.equal(got[0][:, 3:].cpu(), want[0][:, 3:])          # argmax'd one-hot types
    assert torch.allclose(got[1].cpu(), want[1], atol=1e-4 * scale)


def test_inpaint_eager_matches_cpu_wrapper_with_injected_noise():
    T = 4
    cpu = build(T, 'cpu', native=False)
    cpu.sample_gaussian = _NoiseTape(4)
    lig, fixed = make_ligand([8, 6], 3)
    want = cpu.inpaint(lig, make_pocket(), fixed, resamplings=2)
    gpu = build(T, 'cuda', native=True, engine='eager')
    gpu.sample_gaussian = _NoiseTape(4)
    lig_g, fixed_g = make_ligand([8, 6], 3, device='cuda')
    got = gpu.inpaint(lig_g, make_pocket('cuda'), fixed_g, resamplings=2)
    scale = float(want[0][:, :3].abs().max())
    assert torch.allclose(got[0][:, :3].cpu(), want[0][:, :3], atol=1e-4 * scale)
    assert torch.equal(got[0][:, 3:].cpu(), want[0][:, 3:])


def test_graph_loop_matches_eager_loop_same_seed():
    """Both engines draw one randn((N_L, 3+A)) per reverse step from the CUDA generator; with the same seed the
    trajectories must agree (the fused kernel and the torch ops differ only in rounding)."""
    T = 10
    n_lig = torch.tensor([7, 5]).cuda()
    eager = build(T, 'cuda', engine='eager')
    graph = build(T, 'cuda', engine='graph')
    torch.manual_seed(77)
    a = eager.sample_given_pocket(make_pocket('cuda'), n_lig)
    torch.manual_seed(77)
    b = graph.sample_given_pocket(make_pocket('cuda'), n_lig)
    assert graph._graph_cache, 'graph engine did not capture'
    scale = float(a[0][:, :3].abs().max())
    same_rng = torch.allclose(a[0][:, :3], b[0][:, :3], atol=1e-3 * scale)
    if not same_rng:
        pytest.xfail('in-graph normal_() consumes the Philox stream differently from eager randn on this torch build; '
                     'the engines are then compared statistically in test_graph_loop_statistics')
    assert torch.equal(a[0][:, 3:], b[0][:, 3:])
    # second call with fresh mask tensors of the same layout must reuse the captured graph
    g0 = next(iter(graph._graph_cache.values()))['graph']
    graph.sample_given_pocket(make_pocket('cuda'), n_lig.clone())
    assert next(iter(graph._graph_cache.values()))['graph'] is g0


def test_graph_loop_invariants_and_frames():
    T = 12
    ddpm = build(T, 'cuda', engine='graph')
    n_lig = torch.tensor([4, 9]).cuda()
    torch.manual_seed(3)
    xh_lig, xh_pocket, lig_mask, pocket_mask = ddpm.sample_given_pocket(make_pocket('cuda'), n_lig, return_frames=3, timesteps=6)
    assert xh_lig.shape == (3, 13, 13) and xh_pocket.shape[0] == 3
    assert torch.isfinite(xh_lig).all() and torch.isfinite(xh_pocket).all()
    final = xh_lig[0]
    assert torch.all(final[:, 3:].sum(1) == 1)
    com = scatter_mean(final[:, :3], lig_mask)
    assert com.abs().max() < 5e-2
    # pocket moved rigidly: pairwise distances of the pocket are preserved
    p0 = make_pocket('cuda')
    d_before = torch.cdist(p0['x'][:22], p0['x'][:22])
    d_after = torch.cdist(xh_pocket[0][:22, :3], xh_pocket[0][:22, :3])
    assert torch.allclose(d_before, d_after, atol=1e-3)


def test_graph_loop_statistics():
    """Distribution-level agreement of the two engines (independent seeds): per-atom coordinate spread after a short
    loop from the same prior must agree within sampling error."""
    T = 8
    n = torch.full((2,), 40).cuda()
    spreads = {}
    for engine, seed in (('eager', 1), ('graph', 2)):
        ddpm = build(T, 'cuda', engine=engine)
        torch.manual_seed(seed)
        xs = []
        for _ in range(6):
            out = ddpm.sample_given_pocket(make_pocket('cuda'), n)
            xs.append(out[0][:, :3])
        spreads[engine] = float(torch.cat(xs).std())
    assert abs(spreads['eager'] - spreads['graph']) < 0.25 * spreads['eager']


def _build_joint(T, device, native):
    from ddpm_cases import JOINT_CFG
    from diffsbdd_b200.en_diffusion import EnVariationalDiffusion
    sd = syn.synthetic_state_dict(JOINT_CFG, 6)
    if native:
        dyn = EGNNDynamics.from_config(JOINT_CFG, device=device)
        dyn.load_state_dict(sd)
    else:
        dyn = OracleDynamics(JOINT_CFG, sd)
    ddpm = EnVariationalDiffusion(dynamics=dyn, atom_nf=JOINT_CFG.atom_nf, residue_nf=JOINT_CFG.residue_nf,
                                  n_dims=3, timesteps=T, noise_schedule='polynomial_2', noise_precision=5e-4,
                                  loss_type='l2', norm_values=(1, 4), size_histogram=HIST)
    return ddpm.to(device).eval()


class _Recorder(torch.nn.Module):
    """Wraps a denoiser and keeps every (inputs, outputs) pair it was called with."""

    def __init__(self, inner):
        super().__init__()
        self.inner, self.calls = inner, []
        for k in ('update_pocket_coords', 'atom_nf', 'residue_nf', 'n_dims'):
            if hasattr(inner, k):
                setattr(self, k, getattr(inner, k))

    def forward(self, *args):
        out = self.inner(*args)
        self.calls.append(([a.clone() for a in args], [o.clone() for o in out]))
        return out


def test_joint_repaint_inpaint_denoiser_calls_match_oracle():
    """EnVariationalDiffusion.inpaint (en_diffusion.py:677-837, RePaint jumps, pocket partially free) around the joint
    denoiser (update_pocket_coords=True).  The joint trajectory with random weights is chaotic, so instead of the end
    point every denoiser call of the CPU run (inputs as the sampler really produces them: noised, COM-shifted,
    t per graph) is replayed through the native kernels and compared call by call; the GPU wrapper itself must run
    end to end and keep the sampler's invariants."""
    from ddpm_cases import JOINT_CASES, make_pocket_fixed
    spec = JOINT_CASES['joint_inpaint_T6_r2_j2']
    cpu = _build_joint(spec['T'], 'cpu', native=False)
    cpu.dynamics = _Recorder(cpu.dynamics)
    cpu.sample_gaussian = _NoiseTape(12)
    lig, fixed = make_ligand(spec['n_lig'], spec['n_fixed'])
    pocket = make_pocket()
    pfix = make_pocket_fixed(dict(pocket_fixed=False), pocket)
    cpu.inpaint(lig, pocket, fixed, pfix, resamplings=2, jump_length=2)
    calls = cpu.dynamics.calls
    assert len(calls) >= spec['T']
    gpu = _build_joint(spec['T'], 'cuda', native=True)
    for args, want in calls:
        got = gpu.dynamics(*[a.cuda() for a in args])
        for g, w in zip(got, want):
            scale = max(1.0, float(w.abs().max()))
            assert torch.allclose(g.cpu(), w, atol=1e-5 * scale, rtol=1e-4), float((g.cpu() - w).abs().max())
    gpu.sample_gaussian = _NoiseTape(12)
    lig_g, fixed_g = make_ligand(spec['n_lig'], spec['n_fixed'], device='cuda')
    out = gpu.inpaint(lig_g, make_pocket('cuda'), fixed_g, pfix.cuda(), resamplings=2, jump_length=2)
    assert all(torch.isfinite(o).all() for o in out[:2])
    assert torch.all(out[0][:, 3:].sum(1) == 1) and torch.all(out[1][:, 3:].sum(1) == 1)
    com = scatter_mean(torch.cat((out[0][:, :3], out[1][:, :3])), torch.cat((out[2], out[3])))
    assert com.abs().max() < 5e-2 * max(1.0, float(out[1][:, :3].abs().max()))

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
    g0 = next(iter(graph._graph_cache.values()))['graphs']['reverse']
    graph.sample_given_pocket(make_pocket('cuda'), n_lig.clone())
    assert next(iter(graph._graph_cache.values()))['graphs']['reverse'] is g0


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


def test_fused_inpaint_kernel_matches_torch_ops():
    """dsb_ddpm_inpaint_update against the torch ops of the eager RePaint iteration (conditional_model.py:636-666),
    with and without the re-noising step, ragged graphs incl. a graph without fixed atoms."""
    g = torch.Generator().manual_seed(1)
    n_lig, n_poc = [6, 1, 9, 4], [11, 7, 3, 8]
    A, R, B = 10, 10, 4
    lm = torch.repeat_interleave(torch.arange(B), torch.tensor(n_lig)).cuda()
    pm = torch.repeat_interleave(torch.arange(B), torch.tensor(n_poc)).cuda()
    NL, NP = sum(n_lig), sum(n_poc)
    z_unknown = torch.randn((NL, 3 + A), generator=g).cuda()
    pocket = torch.randn((NP, 3 + R), generator=g).cuda()
    known = torch.randn((NL, 3 + A), generator=g).cuda()
    com0 = torch.randn((B, 3), generator=g).cuda()
    fixed = (torch.rand(NL, generator=g) < 0.4).float().cuda()
    fixed[lm == 3] = 0                       # a graph with nothing fixed
    fixed[0] = 1
    n1 = torch.randn((NL, 3 + A), generator=g).cuda()
    n2 = torch.randn((NL, 3 + A), generator=g).cuda()
    coef = (torch.rand((B, 4), generator=g) * 0.8 + 0.1).cuda()
    lib = _native.load()
    for renoise in (False, True):
        # torch ops in eager order
        com_pocket = scatter_mean(pocket[:, :3], pm)
        xk = known.clone()
        xk[:, :3] = known[:, :3] + (com_pocket - com0)[lm]
        zk = coef[lm, 0:1] * xk + coef[lm, 1:2] * n1
        pk = pocket.clone()
        mean = scatter_mean(zk[:, :3], lm)
        zk[:, :3] = zk[:, :3] - mean[lm]
        pk[:, :3] = pk[:, :3] - mean[pm]
        rows = fixed.bool()
        cn = scatter_mean(zk[rows][:, :3], lm[rows], dim_size=B)
        cd = scatter_mean(z_unknown[rows][:, :3], lm[rows], dim_size=B)
        dx = cd - cn
        zk[:, :3] = zk[:, :3] + dx[lm]
        pk[:, :3] = pk[:, :3] + dx[pm]
        want = zk * fixed[:, None] + z_unknown * (1 - fixed[:, None])
        if renoise:
            want = coef[lm, 2:3] * want + coef[lm, 3:4] * n2
            m2 = scatter_mean(want[:, :3], lm)
            want[:, :3] = want[:, :3] - m2[lm]
            pk[:, :3] = pk[:, :3] - m2[pm]
        z, p = z_unknown.clone(), pocket.clone()
        _native.check(lib.dsb_ddpm_inpaint_update(
            z.data_ptr(), p.data_ptr(), known.data_ptr(), com0.data_ptr(), fixed.data_ptr(), n1.data_ptr(),
            n2.data_ptr() if renoise else None, coef.data_ptr(), lm.data_ptr(), pm.data_ptr(), NL, NP, B, A, R,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert torch.allclose(z, want, atol=3e-6, rtol=1e-5), float((z - want).abs().max())
        assert torch.allclose(p, pk, atol=3e-6, rtol=1e-5), float((p - pk).abs().max())


@pytest.mark.parametrize('resamplings,frames', [(1, 1), (3, 2)])
def test_graph_inpaint_matches_eager_inpaint_same_seed(resamplings, frames):
    """Default engine of ``inpaint`` on CUDA (captured denoiser + fused reverse update + fused RePaint iteration) against
    the reference-order eager loop: same seed -> same randn stream (three draws per inner iteration, two on the last
    resampling), so the trajectories agree to rounding."""
    T = 6
    eager = build(T, 'cuda', engine='eager')
    graph = build(T, 'cuda', engine='auto')
    outs = []
    for ddpm in (eager, graph):
        lig, fixed = make_ligand([8, 6], 3, device='cuda')
        torch.manual_seed(5)
        outs.append(ddpm.inpaint(lig, make_pocket('cuda'), fixed, resamplings=resamplings, return_frames=frames,
                                 center='ligand'))
    a, b = outs
    st = next(iter(graph._graph_cache.values()))
    assert 'inpaint_last' in st['graphs'] and (resamplings == 1 or 'inpaint_renoise' in st['graphs'])
    assert a[0].shape == b[0].shape
    scale = float(a[0][..., :3].abs().max())
    assert torch.allclose(a[0][..., :3], b[0][..., :3], atol=1e-3 * scale), float((a[0] - b[0]).abs().max())
    assert torch.allclose(a[1][..., :3], b[1][..., :3], atol=1e-3 * scale)
    fa, fb = (a[0], b[0]) if frames == 1 else (a[0][0], b[0][0])
    assert torch.equal(fa[:, 3:], fb[:, 3:])                      # argmax'd atom types of the final frame
    # second call, fresh tensors, same layout: the captured graphs are reused
    g0 = st['graphs']['inpaint_last']
    lig, fixed = make_ligand([8, 6], 3, device='cuda')
    graph.inpaint(lig, make_pocket('cuda'), fixed, resamplings=resamplings, return_frames=frames)
    assert next(iter(graph._graph_cache.values()))['graphs']['inpaint_last'] is g0


def test_graph_diversify_matches_eager_diversify_same_seed():
    T = 10
    outs = []
    for engine in ('eager', 'auto'):
        ddpm = build(T, 'cuda', engine=engine)
        lig, _ = make_ligand([6, 6], 0, device='cuda')
        torch.manual_seed(8)
        outs.append(ddpm.diversify(lig, make_pocket('cuda'), noising_steps=4))
    a, b = outs
    scale = float(a[0][:, :3].abs().max())
    assert torch.allclose(a[0][:, :3], b[0][:, :3], atol=1e-3 * scale), float((a[0] - b[0]).abs().max())
    assert torch.equal(a[0][:, 3:], b[0][:, 3:])


def test_captured_graph_is_dropped_when_weights_or_math_mode_change():
    """a cached graph bakes in the packed-weight blob and the kernel selection.  After load_state_dict (or a
    math_mode change) the same-shape call must re-capture, not replay freed weights."""
    T = 5
    n_lig = torch.tensor([7, 5]).cuda()
    graph = build(T, 'cuda', engine='graph')
    torch.manual_seed(21)
    first = graph.sample_given_pocket(make_pocket('cuda'), n_lig)
    g_old = next(iter(graph._graph_cache.values()))['graphs']['reverse']
    sd2 = syn.synthetic_state_dict(DDPM_CFG, 99)
    graph.dynamics.load_state_dict(sd2)
    torch.manual_seed(21)
    second = graph.sample_given_pocket(make_pocket('cuda'), n_lig)
    assert next(iter(graph._graph_cache.values()))['graphs']['reverse'] is not g_old
    eager = build(T, 'cuda', engine='eager')
    eager.dynamics.load_state_dict(sd2)
    torch.manual_seed(21)
    want = eager.sample_given_pocket(make_pocket('cuda'), n_lig)
    scale = float(want[0][:, :3].abs().max())
    assert torch.allclose(second[0][:, :3], want[0][:, :3], atol=1e-3 * scale)
    assert not torch.allclose(second[0][:, :3], first[0][:, :3], atol=1e-3 * scale)
    # math-mode change on a hidden_nf=256 model: replay must not keep the old kernel selection
    from diffsbdd_b200.config import CONFIG1
    cfg = CONFIG1.with_(n_layers=2)
    dyn = EGNNDynamics.from_config(cfg, device='cuda')
    dyn.load_state_dict(syn.synthetic_state_dict(cfg, 3))
    ddpm = ConditionalDDPM(dynamics=dyn, atom_nf=cfg.atom_nf, residue_nf=cfg.residue_nf, n_dims=3, timesteps=T,
                           noise_schedule='polynomial_2', noise_precision=5e-4, loss_type='l2', norm_values=(1, 4),
                           size_histogram=HIST).cuda().eval()
    ddpm.sample_given_pocket(make_pocket('cuda'), n_lig)
    g1 = next(iter(ddpm._graph_cache.values()))['graphs']['reverse']
    dyn.math_mode = 'fp32'
    ddpm.sample_given_pocket(make_pocket('cuda'), n_lig)
    assert next(iter(ddpm._graph_cache.values()))['graphs']['reverse'] is not g1


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


def _joint_ref_noise(nx, lm, pm):
    """COM-free position noise as sample_center_gravity_zero_gaussian_batch builds it (en_diffusion.py:940-944)."""
    cm = torch.cat((lm, pm))
    return nx - scatter_mean(nx, cm)[cm]


def test_fused_joint_kernels_match_torch_ops():
    """dsb_ddpm_joint_update / dsb_ddpm_joint_inpaint_update against the torch ops of the eager joint sampler
    (en_diffusion.py:503-557, :741-807), ragged graphs, partially fixed pocket, with and without the jump back."""
    g = torch.Generator().manual_seed(3)
    n_lig, n_poc = [5, 1, 8], [9, 6, 4]
    A, R, B = 10, 10, 3
    lm = torch.repeat_interleave(torch.arange(B), torch.tensor(n_lig)).cuda()
    pm = torch.repeat_interleave(torch.arange(B), torch.tensor(n_poc)).cuda()
    cm = torch.cat((lm, pm))
    NL, NP = sum(n_lig), sum(n_poc)
    rnd = lambda *shape: torch.randn(shape, generator=g).cuda()
    zl, zp, el, ep = rnd(NL, 3 + A), rnd(NP, 3 + R), rnd(NL, 3 + A), rnd(NP, 3 + R)
    nx, nhl, nhp = rnd(NL + NP, 3), rnd(NL, A), rnd(NP, R)
    coef3 = (torch.rand((B, 3), generator=g) + 0.5).cuda()
    lib = _native.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: t.data_ptr()

    ex = _joint_ref_noise(nx, lm, pm)
    eps_l, eps_p = torch.cat((ex[:NL], nhl), 1), torch.cat((ex[NL:], nhp), 1)
    wl = zl / coef3[lm, 0:1] - coef3[lm, 1:2] * el + coef3[lm, 2:3] * eps_l
    wp = zp / coef3[pm, 0:1] - coef3[pm, 1:2] * ep + coef3[pm, 2:3] * eps_p
    mean = scatter_mean(torch.cat((wl[:, :3], wp[:, :3])), cm)
    wl[:, :3] -= mean[lm]; wp[:, :3] -= mean[pm]
    a, b = zl.clone(), zp.clone()
    _native.check(lib.dsb_ddpm_joint_update(P(a), P(b), P(el), P(ep), P(nx), P(nhl), P(nhp), P(coef3), P(lm), P(pm), NL, NP, B, A, R, stream))
    torch.cuda.synchronize()
    assert torch.allclose(a, wl, atol=3e-6, rtol=1e-5) and torch.allclose(b, wp, atol=3e-6, rtol=1e-5)

    x0l, x0p = rnd(NL, 3 + A), rnd(NP, 3 + R)
    fl = (torch.rand(NL, generator=g) < 0.4).float().cuda()
    fp = (torch.rand(NP, generator=g) < 0.7).float().cuda()
    fl[0] = 1
    coef4 = (torch.rand((B, 4), generator=g) * 0.8 + 0.1).cuda()
    n3 = (rnd(NL + NP, 3), rnd(NL, A), rnd(NP, R))
    for jump in (False, True):
        zkl = coef4[lm, 0:1] * x0l + coef4[lm, 1:2] * eps_l
        zkp = coef4[pm, 0:1] * x0p + coef4[pm, 1:2] * eps_p
        sel_l, sel_p = fl.bool(), fp.bool()
        idx = torch.cat((lm[sel_l], pm[sel_p]))
        com_u = scatter_mean(torch.cat((zl[sel_l][:, :3], zp[sel_p][:, :3])), idx, dim_size=B)
        com_k = scatter_mean(torch.cat((zkl[sel_l][:, :3], zkp[sel_p][:, :3])), idx, dim_size=B)
        shift = com_u - com_k
        zkl[:, :3] += shift[lm]; zkp[:, :3] += shift[pm]
        wl = zkl * fl[:, None] + zl * (1 - fl[:, None])
        wp = zkp * fp[:, None] + zp * (1 - fp[:, None])
        if jump:
            e3 = _joint_ref_noise(n3[0], lm, pm)
            wl = coef4[lm, 2:3] * wl + coef4[lm, 3:4] * torch.cat((e3[:NL], n3[1]), 1)
            wp = coef4[pm, 2:3] * wp + coef4[pm, 3:4] * torch.cat((e3[NL:], n3[2]), 1)
            mean = scatter_mean(torch.cat((wl[:, :3], wp[:, :3])), cm)
            wl[:, :3] -= mean[lm]; wp[:, :3] -= mean[pm]
        a, b = zl.clone(), zp.clone()
        j = [P(x) for x in n3] if jump else [None, None, None]
        _native.check(lib.dsb_ddpm_joint_inpaint_update(P(a), P(b), P(x0l), P(x0p), P(fl), P(fp), P(nx), P(nhl), P(nhp), *j,
                                                        P(coef4), P(lm), P(pm), NL, NP, B, A, R, stream))
        torch.cuda.synchronize()
        assert torch.allclose(a, wl, atol=3e-6, rtol=1e-5), float((a - wl).abs().max())
        assert torch.allclose(b, wp, atol=3e-6, rtol=1e-5), float((b - wp).abs().max())


def _joint_pair(T):
    eager, graph = _build_joint(T, 'cuda', native=True), _build_joint(T, 'cuda', native=True)
    eager.loop_engine, graph.loop_engine = 'eager', 'graph'
    return eager, graph


def test_joint_graph_sample_matches_eager_same_seed():
    """EnVariationalDiffusion.sample: captured joint reverse step (native denoiser + dsb_ddpm_joint_update) vs the
    reference-order eager loop with the same CUDA seed (three randn draws per step in the same order)."""
    T = 4
    eager, graph = _joint_pair(T)
    n_lig, n_poc = torch.tensor([6, 4]).cuda(), torch.tensor([14, 11]).cuda()
    outs = []
    for ddpm in (eager, graph):
        torch.manual_seed(31)
        outs.append(ddpm.sample(2, n_lig, n_poc, device='cuda'))
    a, b = outs
    assert graph._joint_cache and 'reverse' in next(iter(graph._joint_cache.values()))['graphs']
    scale = max(1.0, float(a[1][:, :3].abs().max()))
    assert torch.allclose(a[0][:, :3], b[0][:, :3], atol=2e-3 * scale), float((a[0] - b[0]).abs().max())
    assert torch.allclose(a[1][:, :3], b[1][:, :3], atol=2e-3 * scale), float((a[1] - b[1]).abs().max())
    com = scatter_mean(torch.cat((b[0][:, :3], b[1][:, :3])), torch.cat((b[2], b[3])))
    assert com.abs().max() < 5e-2 * scale


@pytest.mark.parametrize('case', ['joint_inpaint_T6_r2_j2', 'joint_inpaint_T4_r3_partial_pocket', 'joint_inpaint_T8_sub4_frames2'])
def test_joint_graph_inpaint_matches_eager_same_seed(case):
    """EnVariationalDiffusion.inpaint (RePaint schedule with jumps, partially free pocket, frames): captured iteration graphs vs
    the eager loop with the same seed.  The end point of a joint trajectory is sensitive, so the comparison uses a
    loose tolerance; the fused kernels themselves are checked exactly in test_fused_joint_kernels_match_torch_ops."""
    from ddpm_cases import JOINT_CASES, make_pocket_fixed
    spec = JOINT_CASES[case]
    eager, graph = _joint_pair(spec['T'])
    outs = []
    for ddpm in (eager, graph):
        lig, fixed = make_ligand(spec['n_lig'], spec['n_fixed'], device='cuda')
        pocket = make_pocket('cuda')
        pfix = make_pocket_fixed(spec, pocket).cuda()
        torch.manual_seed(spec['seed'])
        outs.append(ddpm.inpaint(lig, pocket, fixed, pfix, resamplings=spec['resamplings'], jump_length=spec['jump_length'],
                                 return_frames=spec['frames'], timesteps=spec['timesteps']))
    a, b = outs
    st = next(iter(graph._joint_cache.values()))
    assert 'inpaint' in st['graphs']
    assert a[0].shape == b[0].shape and all(torch.isfinite(o).all() for o in b[:2])
    scale = max(1.0, float(a[1][..., :3].abs().max()))
    assert torch.allclose(a[0][..., :3], b[0][..., :3], atol=2e-2 * scale), float((a[0] - b[0]).abs().max())
    assert torch.allclose(a[1][..., :3], b[1][..., :3], atol=2e-2 * scale), float((a[1] - b[1]).abs().max())

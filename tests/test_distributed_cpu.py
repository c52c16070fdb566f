"""CPU (gloo, world_size 2): the multi-GPU sharding logic — contiguous pocket split, per-rank seeds, final all_gather —
with the CPU oracle as denoiser.  The sharded result must equal the concatenation of the single-process runs of each shard."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddpm_cases import DDPM_CFG, HIST
from diffsbdd_b200 import synthetic as syn
from diffsbdd_b200.conditional_model import ConditionalDDPM
from diffsbdd_b200.distributed import shard_bounds, shard_pocket, sample_given_pocket_sharded
from oracle.cpu_denoiser import OracleDynamics

N_POCKET = [14, 9, 17]
N_LIG = [5, 7, 4]
T = 3


def _build():
    sd = syn.synthetic_state_dict(DDPM_CFG, 5)
    ddpm = ConditionalDDPM(dynamics=OracleDynamics(DDPM_CFG, sd), atom_nf=DDPM_CFG.atom_nf, residue_nf=DDPM_CFG.residue_nf,
                           n_dims=3, timesteps=T, noise_schedule='polynomial_2', noise_precision=5e-4, loss_type='l2',
                           norm_values=(1, 4), size_histogram=HIST)
    return ddpm.eval()


def _pocket():
    return syn.synthetic_pocket(DDPM_CFG, N_POCKET, seed=41, spread=2.0)


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    xh_all, sizes, local = sample_given_pocket_sharded(_build(), _pocket(), torch.tensor(N_LIG), base_seed=100)
    out[rank] = (xh_all.clone(), sizes.clone(), local[0].clone())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_shard_pocket_renumbers_graphs():
    p = _pocket()
    s = shard_pocket(p, 1, 3)
    assert s['size'].tolist() == N_POCKET[1:3] and s['mask'].min() == 0 and s['mask'].max() == 1
    assert s['x'].shape[0] == sum(N_POCKET[1:3])


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process_shards():
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.equal(out[0][0], out[1][0]), 'ranks disagree on the gathered result'
    # reference: each shard sampled alone with that rank's seed
    pieces = []
    for r in range(2):
        lo, hi = shard_bounds(3, 2, r)
        torch.manual_seed(100 + r)
        res = _build().sample_given_pocket(shard_pocket(_pocket(), lo, hi), torch.tensor(N_LIG)[lo:hi])
        pieces.append(res[0])
        assert torch.equal(out[r][2], res[0])
    assert torch.equal(out[0][0], torch.cat(pieces))
    assert out[0][0].shape[0] == sum(N_LIG)

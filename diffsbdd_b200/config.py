"""Hyper-parameter containers for the denoiser hot path.

The field names are the reference's own constructor / yaml names
(reference: equivariant_diffusion/dynamics.py:11-19, lightning_modules.py:137-173,
configs/crossdock_fullatom_cond.yml:30-52) so a reference user finds the same knobs.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict, replace
from typing import Optional


@dataclass(frozen=True)
class DynamicsConfig:
    """Constructor arguments of ``EGNNDynamics`` (reference dynamics.py:11-19)."""
    atom_nf: int = 10
    residue_nf: int = 10
    n_dims: int = 3
    joint_nf: int = 128
    hidden_nf: int = 256
    n_layers: int = 6
    attention: bool = True
    condition_time: bool = True
    tanh: bool = True
    mode: str = 'egnn_dynamics'
    norm_constant: float = 1
    inv_sublayers: int = 1
    sin_embedding: bool = False
    normalization_factor: float = 100
    aggregation_method: str = 'sum'
    update_pocket_coords: bool = False
    edge_cutoff_ligand: Optional[float] = None
    edge_cutoff_pocket: Optional[float] = 5.0
    edge_cutoff_interaction: Optional[float] = 5.0
    reflection_equivariant: bool = False
    edge_embedding_dim: Optional[int] = None

    def kwargs(self) -> dict:
        return asdict(self)

    def with_(self, **kw) -> "DynamicsConfig":
        return replace(self, **kw)

    @property
    def edge_feat_nf(self) -> int:
        """Width of the per-edge attribute vector (reference egnn_new.py:203-210): current and input-geometry d^2 (or their
        sinusoidal embeddings, 2 x 12 features, egnn_new.py:282-293) + the edge-type embedding."""
        return (24 if self.sin_embedding else 2) + (self.edge_embedding_dim or 0)


# BASELINE.json configs (dims from configs/crossdock_{fullatom,ca}_cond.yml:30-52,
# vocabulary sizes constants.py:170-173).
FULLATOM_COND = DynamicsConfig()
CA_COND = DynamicsConfig(residue_nf=20)
# joint models (configs/crossdock_fullatom_joint.yml:30-52 use hidden 128/5 layers; kept generic here)
FULLATOM_JOINT = DynamicsConfig(update_pocket_coords=True, joint_nf=32, hidden_nf=128, n_layers=5)
# config 1 of BASELINE.json: N=64, h_dim=256, 4 layers
CONFIG1 = DynamicsConfig(n_layers=4)


@dataclass(frozen=True)
class DiffusionConfig:
    """``diffusion_params`` of the reference yaml (configs/crossdock_fullatom_cond.yml:45-50)."""
    diffusion_steps: int = 500
    diffusion_noise_schedule: str = 'polynomial_2'
    diffusion_noise_precision: float = 5.0e-4
    diffusion_loss_type: str = 'l2'
    normalize_factors: tuple = (1, 4)

"""Analytical score network (reference: models/score_networks/analytical_score_network.py:28-298).

The exact sigma-normalised score when the atoms are isotropic Gaussian displacements (width ``sigma_d``) around given
equilibrium positions.  It needs no training and produces physically meaningful samples, which makes it the vehicle
that puts the wrapped-Gaussian score kernel on the sampling path (the EGNN never calls it) and that the
trajectory-level parity tests use.  Only the non-permutation-invariant form is on this path: the symmetrised form of
the reference sums over natoms! permutations and is an exploration tool.
"""
from dataclasses import dataclass
from typing import Any, AnyStr, Dict, List

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import (
    ScoreNetwork, ScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL, NOISE, NOISY_AXL_COMPOSITION


@dataclass(kw_only=True)
class AnalyticalScoreNetworkParameters(ScoreNetworkParameters):
    """Hyper-parameters of the analytical score network (reference :28-57)."""

    architecture: str = "analytical"
    number_of_atoms: int
    kmax: int  # lattice translations -kmax .. kmax are summed
    equilibrium_relative_coordinates: List[List[float]]
    sigma_d: float  # width of the data distribution
    use_permutation_invariance: bool = False

    def __post_init__(self):
        super().__post_init__()
        assert self.sigma_d > 0.0, "the sigma_d parameter should be positive."
        assert len(self.equilibrium_relative_coordinates) == self.number_of_atoms, \
            "There should be exactly one list of equilibrium coordinates per atom."
        for x in self.equilibrium_relative_coordinates:
            assert len(x) == self.spatial_dimension, \
                "The equilibrium coordinates should be consistent with the spatial dimension."


class AnalyticalScoreNetwork(ScoreNetwork):
    """Score of a sum of wrapped Gaussians centred on the equilibrium positions, evaluated in one CUDA kernel."""

    def __init__(self, hyper_params: AnalyticalScoreNetworkParameters):
        super().__init__(hyper_params)
        if hyper_params.use_permutation_invariance:
            raise NotImplementedError("the permutation-symmetrised analytical score (natoms! terms) is not part of the "
                                      "sm_100a sampling path")
        self.number_of_atomic_classes = hyper_params.num_atom_types + 1
        self.natoms = hyper_params.number_of_atoms
        self.nd = self.natoms * self.spatial_dimension
        self.kmax = hyper_params.kmax
        self.sigma_d = float(hyper_params.sigma_d)
        self.sigma_d_square = hyper_params.sigma_d ** 2
        self.use_permutation_invariance = False
        self.translations_k = torch.nn.Parameter(torch.arange(-self.kmax, self.kmax + 1), requires_grad=False)
        self.equilibrium_relative_coordinates = torch.tensor(hyper_params.equilibrium_relative_coordinates,
                                                             dtype=torch.float)
        self.all_x0 = torch.nn.Parameter(self.equilibrium_relative_coordinates.unsqueeze(0).clone(),
                                         requires_grad=False)

    def _forward_unchecked(self, batch: Dict[AnyStr, Any], conditional: bool = False) -> AXL:
        composition = batch[NOISY_AXL_COMPOSITION]
        device = _cabi.require_cuda(composition.X)
        x = _cabi.as_f32(composition.X)
        sigmas = _cabi.as_f32(batch[NOISE].to(device)).reshape(-1)
        batch_size, natoms, d = x.shape
        assert natoms == self.natoms, "the batch has a different number of atoms than the equilibrium configuration"
        x0 = _cabi.as_f32(self.all_x0[0]).to(device)
        scores = torch.empty_like(x)
        _cabi.check(_cabi.lib().b200md_analytical_score_f32(
            _cabi.ptr(x), _cabi.ptr(x0), _cabi.ptr(sigmas), self.sigma_d, batch_size, natoms, d, self.kmax,
            _cabi.ptr(scores), _cabi.stream_ptr(device)))
        logits = torch.zeros(batch_size, natoms, self.number_of_atomic_classes, dtype=torch.float32, device=device)
        logits[..., -1] = -torch.inf
        return AXL(A=logits, X=scores, L=torch.zeros_like(scores))

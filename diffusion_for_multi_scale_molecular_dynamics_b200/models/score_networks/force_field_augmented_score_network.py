"""Force-field augmented score network (reference: models/score_networks/force_field_augmented_score_network.py:19-236).

Wraps any score network and adds a repulsive pseudo-force ``F(r) = -2 s (r - r0) r_hat`` (for r < r0) to the coordinate
score, so that random moves cannot park two atoms on top of each other.  The reference materialises the per-image
adjacency, gathers displacements and scatter-adds one Cartesian component at a time; here the whole augmentation is one
kernel (``b200md_repulsive_pseudo_force``, csrc/structure.cu) that sweeps the 27 images of a structure staged in shared
memory and adds the relative-coordinate force to the wrapped network's score in the same pass.
"""
from dataclasses import dataclass
from typing import AnyStr, Dict, Optional

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import ScoreNetwork
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL, NOISY_AXL_COMPOSITION
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.neighbors import (AdjacencyInfo,
                                                                             get_periodic_adjacency_information)

MIN_BOX_SIZE = 1.0  # map_noisy_axl_lattice_parameters_to_unit_cell_vectors(..., min_box_size=1.0), reference :126,147,226


@dataclass(kw_only=True)
class ForceFieldParameters:
    """Force field parameters: phi(r) = strength * (r - radial_cutoff)^2 for r < radial_cutoff."""

    radial_cutoff: float  # Cutoff to the interaction, in Angstrom
    strength: float  # Strength of the repulsion

    def __post_init__(self):
        """Post init."""
        assert self.radial_cutoff > 0.0, "the radial cutoff should be greater than zero."
        assert self.strength > 0.0, "the repulsive strength should be greater than zero."


class ForceFieldAugmentedScoreNetwork(torch.nn.Module):
    """Force Field-Augmented Score Network."""

    def __init__(self, score_network: ScoreNetwork, force_field_parameters: ForceFieldParameters):
        """Init method.

        Args:
            score_network : a score network, to be augmented with a repulsive force.
            force_field_parameters : parameters for the repulsive force.
        """
        super().__init__()
        self._score_network = score_network
        self._force_field_parameters = force_field_parameters
        self._own_status: Optional[torch.Tensor] = None
        self._own_defer = False

    # ---- device-side status word, shared with the wrapped network when it has one ---------------------------------
    @property
    def defer_status_check(self) -> bool:
        inner = self._score_network
        return inner.defer_status_check if isinstance(inner, ScoreNetwork) else self._own_defer

    @defer_status_check.setter
    def defer_status_check(self, value: bool) -> None:
        if isinstance(self._score_network, ScoreNetwork):
            self._score_network.defer_status_check = value
        self._own_defer = value

    def device_status(self, device: torch.device) -> torch.Tensor:
        if isinstance(self._score_network, ScoreNetwork):
            return self._score_network.device_status(device)
        if self._own_status is None or self._own_status.device != device:
            self._own_status = torch.zeros(1, dtype=torch.int32, device=device)
        return self._own_status

    def raise_on_device_status(self) -> None:
        if isinstance(self._score_network, ScoreNetwork):
            self._score_network.raise_on_device_status()
        elif self._own_status is not None:
            _cabi.raise_on_status(self._own_status, AssertionError)

    # ---- forward -----------------------------------------------------------------------------------------------------
    def forward(self, batch: Dict[AnyStr, torch.Tensor], conditional: Optional[bool] = None) -> AXL:
        """Scores of the wrapped network with the repulsive pseudo-force added to the coordinate score."""
        defer = self.defer_status_check
        self.defer_status_check = True   # one status read for both kernels
        try:
            raw_scores = self._score_network(batch, conditional)
            updated_x = self._pseudo_force(batch, add_to=raw_scores.X)
        finally:
            self.defer_status_check = defer
        if not defer:
            self.raise_on_device_status()
        return AXL(A=raw_scores.A, X=updated_x, L=raw_scores.L)

    def get_relative_coordinates_pseudo_force(self, batch: Dict[AnyStr, torch.Tensor]) -> torch.Tensor:
        """Repulsive force in relative coordinates, [batch_size, natoms, spatial_dimension]."""
        forces = self._pseudo_force(batch, add_to=None)
        if not self.defer_status_check:
            self.raise_on_device_status()
        return forces

    def _pseudo_force(self, batch: Dict[AnyStr, torch.Tensor], add_to: Optional[torch.Tensor]) -> torch.Tensor:
        composition = batch[NOISY_AXL_COMPOSITION]
        device = _cabi.require_cuda(composition.X, composition.L)
        x = _cabi.as_f32(composition.X)
        lattice = _cabi.as_f32(composition.L)
        assert x.dim() == 3 and x.shape[-1] == 3 and lattice.shape == (x.shape[0], 6), \
            "the sm_100a pseudo-force kernel supports spatial_dimension == 3 only"
        base = _cabi.as_f32(add_to) if add_to is not None else None
        out = torch.empty_like(x)
        parameters = self._force_field_parameters
        _cabi.check(_cabi.lib().b200md_repulsive_pseudo_force(
            _cabi.ptr(x), _cabi.ptr(lattice), MIN_BOX_SIZE, float(parameters.radial_cutoff), float(parameters.strength),
            x.shape[0], x.shape[1], _cabi.ptr(base), _cabi.ptr(out), _cabi.ptr(self.device_status(device)),
            _cabi.stream_ptr(device)))
        return out

    # ---- the reference's intermediate, kept for inspection -----------------------------------------------------------
    def _get_adjacency_information(self, batch: Dict[AnyStr, torch.Tensor]) -> AdjacencyInfo:
        """Edges (per periodic image) within the interaction cutoff, in the box the force field uses."""
        composition = batch[NOISY_AXL_COMPOSITION]
        lattice = composition.L.clip(min=MIN_BOX_SIZE)
        basis_vectors = torch.diag_embed(lattice[:, :3])
        cartesian_positions = torch.matmul(composition.X, basis_vectors)
        return get_periodic_adjacency_information(cartesian_positions, basis_vectors,
                                                  radial_cutoff=self._force_field_parameters.radial_cutoff)

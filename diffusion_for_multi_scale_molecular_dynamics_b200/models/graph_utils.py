"""Graph-batched adjacency (reference: models/graph_utils.py:10-50)."""
from typing import Tuple

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.neighbors import _radius_graph


def get_adj_matrix(
    positions: torch.Tensor, basis_vectors: torch.Tensor, radial_cutoff: float = 4.0, spatial_dimension: int = 3
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(adjacency [2,E] with graph-batched node indices, shifts [E,3], node batch indices, edges per structure).

    The b * N index shift of the reference (utils/neighbors.py:381-390) is applied inside the fill kernel.
    """
    if spatial_dimension != 3:
        raise _cabi.B200MDError("the sm_100a radius graph supports spatial_dimension == 3 only")
    batch_size, number_of_atoms, _ = positions.shape
    src, dst, shifts, _, per_structure = _radius_graph(
        positions, basis_vectors, radial_cutoff, _cabi.GRAPH_PER_IMAGE, False, number_of_atoms, True)
    node_batch = torch.repeat_interleave(torch.arange(batch_size, device=positions.device), number_of_atoms)
    return torch.stack([src, dst]), shifts, node_batch, per_structure

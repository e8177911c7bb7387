"""Structure utilities of the sampling path (reference: utils/structure_utils.py:41-140).

``compute_distances_in_batch`` feeds the reference's structure Kolmogorov-Smirnov metric: every interatomic distance up
to a cutoff, over the 27 periodic images.  The reference builds a symbolic [B, 27, N, N] KeOps matrix and takes a
K-nearest reduction; here the image sweep of the radius graph (one CTA per structure, positions in shared memory)
counts, then writes the distances in the same order (``b200md_pair_distances_count`` / ``_fill``, csrc/structure.cu).
"""
from typing import List

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi


def compute_distances_in_batch(cartesian_positions: torch.Tensor, unit_cell: torch.Tensor,
                               max_distance: float) -> torch.Tensor:
    """Distances between atoms (and periodic images) of every structure of a batch, up to a cutoff.

    Args:
        cartesian_positions: atomic positions in Angstrom. (batch_size, n_atoms, spatial_dimension)
        unit_cell: lattice vectors. (batch_size, spatial_dimension, spatial_dimension)
        max_distance: cutoff distance

    Returns:
        1-D tensor with all the distances larger than 0 and not larger than max_distance, in the reference's order
        (structure, periodic image, source atom, ascending distance).
    """
    assert cartesian_positions.dim() == 3 and cartesian_positions.shape[-1] == 3, \
        "the sm_100a pair-distance kernel supports spatial_dimension == 3 only"
    device = _cabi.require_cuda(cartesian_positions, unit_cell)
    lib = _cabi.lib()
    pos = _cabi.as_f32(cartesian_positions)
    cell = _cabi.as_f32(unit_cell)
    batch_size, natom, _ = pos.shape
    rows = batch_size * 27 * natom
    counts = torch.empty(rows, dtype=torch.int32, device=device)
    per_structure = torch.empty(batch_size, dtype=torch.int64, device=device)
    stream = _cabi.stream_ptr(device)
    _cabi.check(lib.b200md_pair_distances_count(_cabi.ptr(pos), _cabi.ptr(cell), 0, float(max_distance), batch_size,
                                                natom, _cabi.ptr(counts), _cabi.ptr(per_structure), stream))
    inclusive = torch.cumsum(counts, dim=0, dtype=torch.int64)
    offsets = inclusive - counts
    total = int(inclusive[-1].item()) if rows > 0 else 0  # the reference synchronises here too (int(max_k_array.max()))
    distances = torch.empty(total, dtype=torch.float32, device=device)
    if total > 0:
        _cabi.check(lib.b200md_pair_distances_fill(_cabi.ptr(pos), _cabi.ptr(cell), 0, float(max_distance), batch_size,
                                                   natom, _cabi.ptr(offsets), _cabi.ptr(distances), stream))
    return distances


def get_orthogonal_basis_vectors(batch_size: int, cell_dimensions: List[float]) -> torch.Tensor:
    """Diagonal basis vectors repeated along the batch: [batch_size, spatial_dimension, spatial_dimension]."""
    basis_vectors = torch.diag(torch.Tensor(cell_dimensions)).unsqueeze(0).repeat(batch_size, 1, 1)
    return basis_vectors

"""Periodic neighbour search on sm_100a (reference: utils/neighbors.py:24-390).

The reference evaluates three KeOps reductions over a symbolic [B, 27, N, N] distance matrix and then
compacts with ``nonzero``.  Here one CTA per structure keeps the positions in shared memory, tests the 27
periodic images with the same fp32 rounding chain and compacts with warp ballots
(``b200md_radius_graph_count`` / ``b200md_radius_graph_fill``); see csrc/graph.cu.
"""
from collections import namedtuple
from typing import Tuple

import numpy as np
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi

INDEX_PADDING_VALUE = -1
POSITION_PADDING_VALUE = np.nan

AdjacencyInfo = namedtuple(
    "AdjacencyInfo",
    ["adjacency_matrix", "shifts", "edge_batch_indices", "node_batch_indices", "number_of_edges"],
)


def _radius_graph(
    positions: torch.Tensor,
    cell: torch.Tensor,
    radial_cutoff: float,
    mode: int,
    positions_are_relative: bool,
    node_index_offset: int,
    want_shifts: bool,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """count -> (host) exclusive scan -> fill.  Returns (src, dst, shifts, edge_batch, edges_per_structure)."""
    device = _cabi.require_cuda(positions, cell)
    lib = _cabi.lib()
    pos = _cabi.as_f32(positions)
    basis = _cabi.as_f32(cell)
    batch_size, natom, _ = pos.shape
    rows = batch_size * natom * (27 if mode == _cabi.GRAPH_PER_IMAGE else 1)
    counts = torch.empty(rows, dtype=torch.int32, device=device)
    per_structure = torch.empty(batch_size, dtype=torch.int64, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    stream = _cabi.stream_ptr(device)
    # unique pairs without shifts: ONE sweep -- the count pass keeps the hit bit row of every source atom and the fill
    # expands those rows; the other forms need the image (shifts) or the distances (per-image order) again and sweep twice
    one_sweep = mode == _cabi.GRAPH_UNIQUE_PAIRS and not want_shifts
    if one_sweep:
        hit_mask = torch.empty(batch_size * natom, (natom + 31) // 32, dtype=torch.int32, device=device)
        _cabi.check(lib.b200md_radius_graph_count_masked(
            _cabi.ptr(pos), _cabi.ptr(basis), int(positions_are_relative), float(radial_cutoff), batch_size, natom,
            _cabi.ptr(counts), _cabi.ptr(per_structure), _cabi.ptr(status), _cabi.ptr(hit_mask), stream))
    else:
        _cabi.check(lib.b200md_radius_graph_count(
            _cabi.ptr(pos), _cabi.ptr(basis), int(positions_are_relative), float(radial_cutoff), batch_size, natom, mode,
            _cabi.ptr(counts), _cabi.ptr(per_structure), _cabi.ptr(status), stream))
    # The reference also synchronises here (int(max_k_array.max()), neighbors.py:181, and the assert :110).
    _cabi.raise_on_status(status, AssertionError)
    inclusive = torch.cumsum(counts, dim=0, dtype=torch.int64)
    offsets = inclusive - counts
    total = int(inclusive[-1].item()) if rows > 0 else 0
    src = torch.empty(total, dtype=torch.int64, device=device)
    dst = torch.empty(total, dtype=torch.int64, device=device)
    shifts = torch.empty(total, 3, dtype=torch.float32, device=device) if want_shifts else None
    edge_batch = torch.empty(total, dtype=torch.int64, device=device)
    if total > 0 and one_sweep:
        _cabi.check(lib.b200md_radius_graph_fill_from_mask(
            _cabi.ptr(hit_mask), batch_size, natom, _cabi.ptr(offsets), int(node_index_offset), _cabi.ptr(src),
            _cabi.ptr(dst), _cabi.ptr(edge_batch), stream))
    elif total > 0:
        _cabi.check(lib.b200md_radius_graph_fill(
            _cabi.ptr(pos), _cabi.ptr(basis), int(positions_are_relative), float(radial_cutoff), batch_size, natom,
            mode, _cabi.ptr(offsets), int(node_index_offset), _cabi.ptr(src), _cabi.ptr(dst), _cabi.ptr(shifts),
            _cabi.ptr(edge_batch), stream))
    return src, dst, shifts, edge_batch, per_structure


def get_periodic_adjacency_information(
    cartesian_positions: torch.Tensor, basis_vectors: torch.Tensor, radial_cutoff: float, spatial_dimension: int = 3
) -> AdjacencyInfo:
    """All neighbour pairs within ``radial_cutoff`` under periodic boundary conditions (reference :36-224).

    Same contract as the reference: indices are local to each structure, one edge per (image, src, dst),
    ordered by (batch, image, src, ascending distance); ``AssertionError`` when the cutoff reaches beyond
    the first shell of images.
    """
    assert len(cartesian_positions.shape) == 3, "Wrong number of dimensions for relative_coordinates"
    assert len(basis_vectors.shape) == 3, "Wrong number of dimensions for basis_vectors"
    batch_size, max_natom, spatial_dimension_ = cartesian_positions.shape
    assert spatial_dimension_ == spatial_dimension, "Wrong spatial dimension for relative_coordinates "
    assert basis_vectors.shape == (batch_size, spatial_dimension, spatial_dimension), "Wrong shape for basis vectors"
    assert radial_cutoff > 0.0, "The radial cutoff should be greater than zero"
    if spatial_dimension != 3:
        raise _cabi.B200MDError("the sm_100a radius graph supports spatial_dimension == 3 only")
    src, dst, shifts, edge_batch, per_structure = _radius_graph(
        cartesian_positions, basis_vectors, radial_cutoff, _cabi.GRAPH_PER_IMAGE, False, 0, True)
    device = cartesian_positions.device
    return AdjacencyInfo(
        adjacency_matrix=torch.stack([src, dst]),
        shifts=shifts,
        edge_batch_indices=edge_batch,
        node_batch_indices=torch.repeat_interleave(torch.arange(batch_size, device=device), max_natom),
        number_of_edges=per_structure,
    )


def shift_adjacency_matrix_indices_for_graph_batching(
    adjacency_matrix: torch.Tensor, num_edges: torch.Tensor, number_of_atoms: int
) -> torch.Tensor:
    """Add b * number_of_atoms to the node indices of structure b (reference :381-390)."""
    offsets = torch.arange(len(num_edges), device=adjacency_matrix.device) * number_of_atoms
    return adjacency_matrix + torch.repeat_interleave(offsets, num_edges.to(adjacency_matrix.device)).repeat(2, 1)

"""Segment reductions and edge lists for EGNN (reference: models/egnn_utils.py:11-144)."""
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.neighbors import _radius_graph


def _segment_reduce(data: torch.Tensor, segment_ids: torch.Tensor, num_segments: int, mean: bool) -> torch.Tensor:
    device = _cabi.require_cuda(data, segment_ids)
    values = _cabi.as_f32(data)
    ids = _cabi.as_i64(segment_ids)
    n_elements, n_features = values.shape
    out = torch.empty(num_segments, n_features, dtype=torch.float32, device=device)
    counts = torch.empty(num_segments, dtype=torch.float32, device=device) if mean else None
    _cabi.check(_cabi.lib().b200md_segment_reduce(
        _cabi.ptr(values), _cabi.ptr(ids), n_elements, n_features, num_segments, int(mean), _cabi.ptr(out),
        _cabi.ptr(counts), _cabi.stream_ptr(device)))
    return out


def unsorted_segment_sum(data: torch.Tensor, segment_ids: torch.Tensor, num_segments: int) -> torch.Tensor:
    """out[s] = sum of data[e] over segment_ids[e] == s (reference :11-38)."""
    return _segment_reduce(data, segment_ids, num_segments, mean=False)


def unsorted_segment_mean(data: torch.Tensor, segment_ids: torch.Tensor, num_segments: int) -> torch.Tensor:
    """Segment mean with empty segments giving 0 (count clamped to 1) (reference :41-70)."""
    return _segment_reduce(data, segment_ids, num_segments, mean=True)


def get_edges_batch(n_nodes: int, batch_size: int, device=None) -> torch.Tensor:
    """Fully connected directed edges (x, y != x) of ``batch_size`` disjoint graphs, [E, 2] (reference :73-104).

    Built arithmetically on the device instead of with the reference's Python list comprehension.
    """
    src = torch.arange(n_nodes, device=device).repeat_interleave(n_nodes - 1)
    k = torch.arange(n_nodes - 1, device=device).repeat(n_nodes)
    dst = k + (k >= src).long()  # skip the diagonal
    local = torch.stack([src, dst], dim=1)
    offsets = (torch.arange(batch_size, device=device) * n_nodes).view(-1, 1, 1)
    return (local.unsqueeze(0) + offsets).reshape(-1, 2)


def get_edges_with_radial_cutoff(
    relative_coordinates: torch.Tensor,
    unit_cell: torch.Tensor,
    radial_cutoff: float = 4.0,
    drop_duplicate_edges: bool = True,
    spatial_dimension: int = 3,
) -> torch.Tensor:
    """[E, 2] int64 graph-batched edges within the cutoff (reference :107-144).

    With ``drop_duplicate_edges`` the result is the sorted, de-duplicated (src, dst) list that
    ``torch.unique(dim=1)`` yields in the reference -- produced directly in that order by the kernel.
    """
    if spatial_dimension != 3:
        raise _cabi.B200MDError("the sm_100a radius graph supports spatial_dimension == 3 only")
    number_of_atoms = relative_coordinates.shape[1]
    mode = _cabi.GRAPH_UNIQUE_PAIRS if drop_duplicate_edges else _cabi.GRAPH_PER_IMAGE
    src, dst, _, _, _ = _radius_graph(relative_coordinates, unit_cell, radial_cutoff, mode, True, number_of_atoms, False)
    return torch.stack([src, dst], dim=1)

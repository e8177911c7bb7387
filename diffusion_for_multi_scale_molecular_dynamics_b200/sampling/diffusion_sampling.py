"""Batch driver of the sampling path (reference: sampling/diffusion_sampling.py:16-73) and its multi-GPU shard point.

Samples are independent Markov chains, so the batch is partitioned across ranks with no communication inside the
trajectory: rank r of R draws the contiguous global sample range ``shard_range(number_of_samples, r, R)`` (the
same split the reference makes sequentially with ``torch.split``), the counter-based noise is keyed by the GLOBAL
sample index (``generator.sample_offset``), and one ``all_gather`` of (A, X, L) at the end assembles the batch on
every rank.  With one process this is exactly the reference's chunk loop.
"""
import logging
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.axl_generator import AXLGenerator, SamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL, AXL_COMPOSITION, CARTESIAN_POSITIONS
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import (
    get_positions_from_coordinates, map_lattice_parameters_to_unit_cell_vectors)

logger = logging.getLogger(__name__)


def shard_range(number_of_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of rank ``rank``: contiguous, sizes differ by at most one, earlier ranks take the remainder."""
    base, extra = divmod(number_of_samples, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def chunk_ranges(start: int, stop: int, chunk: int) -> List[Tuple[int, int]]:
    """Split [start, stop) like ``torch.split(torch.arange(n), chunk)`` (reference :44-46)."""
    return [(lo, min(lo + chunk, stop)) for lo in range(start, stop, chunk)]


def _all_gather_variable(local: torch.Tensor, counts: List[int], group) -> torch.Tensor:
    """all_gather of tensors whose first dimension differs per rank by at most one (pad to the maximum)."""
    longest = max(counts)
    padded = local
    if local.shape[0] < longest:
        pad = torch.zeros((longest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad])
    gathered = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(gathered, padded.contiguous(), group=group)
    return torch.cat([g[:n] for g, n in zip(gathered, counts)])


def create_batch_of_samples(generator: AXLGenerator, sampling_parameters: SamplingParameters, device: torch.device,
                            process_group: Optional["dist.ProcessGroup"] = None):
    """Draw ``sampling_parameters.number_of_samples`` samples in chunks of ``sample_batchsize``.

    Returns ``{CARTESIAN_POSITIONS, AXL_COMPOSITION}`` like the reference.  When ``torch.distributed`` is
    initialised with more than one rank the samples are sharded across ranks and gathered at the end.
    """
    logger.info("Creating a batch of samples")
    number_of_samples = sampling_parameters.number_of_samples
    chunk = sampling_parameters.sample_batchsize or number_of_samples
    world_size, rank = 1, 0
    if dist.is_available() and dist.is_initialized():
        world_size, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    start, stop = shard_range(number_of_samples, rank, world_size)

    atom_types, coordinates, lattices = [], [], []
    base_offset = getattr(generator, "sample_offset", 0)
    try:
        for lo, hi in chunk_ranges(start, stop, chunk):
            if hasattr(generator, "sample_offset"):
                generator.sample_offset = base_offset + lo
            sampled = generator.sample(hi - lo, device=device)
            atom_types.append(sampled.A)
            coordinates.append(sampled.X)
            lattices.append(sampled.L)
    finally:
        if hasattr(generator, "sample_offset"):
            generator.sample_offset = base_offset
    if atom_types:
        a, x, lattice = torch.concat(atom_types), torch.concat(coordinates), torch.concat(lattices)
    else:  # a rank with no work (more ranks than samples)
        natoms, d = sampling_parameters.number_of_atoms, sampling_parameters.spatial_dimension
        a = torch.zeros(0, natoms, dtype=torch.int64, device=device)
        x = torch.zeros(0, natoms, d, dtype=torch.float32, device=device)
        lattice = torch.zeros(0, int(d * (d + 1) / 2), dtype=torch.float32, device=device)

    if world_size > 1:
        counts = [shard_range(number_of_samples, r, world_size) for r in range(world_size)]
        counts = [hi - lo for lo, hi in counts]
        a = _all_gather_variable(a, counts, process_group)
        x = _all_gather_variable(x, counts, process_group)
        lattice = _all_gather_variable(lattice, counts, process_group)

    spatial_dimension = x.shape[-1]
    lattice[..., spatial_dimension:] = 0  # orthogonal boxes only, as in the reference (:60)
    basis_vectors = map_lattice_parameters_to_unit_cell_vectors(lattice)
    cartesian_positions = get_positions_from_coordinates(x, basis_vectors)
    return {CARTESIAN_POSITIONS: cartesian_positions, AXL_COMPOSITION: AXL(A=a, X=x, L=lattice)}

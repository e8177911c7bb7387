"""Relative <-> Cartesian helpers on the sampling path (reference: utils/basis_transformations.py).

``map_relative_coordinates_to_unit_cell`` is a CUDA kernel; the lattice-parameter bookkeeping functions are
shape-only host utilities used when configurations are assembled.
"""
from typing import Union

import numpy as np
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL


def get_number_of_lattice_parameters(spatial_dimension: int) -> int:
    """d (d + 1) / 2 independent lattice parameters (reference :172-174)."""
    return int(spatial_dimension * (spatial_dimension + 1) / 2)


def get_spatial_dimension_from_number_of_lattice_parameters(number_of_lattice_parameters: int) -> int:
    """Inverse of :func:`get_number_of_lattice_parameters` (reference :177-181)."""
    return int((-1 + np.sqrt(1 + 8 * number_of_lattice_parameters)) / 2)


def map_relative_coordinates_to_unit_cell(relative_coordinates: torch.Tensor) -> torch.Tensor:
    """Bring relative coordinates back into [0, 1) (reference :95-119).

    ``remainder(x, 1)`` followed by the ``== 1 -> 0`` fix-up for tiny negative inputs, in one kernel
    (``b200md_wrap_to_unit_cell``).  float32 CUDA tensors only.
    """
    _cabi.require_cuda(relative_coordinates)
    x = _cabi.as_f32(relative_coordinates)
    out = torch.empty_like(x)
    _cabi.check(_cabi.lib().b200md_wrap_to_unit_cell(_cabi.ptr(x), _cabi.ptr(out), x.numel(), _cabi.stream_ptr(x.device)))
    return out.view(relative_coordinates.shape)


def map_axl_composition_to_unit_cell(composition: AXL, device: torch.device) -> AXL:
    """AXL version of the wrap (reference :122-138)."""
    wrapped = map_relative_coordinates_to_unit_cell(composition.X.to(device))
    return AXL(A=composition.A, X=wrapped, L=composition.L)


def map_lattice_parameters_to_unit_cell_vectors(lattice_parameters: torch.Tensor) -> torch.Tensor:
    """Lattice parameters [..., d(d+1)/2] -> cell vectors [..., d, d]; orthogonal boxes only (reference :141-170)."""
    d = get_spatial_dimension_from_number_of_lattice_parameters(lattice_parameters.shape[-1])
    angles = lattice_parameters[..., d:]
    assert torch.allclose(angles, torch.zeros_like(angles)), "Only orthogonal boxes are supported for now."
    return torch.diag_embed(lattice_parameters[..., :d])


def map_unit_cell_to_lattice_parameters(
    unit_cell: Union[np.ndarray, torch.Tensor], engine: str = "torch"
) -> Union[np.ndarray, torch.Tensor]:
    """Cell vectors [..., d, d] -> lattice parameters (diagonal lengths then zeros) (reference :184-218)."""
    assert engine in ["torch", "numpy"], f"Mapping can be done for numpy or torch. Got {engine}."
    d = unit_cell.shape[-1]
    n_lp = get_number_of_lattice_parameters(d)
    if engine == "torch":
        out = torch.zeros(*unit_cell.shape[:-2], n_lp).to(unit_cell)
        out[..., :d] = torch.diagonal(unit_cell, dim1=-2, dim2=-1)
    else:
        out = np.zeros(n_lp)
        out[..., :d] = np.diag(unit_cell)
    return out


def map_numpy_unit_cell_to_lattice_parameters(unit_cell: np.ndarray) -> np.ndarray:
    return map_unit_cell_to_lattice_parameters(unit_cell, engine="numpy")


def get_positions_from_coordinates(relative_coordinates: torch.Tensor, basis_vectors: torch.Tensor) -> torch.Tensor:
    """p = x @ cell (reference :34-57).  Used once per sample() when results are converted for output."""
    return torch.matmul(relative_coordinates, basis_vectors)

"""Sampling dataclasses and the generator interface (reference: generators/axl_generator.py:13-95)."""
import warnings
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import List, Optional

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import \
    map_unit_cell_to_lattice_parameters


@dataclass(kw_only=True)
class SamplingParameters:
    """Hyper-parameters for diffusion sampling; field names/defaults are those of the reference (:13-67)."""

    algorithm: str
    spatial_dimension: int = 3
    num_atom_types: int  # excluding MASK
    number_of_atoms: int
    number_of_samples: int
    sample_batchsize: Optional[int] = None  # None -> one chunk of number_of_samples
    use_fixed_lattice_parameters: bool = False
    cell_dimensions: Optional[List[float]] = None
    record_samples: bool = False
    record_samples_corrector_steps: bool = False
    record_atom_type_update: bool = False

    def __post_init__(self):
        if not self.use_fixed_lattice_parameters:
            warnings.warn("Using diffusion on lattice parameters. This is experimental and not fully tested.")
            self.fixed_lattice_parameters = None
            return
        assert self.cell_dimensions is not None, \
            "If use_fixed_lattice_parameters is True, then cell_dimensions must be provided."
        cell = torch.tensor(self.cell_dimensions)
        if cell.dim() == 1:  # a list of lengths means an orthogonal box
            cell = torch.diag(cell)
        assert cell.dim() == 2, f"Provided cell_dimensions must be a 2D tensor. Got {cell.shape}."
        assert cell.shape[0] == cell.shape[1] == self.spatial_dimension, \
            "The cell_dimensions tensor must have shape [spatial_dimension, spatial_dimension]."
        self.fixed_lattice_parameters = map_unit_cell_to_lattice_parameters(cell)


class AXLGenerator(ABC):
    """Interface of AXL (atom types, relative coordinates, lattice) generators."""

    @abstractmethod
    def sample(self, number_of_samples: int, device: torch.device) -> AXL:
        """Draw ``number_of_samples`` configurations on ``device``."""

    @abstractmethod
    def initialize(self, number_of_samples: int, device: torch.device) -> AXL:
        """Draw the fully-noised starting configurations."""

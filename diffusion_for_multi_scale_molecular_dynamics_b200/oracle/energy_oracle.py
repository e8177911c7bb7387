"""Energy oracle base class (reference: oracle/energy_oracle.py:21-140).

The reference's oracles evaluate one configuration at a time on the host; the interface here is batched: a subclass
implements ``_compute_batch_energies`` on device tensors.  ``compute_oracle_energies`` accepts the same sample dictionary
(``AXL_COMPOSITION`` or the separate ``RELATIVE_COORDINATES`` / ``ATOM_TYPES`` / ``LATTICE_PARAMETERS`` fields).
"""
from dataclasses import dataclass
from typing import AnyStr, Dict, List

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200.data.element_types import ElementTypes
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (ATOM_TYPES, AXL_COMPOSITION,
                                                                       LATTICE_PARAMETERS, RELATIVE_COORDINATES)


@dataclass(kw_only=True)
class OracleParameters:
    """Oracle parameters."""

    name: str  # what kind of Oracle
    elements: List[str]  # unique elements


class EnergyOracle:
    """Energy oracle base class."""

    def __init__(self, oracle_parameters: OracleParameters, **kwargs):
        self._oracle_parameters = oracle_parameters
        self._element_types = ElementTypes(oracle_parameters.elements)

    def _compute_batch_energies(self, relative_coordinates: torch.Tensor, lattice_parameters: torch.Tensor,
                                atom_types: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("This method must be implemented")

    def compute_oracle_energies(self, samples: Dict[AnyStr, torch.Tensor]) -> torch.Tensor:
        """Energies [batch_size] (eV) of samples in the standard format (reference :46-140, energies only)."""
        assert LATTICE_PARAMETERS in samples or AXL_COMPOSITION in samples, \
            f"the field '{LATTICE_PARAMETERS}' or '{AXL_COMPOSITION}' must be present in the sample dictionary"
        assert AXL_COMPOSITION in samples or ATOM_TYPES in samples, \
            f"the field '{AXL_COMPOSITION}' or '{ATOM_TYPES}' must be present in the sample dictionary"
        relative_coordinates = samples[RELATIVE_COORDINATES] if RELATIVE_COORDINATES in samples \
            else samples[AXL_COMPOSITION].X
        lattice_parameters = samples[LATTICE_PARAMETERS] if LATTICE_PARAMETERS in samples else samples[AXL_COMPOSITION].L
        atom_types = samples[ATOM_TYPES] if ATOM_TYPES in samples else samples[AXL_COMPOSITION].A
        return self._compute_batch_energies(torch.as_tensor(relative_coordinates), torch.as_tensor(lattice_parameters),
                                            torch.as_tensor(atom_types))

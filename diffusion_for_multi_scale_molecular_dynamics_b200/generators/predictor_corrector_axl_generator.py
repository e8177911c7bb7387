"""Predictor-corrector driver (reference: generators/predictor_corrector_axl_generator.py:21-204).

``sample_from_noisy_composition`` keeps the reference's loop order -- for i = start-1 .. end: predictor(i+1) then
``number_of_corrector_steps`` x corrector(i) -- and is the generic (step-by-step) driver.  Subclasses that can
run the whole trajectory in one C-ABI call override ``_run_fused`` (see LangevinGenerator).
"""
import logging
from abc import abstractmethod
from dataclasses import dataclass
from typing import Optional

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.axl_generator import AXLGenerator, SamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import (
    FullRandomTrajectoryInitializer, TrajectoryInitializer, TrajectoryInitializerParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import \
    get_number_of_lattice_parameters

logger = logging.getLogger(__name__)


@dataclass(kw_only=True)
class PredictorCorrectorSamplingParameters(SamplingParameters):
    """Hyper-parameters of the predictor-corrector sampler (reference :21-30)."""

    algorithm: str = "predictor_corrector"
    number_of_corrector_steps: int = 1
    small_epsilon: float = 1e-8
    one_atom_type_transition_per_step: bool = True
    atom_type_greedy_sampling: bool = True
    atom_type_transition_in_corrector: bool = False


class PredictorCorrectorAXLGenerator(AXLGenerator):
    """Interface + time loop of predictor-corrector AXL generators."""

    def __init__(
        self,
        number_of_discretization_steps: int,
        number_of_corrector_steps: int,
        spatial_dimension: int,
        num_atom_types: int,
        number_of_atoms: int,
        use_fixed_lattice_parameters: bool = False,
        fixed_lattice_parameters: Optional[torch.Tensor] = None,
        trajectory_initializer: Optional[TrajectoryInitializer] = None,
        **kwargs,
    ):
        assert number_of_discretization_steps > 1, "The number of discretization steps should be larger than one"
        assert number_of_corrector_steps >= 0, "The number of corrector steps should be non-negative"
        self.number_of_discretization_steps = number_of_discretization_steps
        self.number_of_corrector_steps = number_of_corrector_steps
        self.spatial_dimension = spatial_dimension
        self.num_classes = num_atom_types + 1  # + MASK
        self.num_lattice_parameters = get_number_of_lattice_parameters(spatial_dimension)
        if trajectory_initializer is None:
            trajectory_initializer = FullRandomTrajectoryInitializer(TrajectoryInitializerParameters(
                spatial_dimension=spatial_dimension, num_atom_types=num_atom_types, number_of_atoms=number_of_atoms,
                use_fixed_lattice_parameters=use_fixed_lattice_parameters,
                fixed_lattice_parameters=fixed_lattice_parameters))
        self.trajectory_initializer = trajectory_initializer

    def initialize(self, number_of_samples: int, device: torch.device) -> AXL:
        return self.trajectory_initializer.initialize(number_of_samples, device)

    def sample(self, number_of_samples: int, device: torch.device) -> AXL:
        """Draw samples with the PC algorithm (reference :81-110)."""
        start_composition = self.initialize(number_of_samples, device)
        start = self.trajectory_initializer.create_start_time_step_index(self.number_of_discretization_steps)
        end = self.trajectory_initializer.create_end_time_step_index()
        return self.sample_from_noisy_composition(starting_noisy_composition=start_composition,
                                                  starting_step_index=start, ending_step_index=end)

    def sample_from_noisy_composition(self, starting_noisy_composition: AXL, starting_step_index: int,
                                      ending_step_index: int) -> AXL:
        """Denoise from time index ``starting_step_index`` down to ``ending_step_index`` (reference :112-161)."""
        assert starting_step_index > ending_step_index, \
            "It is nonsensical for starting_step_index to be smaller or equal to ending_step_index."
        assert starting_step_index > 0, "Starting step should be larger than zero."
        assert ending_step_index >= 0, "ending step should be larger or equal to zero."
        fused = self._run_fused(starting_noisy_composition, starting_step_index, ending_step_index)
        if fused is not None:
            return fused
        composition_ip1 = starting_noisy_composition
        forces = torch.zeros_like(composition_ip1.X)
        composition_i = composition_ip1
        for i in range(starting_step_index - 1, max(ending_step_index, 0) - 1, -1):
            composition_i = self.predictor_step(composition_ip1, i + 1, forces)
            for _ in range(self.number_of_corrector_steps):
                composition_i = self.corrector_step(composition_i, i, forces)
            composition_ip1 = composition_i
        self._end_of_trajectory()
        return composition_i

    def _run_fused(self, starting_noisy_composition: AXL, starting_step_index: int,
                   ending_step_index: int) -> Optional[AXL]:
        """Hook: run the whole loop natively and return the result, or None to use the step-by-step loop."""
        return None

    def _end_of_trajectory(self) -> None:
        """Hook: called once after the step-by-step loop (deferred device-side checks)."""

    @abstractmethod
    def predictor_step(self, composition_ip1: AXL, ip1: int, cartesian_forces: torch.Tensor) -> AXL:
        """One predictor step from time index ``ip1`` to ``ip1 - 1``."""

    @abstractmethod
    def corrector_step(self, composition_i: AXL, i: int, cartesian_forces: torch.Tensor) -> AXL:
        """One corrector step at time index ``i``."""

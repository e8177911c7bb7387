"""Starting compositions of a sampling trajectory (reference: generators/trajectory_initializer.py:14-214).

The fully random initial coordinates are drawn on the device by the counter-based generator of the extension
(``b200md_draw_uniform``, draw id INIT_UNIFORM) keyed by the global sample index, so that a batch sharded over
several GPUs starts from the same configurations as the unsharded batch.
"""
import os
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Optional, Union

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.axl_generator import SamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL, NOISY_AXL_COMPOSITION
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import \
    get_number_of_lattice_parameters


@dataclass(kw_only=True)
class TrajectoryInitializerParameters:
    """Parameters for trajectory initialization (reference :14-44)."""

    spatial_dimension: int = 3
    num_atom_types: int
    use_fixed_lattice_parameters: bool = False
    fixed_lattice_parameters: Optional[torch.Tensor] = None
    number_of_atoms: int
    path_to_starting_configuration_data_pickle: Optional[str] = None

    def __post_init__(self):
        if self.use_fixed_lattice_parameters:
            assert self.fixed_lattice_parameters is not None, \
                "If use_fixed_lattice_parameters is True, then fixed_lattice_parameters must be provided."
            expected = get_number_of_lattice_parameters(self.spatial_dimension)
            assert self.fixed_lattice_parameters.shape[0] == expected, (
                "The fixed_lattice_parameters tensor must have shape [spatial_dimension * (spatial_dimension + 1) / 2]."
                f"Got {self.fixed_lattice_parameters.shape}.")
        else:
            assert self.fixed_lattice_parameters is None, \
                "fixed_lattice_parameters must be None if use_fixed_lattice_parameters is False."


class TrajectoryInitializer(ABC):
    """Creates the first composition of a trajectory and names its first / last time indices."""

    def __init__(self, trajectory_initializer_parameters: TrajectoryInitializerParameters) -> None:
        p = trajectory_initializer_parameters
        self.trajectory_initializer_parameters = p
        self.spatial_dimension = p.spatial_dimension
        self.number_of_atoms = p.number_of_atoms
        self.masked_atom_type_index = p.num_atom_types
        self.num_lattice_parameters = get_number_of_lattice_parameters(p.spatial_dimension)
        self.use_fixed_lattice_parameters = p.use_fixed_lattice_parameters
        self.fixed_lattice_parameters = p.fixed_lattice_parameters
        # set by the owning generator before initialize(): keys of the counter-based draws
        self.seed = 0
        self.sample_offset = 0

    @abstractmethod
    def initialize(self, number_of_samples: int, device: torch.device) -> AXL:
        pass

    @abstractmethod
    def create_start_time_step_index(self, number_of_discretization_steps: int) -> int:
        pass

    @abstractmethod
    def create_end_time_step_index(self) -> int:
        pass


class FullRandomTrajectoryInitializer(TrajectoryInitializer):
    """All atoms MASK, coordinates uniform in [0,1), lattice fixed or N(0,1) (reference :94-131)."""

    def initialize(self, number_of_samples: int, device: torch.device) -> AXL:
        device = torch.device(device)
        if device.type != "cuda":
            raise _cabi.B200MDError("trajectories are initialised on a CUDA device (counter-based draw kernel); "
                                    f"got '{device}'. There is no CPU path.")
        lib = _cabi.lib()
        atom_types = torch.full((number_of_samples, self.number_of_atoms), self.masked_atom_type_index,
                                dtype=torch.int64, device=device)
        x = torch.empty(number_of_samples, self.number_of_atoms, self.spatial_dimension, dtype=torch.float32,
                        device=device)
        _cabi.check(lib.b200md_draw_uniform(_cabi.ptr(x), number_of_samples, self.number_of_atoms,
                                            self.spatial_dimension, self.seed, 0, _cabi.DRAW_INIT_UNIFORM,
                                            self.sample_offset, _cabi.stream_ptr(device)))
        if self.use_fixed_lattice_parameters:
            lattice = self.fixed_lattice_parameters.to(device=device, dtype=torch.float32).repeat(number_of_samples, 1)
        else:
            lattice = torch.empty(number_of_samples, self.num_lattice_parameters, dtype=torch.float32, device=device)
            _cabi.check(lib.b200md_draw_normal(_cabi.ptr(lattice), number_of_samples, 1, self.num_lattice_parameters,
                                               self.seed, 0, _cabi.DRAW_INIT_LATTICE, self.sample_offset,
                                               _cabi.stream_ptr(device)))
        return AXL(A=atom_types, X=x, L=lattice)

    def create_start_time_step_index(self, number_of_discretization_steps: int) -> int:
        return number_of_discretization_steps

    def create_end_time_step_index(self) -> int:
        return 0


class StartFromGivenConfigurationTrajectoryInitializer(TrajectoryInitializer):
    """Restart from a pickled ``{noisy_axl, start_time_step_index}`` (reference :134-186)."""

    def __init__(self, trajectory_initializer_parameters: TrajectoryInitializerParameters) -> None:
        super().__init__(trajectory_initializer_parameters)
        self.start_time_step_index, self.noisy_starting_composition = self._read_pickle_data(
            trajectory_initializer_parameters.path_to_starting_configuration_data_pickle)

    def _read_pickle_data(self, path_to_starting_configuration_data_pickle: str):
        assert os.path.isfile(path_to_starting_configuration_data_pickle), \
            f"The file {path_to_starting_configuration_data_pickle} does not exist. Review input."
        data = torch.load(path_to_starting_configuration_data_pickle, weights_only=False)
        return data["start_time_step_index"], data[NOISY_AXL_COMPOSITION]

    def initialize(self, number_of_samples: int, device: torch.device) -> AXL:
        start = self.noisy_starting_composition
        assert number_of_samples == start.X.shape[0], (
            "The number of samples requested is inconsistent with the number of starting configurations in the data "
            "pickle. Something is probably inconsistent: stopping here, review inputs.")
        return AXL(A=start.A.to(device), X=start.X.to(device), L=start.L.to(device))

    def create_start_time_step_index(self, number_of_discretization_steps: int) -> int:
        return self.start_time_step_index

    def create_end_time_step_index(self) -> int:
        return 0


def instantiate_trajectory_initializer(
    sampling_parameters: SamplingParameters,
    path_to_starting_configuration_data_pickle: Union[str, None] = None,
) -> TrajectoryInitializer:
    """Factory (reference :189-214)."""
    params = TrajectoryInitializerParameters(
        spatial_dimension=sampling_parameters.spatial_dimension,
        num_atom_types=sampling_parameters.num_atom_types,
        number_of_atoms=sampling_parameters.number_of_atoms,
        use_fixed_lattice_parameters=sampling_parameters.use_fixed_lattice_parameters,
        fixed_lattice_parameters=sampling_parameters.fixed_lattice_parameters,
        path_to_starting_configuration_data_pickle=path_to_starting_configuration_data_pickle,
    )
    if path_to_starting_configuration_data_pickle:
        return StartFromGivenConfigurationTrajectoryInitializer(params)
    return FullRandomTrajectoryInitializer(params)

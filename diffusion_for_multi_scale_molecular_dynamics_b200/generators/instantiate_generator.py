"""Generator factory (reference: generators/instantiate_generator.py:25-82) for the predictor-corrector branch."""
from typing import Optional

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.axl_generator import SamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.constrained_langevin_generator import \
    ConstrainedLangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.sampling_constraint import SamplingConstraint
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import TrajectoryInitializer
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import ScoreNetwork
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters


def instantiate_generator(
    sampling_parameters: SamplingParameters,
    noise_parameters: NoiseParameters,
    axl_network: ScoreNetwork,
    trajectory_initializer: Optional[TrajectoryInitializer] = None,
    sampling_constraints: Optional[SamplingConstraint] = None,
):
    """Build the generator named by ``sampling_parameters.algorithm``; only "predictor_corrector" is on this path."""
    if sampling_parameters.algorithm != "predictor_corrector":
        raise NotImplementedError(
            f"Generator type {sampling_parameters.algorithm} is not part of the sm_100a sampling path "
            "(predictor_corrector only).")
    if sampling_constraints is None:
        return LangevinGenerator(noise_parameters=noise_parameters, sampling_parameters=sampling_parameters,
                                 axl_network=axl_network, trajectory_initializer=trajectory_initializer)
    return ConstrainedLangevinGenerator(noise_parameters=noise_parameters, sampling_parameters=sampling_parameters,
                                        axl_network=axl_network, sampling_constraints=sampling_constraints,
                                        trajectory_initializer=trajectory_initializer)

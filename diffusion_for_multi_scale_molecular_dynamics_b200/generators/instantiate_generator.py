"""Generator factory (reference: generators/instantiate_generator.py:25-82) for the predictor-corrector branch."""
from typing import Optional

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.adaptive_corrector import AdaptiveCorrectorGenerator
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
    """Build the generator named by ``sampling_parameters.algorithm``: "predictor_corrector" and "adaptive_corrector"
    are on this path; the ODE / SDE solvers of the reference are not."""
    assert sampling_parameters.algorithm in ["ode", "sde", "predictor_corrector", "adaptive_corrector"], \
        "Unknown algorithm. Possible choices are 'ode', 'sde', 'predictor_corrector' and 'adaptive_corrector'"
    if sampling_constraints is not None:
        assert sampling_parameters.algorithm == "predictor_corrector", \
            "Only the 'predictor_corrector' scheme supports sampling constraints."
        return ConstrainedLangevinGenerator(noise_parameters=noise_parameters, sampling_parameters=sampling_parameters,
                                            axl_network=axl_network, sampling_constraints=sampling_constraints,
                                            trajectory_initializer=trajectory_initializer)
    if sampling_parameters.algorithm == "predictor_corrector":
        return LangevinGenerator(noise_parameters=noise_parameters, sampling_parameters=sampling_parameters,
                                 axl_network=axl_network, trajectory_initializer=trajectory_initializer)
    if sampling_parameters.algorithm == "adaptive_corrector":
        return AdaptiveCorrectorGenerator(noise_parameters=noise_parameters, sampling_parameters=sampling_parameters,
                                          axl_network=axl_network, trajectory_initializer=trajectory_initializer)
    raise NotImplementedError(
        f"Generator type {sampling_parameters.algorithm} is not part of the sm_100a sampling path "
        "(predictor_corrector and adaptive_corrector only).")

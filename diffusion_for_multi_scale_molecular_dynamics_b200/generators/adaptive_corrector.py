"""Corrector-only Langevin sampler with an adaptive step size (reference: generators/adaptive_corrector.py:17-148).

Song et al. 2021, "Score-based generative modeling through stochastic differential equations": the predictor leaves
coordinates and lattice untouched; each corrector step uses

    eps_i = 2 (r ||z||_2 / ||s(x_i, t_i)||_2)^2

with the norms averaged over the batch.  The two batch-wide norms are reduced on the device
(``b200md_adaptive_corrector_step_size``: per-sample partial norms, then one CTA, fixed order) and the update kernel reads
the resulting weights from device memory, so a corrector step has no host synchronisation -- the reference's
``torch.linalg.norm(...).mean()`` chain costs six small launches and the generic update path a dozen more.
"""
from typing import Optional

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import TrajectoryInitializer
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import ScoreNetwork
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters


class AdaptiveCorrectorGenerator(LangevinGenerator):
    """Langevin Dynamics Generator using only a corrector step with adaptive step size for relative coordinates."""

    def __init__(self, noise_parameters: NoiseParameters, sampling_parameters: PredictorCorrectorSamplingParameters,
                 axl_network: ScoreNetwork, trajectory_initializer: Optional[TrajectoryInitializer] = None):
        """Init method."""
        super().__init__(noise_parameters=noise_parameters, sampling_parameters=sampling_parameters,
                         axl_network=axl_network, trajectory_initializer=trajectory_initializer)
        self.corrector_r = noise_parameters.corrector_r

    def _relative_coordinates_update_predictor_step(self, relative_coordinates: torch.Tensor,
                                                    sigma_normalized_scores: torch.Tensor, sigma_i, score_weight,
                                                    gaussian_noise_weight, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Do not update the relative coordinates in the predictor."""
        return relative_coordinates

    def _lattice_parameters_update_predictor_step(self, lattice_parameters: torch.Tensor,
                                                  sigma_normalized_scores: torch.Tensor, sigma_n_i, score_weight,
                                                  gaussian_noise_weight, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Do not update the lattice parameters in the predictor."""
        return lattice_parameters

    def _get_coordinates_corrector_step_size(self, index_i: int, sigma_i, model_predictions_i: torch.Tensor,
                                             z: torch.Tensor) -> torch.Tensor:
        """Adaptive step size for the relative coordinates: norms over (atoms, space) per sample."""
        return self._generic_corrector_step_size(index_i, sigma_i, model_predictions_i, z, n_dim=3)

    def _get_lattice_parameters_corrector_step_size(self, index_i: int, sigma_n_i, model_predictions_i: torch.Tensor,
                                                    z: torch.Tensor) -> torch.Tensor:
        """Adaptive step size for the lattice parameters: norms over the last dimension."""
        return self._generic_corrector_step_size(index_i, sigma_n_i, model_predictions_i, z, n_dim=2)

    def _generic_corrector_step_size(self, index_i: int, sigma_n_i, model_predictions_i: torch.Tensor, z: torch.Tensor,
                                     n_dim: int) -> torch.Tensor:
        r"""eps_i = 2 (r ||z||_2 / ||s(x_i, t_i)||_2)^2 as a device tensor of shape (1, 1, 1) or (1, 1).

        ``model_predictions_i`` is the sigma-normalised score sigma * s, hence the division of its norm by sigma
        (reference :130-134); the score norm is clipped from below by ``small_epsilon`` (:143).
        """
        device = _cabi.require_cuda(model_predictions_i, z)
        scores = _cabi.as_f32(model_predictions_i)
        noise = _cabi.as_f32(z.to(device))
        assert scores.shape == noise.shape, "The score and the noise must have the same shape."
        number_of_samples = scores.shape[0]
        units, d = (scores.shape[1], scores.shape[2]) if n_dim == 3 else (1, scores.shape[1])
        partial = torch.empty(2 * number_of_samples, dtype=torch.float32, device=device)
        step = torch.empty(2, dtype=torch.float32, device=device)
        sigma = float(torch.as_tensor(sigma_n_i, dtype=torch.float32).reshape(-1)[0]) if not _is_device(sigma_n_i) \
            else float(sigma_n_i.reshape(-1)[0].item())
        _cabi.check(_cabi.lib().b200md_adaptive_corrector_step_size(
            _cabi.ptr(scores), _cabi.ptr(noise), number_of_samples, units, d, sigma, float(self.corrector_r),
            float(self.small_epsilon), _cabi.ptr(partial), _cabi.ptr(step), _cabi.stream_ptr(device)))
        return step[0:1].view(1, 1, 1) if n_dim == 3 else step[0:1].view(1, 1)


def _is_device(value) -> bool:
    return torch.is_tensor(value) and value.is_cuda

"""Variance scheduler facade (reference: noise_schedulers/exploding_variance.py:9-64)."""
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.sigma_calculator import \
    instantiate_sigma_calculator


class VarianceScheduler(torch.nn.Module):
    """sigma(t), d sigma/dt and g^2(t) = d sigma^2/dt = 2 sigma sigma'."""

    def __init__(self, noise_parameters: NoiseParameters):
        super().__init__()
        self.sigma_calculator = instantiate_sigma_calculator(
            noise_parameters.sigma_min, noise_parameters.sigma_max, noise_parameters.schedule_type
        )

    def get_sigma(self, times: torch.Tensor) -> torch.Tensor:
        return self.sigma_calculator.get_sigma(times)

    def get_sigma_time_derivative(self, times: torch.Tensor) -> torch.Tensor:
        return self.sigma_calculator.get_sigma_time_derivative(times)

    def get_g_squared(self, times: torch.Tensor) -> torch.Tensor:
        return 2.0 * self.get_sigma(times) * self.get_sigma_time_derivative(times)

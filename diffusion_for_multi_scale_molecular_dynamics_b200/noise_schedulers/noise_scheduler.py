"""Noise tables for sampling (reference: noise_schedulers/noise_scheduler.py:11-27,112-378).

All tables are tiny (T <= a few thousand entries) and are built ONCE on the host with float32 torch ops in
the reference's operation order, so they are bit-identical to the reference (tests/golden pins this).  The
sampling kernels receive the per-step scalars as launch arguments and the [T,C,C] transition tables as
device pointers.

Index conventions (reference docstring :60-108): arrays are indexed 0..T-1 for i = 1..T, except epsilon /
sqrt_2_epsilon which are indexed by i = 0..T-1 directly (corrector steps), with t_0 = 0, sigma_0 = sigma_min.
"""
from collections import namedtuple
from typing import Tuple

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.exploding_variance import VarianceScheduler
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters

Noise = namedtuple(
    "Noise",
    ["time", "sigma", "sigma_squared", "g", "g_squared", "beta", "alpha_bar", "q_matrix", "q_bar_matrix",
     "q_bar_tm1_matrix", "indices"],
)
LangevinDynamics = namedtuple("LangevinDynamics", ["epsilon", "sqrt_2_epsilon"])


def _frozen(t: torch.Tensor) -> torch.nn.Parameter:
    return torch.nn.Parameter(t, requires_grad=False)


class NoiseScheduler(torch.nn.Module):
    """Builds and serves the time / sigma / g / epsilon / D3PM transition tables."""

    def __init__(self, noise_parameters: NoiseParameters, num_classes: int):
        super().__init__()
        self.noise_parameters = noise_parameters
        self.num_classes = num_classes
        self._exploding_variance = VarianceScheduler(noise_parameters)
        T = noise_parameters.total_time_steps

        times = self._get_time_array(noise_parameters)
        self._time_array = _frozen(times)
        self._sigma_array = _frozen(self._exploding_variance.get_sigma(times))
        self._sigma_squared_array = _frozen(self._sigma_array**2)
        self._g_squared_array = _frozen(
            self._create_discretized_g_squared_array(self._sigma_squared_array, noise_parameters.sigma_min)
        )
        self._g_array = _frozen(torch.sqrt(self._g_squared_array))
        self._epsilon_array = _frozen(self._create_epsilon_array(noise_parameters, self._sigma_squared_array))
        self._sqrt_two_epsilon_array = _frozen(torch.sqrt(2.0 * self._epsilon_array))
        self._maximum_random_index = _frozen(torch.tensor(T - 1))
        self._minimum_random_index = _frozen(torch.tensor(0))
        self._beta_array = _frozen(self._create_beta_array(T))
        self._alpha_bar_array = _frozen(self._create_alpha_bar_array(self._beta_array))
        self._q_matrix_array = _frozen(self._create_q_matrix_array(self._beta_array, num_classes))
        self._q_bar_matrix_array = _frozen(self._create_q_bar_matrix_array(self._q_matrix_array))
        self._q_bar_tm1_matrix_array = _frozen(self._create_q_bar_tm1_matrix_array(self._q_bar_matrix_array))

    # ---- table builders -------------------------------------------------------------------------
    @staticmethod
    def _get_time_array(noise_parameters: NoiseParameters) -> torch.Tensor:
        return torch.linspace(noise_parameters.time_delta, 1.0, noise_parameters.total_time_steps)

    @staticmethod
    def _create_discretized_g_squared_array(sigma_squared_array: torch.Tensor, sigma_min: float) -> torch.Tensor:
        # g^2_i = sigma^2_i - sigma^2_{i-1}, with sigma_0 = sigma_min
        head = torch.tensor([sigma_squared_array[0] - sigma_min**2])
        return torch.cat([head, sigma_squared_array[1:] - sigma_squared_array[:-1]])

    @staticmethod
    def _create_epsilon_array(noise_parameters: NoiseParameters, sigma_squared_array: torch.Tensor) -> torch.Tensor:
        # eps_i = 0.5 * epsilon * sigma_i^2 / sigma_1^2 for i = 0..T-1 (array shifted by one, sigma_0 = sigma_min)
        eps = noise_parameters.corrector_step_epsilon
        first = sigma_squared_array[0]
        head = torch.tensor([0.5 * eps * noise_parameters.sigma_min**2 / first])
        return torch.cat([head, 0.5 * eps * sigma_squared_array[:-1] / first])

    @staticmethod
    def _create_beta_array(num_time_steps: int) -> torch.Tensor:
        return 1.0 / (num_time_steps - torch.arange(1, num_time_steps + 1) + 1)

    @staticmethod
    def _create_alpha_bar_array(beta_array: torch.Tensor) -> torch.Tensor:
        return torch.cumprod(1 - beta_array, 0)

    @staticmethod
    def _create_q_matrix_array(beta_array: torch.Tensor, num_classes: int) -> torch.Tensor:
        # Q_t = (1 - beta_t) I + beta_t 1 e_MASK^T   (absorbing-state D3PM, MASK = last class)
        beta = beta_array.unsqueeze(-1).unsqueeze(-1)
        mask_row = torch.nn.functional.one_hot(torch.LongTensor([num_classes - 1]), num_classes=num_classes).squeeze(0)
        q = (1 - beta) * torch.eye(num_classes)
        q += beta * torch.outer(torch.ones(num_classes), mask_row)
        return q

    @staticmethod
    def _create_q_bar_matrix_array(q_matrix_array: torch.Tensor) -> torch.Tensor:
        q_bar = torch.empty_like(q_matrix_array)
        q_bar[0] = q_matrix_array[0]
        for i in range(1, q_matrix_array.size(0)):
            q_bar[i] = torch.matmul(q_bar[i - 1], q_matrix_array[i])
        return q_bar

    @staticmethod
    def _create_q_bar_tm1_matrix_array(q_bar_matrix_array: torch.Tensor) -> torch.Tensor:
        identity = torch.eye(q_bar_matrix_array.size(-1)).unsqueeze(0)
        return torch.cat((identity, q_bar_matrix_array[:-1]), dim=0)

    # ---- accessors ----------------------------------------------------------------------------------
    def _get_random_time_step_indices(self, shape: Tuple[int]) -> torch.Tensor:
        return torch.randint(
            self._minimum_random_index, self._maximum_random_index + 1, size=shape,
            device=self._minimum_random_index.device,
        )

    def get_random_noise_sample(self, batch_size: int) -> Noise:
        return self.get_noise_from_indices(self._get_random_time_step_indices((batch_size,)))

    def get_noise_from_indices(self, indices: torch.Tensor) -> Noise:
        return Noise(
            time=self._time_array.take(indices),
            sigma=self._sigma_array.take(indices),
            sigma_squared=self._sigma_squared_array.take(indices),
            g=self._g_array.take(indices),
            g_squared=self._g_squared_array.take(indices),
            beta=self._beta_array.take(indices),
            alpha_bar=self._alpha_bar_array.take(indices),
            q_matrix=self._q_matrix_array.index_select(dim=0, index=indices),
            q_bar_matrix=self._q_bar_matrix_array.index_select(dim=0, index=indices),
            q_bar_tm1_matrix=self._q_bar_tm1_matrix_array.index_select(dim=0, index=indices),
            indices=indices,
        )

    def get_all_sampling_parameters(self) -> Tuple[Noise, LangevinDynamics]:
        noise = Noise(
            time=self._time_array,
            sigma=self._sigma_array,
            sigma_squared=self._sigma_squared_array,
            g=self._g_array,
            g_squared=self._g_squared_array,
            beta=self._beta_array,
            alpha_bar=self._alpha_bar_array,
            q_matrix=self._q_matrix_array,
            q_bar_matrix=self._q_bar_matrix_array,
            q_bar_tm1_matrix=self._q_bar_tm1_matrix_array,
            indices=torch.arange(self._minimum_random_index, self._maximum_random_index + 1),
        )
        langevin_dynamics = LangevinDynamics(epsilon=self._epsilon_array, sqrt_2_epsilon=self._sqrt_two_epsilon_array)
        return noise, langevin_dynamics

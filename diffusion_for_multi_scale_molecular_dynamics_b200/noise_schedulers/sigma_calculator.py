"""sigma(t) for the exploding-variance schedules (reference: noise_schedulers/sigma_calculator.py:5-108).

Host-side, evaluated once when the tables are built.  Arithmetic is kept in float32 tensors in the same
operation order as the reference so that the tables are bit-identical (pinned by tests/golden).
"""
import torch


class SigmaCalculator(torch.nn.Module):
    """Base class: holds sigma_min / sigma_max as frozen float32 scalars."""

    def __init__(self, sigma_min: float, sigma_max: float):
        super().__init__()
        self.sigma_min = torch.nn.Parameter(torch.tensor(sigma_min), requires_grad=False)
        self.sigma_max = torch.nn.Parameter(torch.tensor(sigma_max), requires_grad=False)

    def get_sigma(self, times: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("This method must be implemented in a child class.")

    def get_sigma_time_derivative(self, times: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("This method must be implemented in a child class.")

    def forward(self, times: torch.Tensor) -> torch.Tensor:
        return self.get_sigma(times)


class ExponentialSigmaCalculator(SigmaCalculator):
    """sigma(t) = sigma_min * (sigma_max / sigma_min) ** t."""

    def __init__(self, sigma_min: float, sigma_max: float):
        super().__init__(sigma_min, sigma_max)
        quotient = self.sigma_max / self.sigma_min
        self.ratio = torch.nn.Parameter(quotient, requires_grad=False)
        self.log_ratio = torch.nn.Parameter(torch.log(quotient), requires_grad=False)

    def get_sigma(self, times: torch.Tensor) -> torch.Tensor:
        return self.sigma_min * self.ratio**times

    def get_sigma_time_derivative(self, times: torch.Tensor) -> torch.Tensor:
        return self.log_ratio * self.get_sigma(times)


class LinearSigmaCalculator(SigmaCalculator):
    """sigma(t) = sigma_min + (sigma_max - sigma_min) * t."""

    def __init__(self, sigma_min: float, sigma_max: float):
        super().__init__(sigma_min, sigma_max)
        self.sigma_difference = torch.nn.Parameter(self.sigma_max - self.sigma_min, requires_grad=False)

    def get_sigma(self, times: torch.Tensor) -> torch.Tensor:
        return self.sigma_min + self.sigma_difference * times

    def get_sigma_time_derivative(self, times: torch.Tensor) -> torch.Tensor:
        return self.sigma_difference * torch.ones_like(times)


_CALCULATORS = {"exponential": ExponentialSigmaCalculator, "linear": LinearSigmaCalculator}


def instantiate_sigma_calculator(sigma_min: float, sigma_max: float, schedule_type: str) -> SigmaCalculator:
    try:
        return _CALCULATORS[schedule_type](sigma_min, sigma_max)
    except KeyError:
        raise NotImplementedError(f"The schedule type {schedule_type} is not implemented")

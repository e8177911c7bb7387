"""Noise-schedule hyper-parameters (reference: noise_schedulers/noise_parameters.py:4-36).

Field names and defaults are part of the drop-in contract (yaml configs are splatted into this dataclass).
"""
from dataclasses import dataclass

SUPPORTED_SCHEDULES = ("exponential", "linear")


@dataclass
class NoiseParameters:
    """Exploding-variance schedule on t in [time_delta, 1] discretised in total_time_steps points."""

    total_time_steps: int
    schedule_type: str = "exponential"  # "exponential": sigma_min (sigma_max/sigma_min)^t ; "linear"
    time_delta: float = 1e-5  # t = 0 is singular for exploding variance, start slightly above
    sigma_min: float = 0.005  # relative-coordinate units: the torsional-diffusion values divided by 2 pi
    sigma_max: float = 0.5
    corrector_step_epsilon: float = 2e-5  # Langevin corrector step scale (Song & Ermon 2019)
    corrector_r: float = 0.17  # only used by the adaptive corrector variant (out of scope here)

    def __post_init__(self):
        assert self.schedule_type in SUPPORTED_SCHEDULES, f"The schedule type {self.schedule_type} is not supported."

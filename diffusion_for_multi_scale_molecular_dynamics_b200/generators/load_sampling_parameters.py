"""Sampling-parameter loader (reference: generators/load_sampling_parameters.py:13-50)."""
from typing import Any, AnyStr, Dict

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.axl_generator import SamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters


def load_sampling_parameters(sampling_parameter_dictionary: Dict[AnyStr, Any]) -> SamplingParameters:
    """Dataclass for the ``algorithm`` named in a yaml ``sampling`` block."""
    assert "algorithm" in sampling_parameter_dictionary, "The sampling parameters must select an algorithm."
    algorithm = sampling_parameter_dictionary["algorithm"]
    assert algorithm in ["ode", "sde", "adaptive_corrector", "predictor_corrector"], \
        "Unknown algorithm. Possible choices are 'ode', 'sde', 'predictor_corrector' and 'adaptive_corrector'"
    if algorithm in ("predictor_corrector", "adaptive_corrector"):
        return PredictorCorrectorSamplingParameters(**sampling_parameter_dictionary)
    raise NotImplementedError(f"algorithm '{algorithm}' is not part of the sm_100a sampling path "
                              "(predictor_corrector and adaptive_corrector only)")

"""Registry of score networks on the sm_100a sampling path (reference: models/score_networks/score_network_factory.py).

The EGNN and the (non-symmetrised) analytical score network are implemented natively; the reference's other
architectures (mlp, mace, diffusion_mace, ...) are outside the accelerated path and raise ``NotImplementedError``.
"""
import dataclasses
from typing import Any, AnyStr, Dict

from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.analytical_score_network import (
    AnalyticalScoreNetwork, AnalyticalScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import (
    ScoreNetwork, ScoreNetworkParameters)

SCORE_NETWORKS_BY_ARCH = dict(egnn=EGNNScoreNetwork, analytical=AnalyticalScoreNetwork)
SCORE_NETWORK_PARAMETERS_BY_ARCH = dict(egnn=EGNNScoreNetworkParameters, analytical=AnalyticalScoreNetworkParameters)


def create_score_network(score_network_parameters: ScoreNetworkParameters) -> ScoreNetwork:
    """Instantiate the network named by ``score_network_parameters.architecture`` (reference :33-62)."""
    architecture = score_network_parameters.architecture
    if architecture not in SCORE_NETWORKS_BY_ARCH:
        raise NotImplementedError(f"Architecture {architecture} is not available on the sm_100a sampling path. "
                                  f"Choices are {list(SCORE_NETWORKS_BY_ARCH)}")
    expected = SCORE_NETWORK_PARAMETERS_BY_ARCH[architecture]
    assert isinstance(score_network_parameters, expected), \
        f"The parameters for architecture {architecture} should be of type {expected.__name__}."
    return SCORE_NETWORKS_BY_ARCH[architecture](score_network_parameters)


def create_score_network_parameters(score_network_dictionary: Dict[AnyStr, Any],
                                    global_parameters_dictionary: Dict[AnyStr, Any]) -> ScoreNetworkParameters:
    """Dataclass from a yaml ``score_network`` block, pulling shared fields from the global block (reference :65-125)."""
    assert "architecture" in score_network_dictionary, "The architecture of the score network must be specified."
    architecture = score_network_dictionary["architecture"]
    if architecture not in SCORE_NETWORK_PARAMETERS_BY_ARCH:
        raise NotImplementedError(f"Architecture {architecture} is not available on the sm_100a sampling path.")
    data_class = SCORE_NETWORK_PARAMETERS_BY_ARCH[architecture]
    field_names = {f.name for f in dataclasses.fields(data_class)}
    merged = dict(score_network_dictionary)
    for key, value in global_parameters_dictionary.items():
        if key in field_names:
            if key in merged:
                assert merged[key] == value, f"inconsistent configuration values for key '{key}'"
            merged[key] = value
    return data_class(**merged)

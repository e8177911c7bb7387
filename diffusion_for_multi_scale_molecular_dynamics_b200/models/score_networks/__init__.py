from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import (  # noqa: F401
    ScoreNetwork, ScoreNetworkParameters)

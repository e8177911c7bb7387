"""AdaptiveCorrectorGenerator against the golden vectors of the reference's
generators/adaptive_corrector.py: the step-size reduction, and a corrector-only trajectory with injected noise."""
import dataclasses

import pytest
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.adaptive_corrector import AdaptiveCorrectorGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.instantiate_generator import instantiate_generator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
from oracle import pc_oracle as oracle
from tests.conftest import load_golden, sub_dict
from tests.egnn_variants import EGNN_VARIANTS


def build(g, device=None):
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]))
    net.load_state_dict(sub_dict(g, "sd"), strict=True)
    if device is not None:
        net = net.to(device)
    B, N, _ = g["x0"].shape
    a = float(g["l"][0, 0])
    sampling = PredictorCorrectorSamplingParameters(
        algorithm="adaptive_corrector", number_of_samples=B, number_of_atoms=N, num_atom_types=2,
        number_of_corrector_steps=int(g["n_corr"]), use_fixed_lattice_parameters=True, cell_dimensions=[a, a, a])
    noise = NoiseParameters(total_time_steps=int(g["T"]), sigma_min=0.005, sigma_max=0.5, corrector_r=0.17)
    return instantiate_generator(sampling, noise, net), int(g["T"])


def test_factory_and_host_plumbing():
    g = sub_dict(load_golden("adaptive_corrector"), "traj")
    gen, _ = build(g)
    assert type(gen) is AdaptiveCorrectorGenerator and gen.corrector_r == 0.17
    assert gen._hooks_overridden()  # the corrector-only scheme runs through the step-level hooks, not the fused loop
    x = torch.rand(2, 16, 3)
    assert gen._relative_coordinates_update_predictor_step(x, x, 0.1, 0.2, 0.3, None) is x
    lattice = torch.rand(2, 6)
    assert gen._lattice_parameters_update_predictor_step(lattice, lattice, 0.1, 0.2, 0.3, None) is lattice
    sampling = dataclasses.replace(gen.sampling_parameters, algorithm="sde") if hasattr(gen, "sampling_parameters") else None
    if sampling is not None:
        with pytest.raises(NotImplementedError):
            instantiate_generator(sampling, gen.noise_parameters, gen.axl_network)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n_dim", [("x", 3), ("l", 2), ("tiny", 3)])
def test_adaptive_step_size_golden(cuda_device, tag, n_dim):
    """eps = 2 (r |z| / max(|sigma s| / sigma, small_eps))^2: same value as the reference to fp32 reduction order."""
    g = load_golden("adaptive_corrector")
    gen, _ = build(sub_dict(g, "traj"), cuda_device)
    scores, z, sigma = (g[f"step.{tag}.{k}"] for k in ("scores", "z", "sigma"))
    eps = gen._generic_corrector_step_size(2, float(sigma), scores.to(cuda_device), z.to(cuda_device), n_dim=n_dim)
    expected = g[f"step.{tag}.eps"]
    assert eps.is_cuda and eps.shape == expected.shape
    torch.testing.assert_close(eps.cpu(), expected, rtol=2e-6, atol=0.0)


@pytest.mark.gpu
def test_adaptive_corrector_trajectory_golden(cuda_device):
    """Corrector-only trajectory of the reference reproduced with the same injected noise."""
    from tests.test_gpu_parity import TapedHooks, torus_distance
    g = sub_dict(load_golden("adaptive_corrector"), "traj")
    gen, T = build(g, cuda_device)
    start = AXL(A=g["a0"].clone().to(cuda_device), X=g["x0"].clone().to(cuda_device), L=g["l"].to(cuda_device))
    with TapedHooks(gen, oracle.NoiseTape(79)):
        final = gen.sample_from_noisy_composition(start, T, 0)
    assert torch.equal(final.A.cpu(), g["final_a"])
    assert torus_distance(final.X, g["final_x"]).max() < 2e-5
    assert not torch.equal(final.X.cpu(), g["x0"])

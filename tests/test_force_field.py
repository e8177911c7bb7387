"""ForceFieldAugmentedScoreNetwork against the golden vectors of the reference's
models/score_networks/force_field_augmented_score_network.py."""
import pytest
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.force_field_augmented_score_network import (
    ForceFieldAugmentedScoreNetwork, ForceFieldParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, CARTESIAN_FORCES, NOISE,
                                                                       NOISY_AXL_COMPOSITION, TIME)
from oracle import pc_oracle as oracle
from tests.conftest import load_golden, sub_dict
from tests.egnn_variants import EGNN_VARIANTS

CASES = [(1.5, 1.0), (2.5, 0.3)]


def batch_on(g, device):
    B, N, _ = g["x"].shape
    return {NOISY_AXL_COMPOSITION: AXL(A=g["a"].to(device), X=g["x"].to(device), L=g["l"].to(device)),
            TIME: torch.full((B, 1), 0.5, device=device), NOISE: g["sigma"].to(device),
            CARTESIAN_FORCES: torch.zeros(B, N, 3, device=device)}


@pytest.mark.parametrize("rc,strength", CASES)
def test_oracle_pseudo_force_against_reference(rc, strength):
    g = load_golden("force_field")
    out = oracle.repulsive_pseudo_force(g["x"], g["l"], rc, strength)
    expected = g[f"rc{rc}.force"]
    scale = expected.abs().max()
    assert (out - expected).abs().max() <= 2e-6 * scale


def test_force_field_parameters_validation():
    with pytest.raises(AssertionError):
        ForceFieldParameters(radial_cutoff=0.0, strength=1.0)
    with pytest.raises(AssertionError):
        ForceFieldParameters(radial_cutoff=1.0, strength=-1.0)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]))
    ff = ForceFieldAugmentedScoreNetwork(net, ForceFieldParameters(radial_cutoff=1.5, strength=1.0))
    ff.defer_status_check = True
    assert net.defer_status_check
    ff.defer_status_check = False
    g = load_golden("force_field")
    with pytest.raises(_cabi.B200MDError):  # no CPU path
        ff.get_relative_coordinates_pseudo_force(batch_on(g, torch.device("cpu")))


@pytest.mark.gpu
@pytest.mark.parametrize("rc,strength", CASES)
def test_gpu_pseudo_force_and_augmented_scores(cuda_device, rc, strength):
    """Pseudo force within 1e-5 of the largest reference component (a different summation order over the 27 images);
    augmented score = EGNN score + force; logits untouched."""
    g = load_golden("force_field")
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]))
    net.load_state_dict(sub_dict(g, "sd"), strict=True)
    net = net.to(cuda_device)
    ff = ForceFieldAugmentedScoreNetwork(net, ForceFieldParameters(radial_cutoff=rc, strength=strength))
    batch = batch_on(g, cuda_device)
    force = ff.get_relative_coordinates_pseudo_force(batch).cpu()
    expected = g[f"rc{rc}.force"]
    assert (force - expected).abs().max() <= 1e-5 * expected.abs().max()
    raw = net(batch, conditional=False)
    augmented = ff(batch, conditional=False)
    expected_x = g[f"rc{rc}.augmented_x"]
    assert (augmented.X.cpu() - expected_x).abs().max() <= 1e-5 * expected_x.abs().max()
    torch.testing.assert_close(augmented.X - raw.X, force.to(cuda_device), rtol=0, atol=1e-5 * float(expected_x.abs().max()))
    assert torch.equal(augmented.A, raw.A) and torch.equal(augmented.L, raw.L)


@pytest.mark.gpu
def test_gpu_pseudo_force_cutoff_guard_and_sampling(cuda_device):
    """The adjacency assert of the reference (cutoff beyond the first image shell), and a short constrained-free
    sampling run through the generator with the augmented network."""
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
        PredictorCorrectorSamplingParameters
    from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
    g = load_golden("force_field")
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]))
    net.load_state_dict(sub_dict(g, "sd"), strict=True)
    net = net.to(cuda_device)
    too_large = ForceFieldAugmentedScoreNetwork(net, ForceFieldParameters(radial_cutoff=8.0, strength=1.0))
    with pytest.raises(AssertionError):
        too_large.get_relative_coordinates_pseudo_force(batch_on(g, cuda_device))
    ff = ForceFieldAugmentedScoreNetwork(net, ForceFieldParameters(radial_cutoff=1.5, strength=1.0))
    sampling = PredictorCorrectorSamplingParameters(number_of_samples=6, number_of_atoms=16, num_atom_types=2,
                                                    use_fixed_lattice_parameters=True, cell_dimensions=[7.0] * 3)
    generator = LangevinGenerator(NoiseParameters(total_time_steps=8), sampling, ff)
    sample = generator.sample(6, cuda_device)
    assert sample.X.shape == (6, 16, 3) and bool(((sample.X >= 0) & (sample.X < 1)).all())
    assert bool((sample.A != generator.masked_atom_type_index).all())

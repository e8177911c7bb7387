"""Structure statistics next to the sampling path: compute_distances_in_batch and the
Kolmogorov-Smirnov metric.  CPU: the oracle against the reference's golden vectors; GPU: the CUDA path against both."""
import numpy as np
import pytest
import scipy.stats as ss
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.metrics.kolmogorov_smirnov_metrics import KolmogorovSmirnovMetrics
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.structure_utils import (compute_distances_in_batch,
                                                                                   get_orthogonal_basis_vectors)
from oracle import pc_oracle as oracle
from tests.conftest import load_golden

CASES = [("tri", 1.7), ("tri", 2.6), ("si", 5.0)]


@pytest.mark.parametrize("tag,rc", CASES)
def test_oracle_pair_distances_against_reference(tag, rc):
    """Bit-exact: same rounding chain for d^2, same (CPU) torch.sqrt, same order."""
    g = load_golden("pair_distances")
    out = oracle.compute_distances_in_batch(g[f"{tag}.cart"], g[f"{tag}.basis"], rc)
    expected = g[f"{tag}.distances_rc{rc}"]
    assert torch.equal(out, expected)
    ieee = oracle.compute_distances_in_batch(g[f"{tag}.cart"], g[f"{tag}.basis"], rc, ieee_sqrt=True)
    torch.testing.assert_close(ieee, expected, rtol=1.2e-7, atol=0.0)  # torch's CPU sqrt is one ulp off now and then
    assert bool((out > 0).all()) and bool((out <= rc).all())


def test_kolmogorov_smirnov_metric_matches_scipy():
    rng = np.random.default_rng(3)
    reference = [torch.from_numpy(rng.normal(size=400)), torch.from_numpy(rng.normal(size=300))]
    predicted = [torch.from_numpy(rng.normal(loc=0.3, size=500))]
    metric = KolmogorovSmirnovMetrics()
    for chunk in reference:
        metric.register_reference_samples(chunk)
    for chunk in predicted:
        metric.register_predicted_samples(chunk)
    distance, p_value = metric.compute_kolmogorov_smirnov_distance_and_pvalue()
    expected = ss.ks_2samp(torch.cat(predicted).numpy(), torch.cat(reference).numpy())
    assert distance == expected.statistic and p_value == expected.pvalue
    assert metric.reference_count == 700 and metric.predicted_count == 500
    metric.reset()
    assert metric.reference_count == 0 and metric.predicted_count == 0
    capped = KolmogorovSmirnovMetrics(maximum_number_of_samples=10)
    capped.register_reference_samples(torch.zeros(8))
    capped.register_reference_samples(torch.zeros(8))
    capped.register_reference_samples(torch.zeros(8))  # ignored: 16 >= 10 already registered
    assert capped.reference_count == 16


def test_orthogonal_basis_vectors():
    basis = get_orthogonal_basis_vectors(3, [5.0, 6.0, 7.0])
    assert basis.shape == (3, 3, 3) and torch.equal(basis[2], torch.diag(torch.tensor([5.0, 6.0, 7.0])))


def test_pair_distances_have_no_cpu_fallback():
    with pytest.raises(_cabi.B200MDError):
        compute_distances_in_batch(torch.rand(1, 4, 3), torch.eye(3).unsqueeze(0), 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,rc", CASES)
def test_gpu_pair_distances_against_reference(cuda_device, tag, rc):
    """Same count and order as the reference run on CPU, values to one ulp (torch's CPU sqrt is not correctly rounded);
    bit-exact against the oracle with the correctly rounded square root CUDA uses."""
    g = load_golden("pair_distances")
    out = compute_distances_in_batch(g[f"{tag}.cart"].to(cuda_device), g[f"{tag}.basis"].to(cuda_device), rc)
    expected = g[f"{tag}.distances_rc{rc}"]
    assert out.shape == expected.shape
    torch.testing.assert_close(out.cpu(), expected, rtol=2e-7, atol=0.0)
    assert torch.equal(out.cpu(), oracle.compute_distances_in_batch(g[f"{tag}.cart"], g[f"{tag}.basis"], rc, ieee_sqrt=True))


@pytest.mark.gpu
def test_gpu_pair_distances_at_sampling_size(cuda_device):
    """512 structures of 64 atoms against the oracle, and the structure KS distance of two ensembles."""
    torch.manual_seed(9)
    B, N, length = 512, 64, 10.86
    x = torch.rand(B, N, 3)
    cell = get_orthogonal_basis_vectors(B, [length] * 3)
    cart = x * length
    out = compute_distances_in_batch(cart.to(cuda_device), cell.to(cuda_device), 4.0).cpu()
    expected = oracle.compute_distances_in_batch(cart[:16], cell[:16], 4.0, ieee_sqrt=True)
    assert torch.equal(out[:expected.numel()], expected)
    assert bool((out > 0).all()) and bool((out <= 4.0).all())
    metric = KolmogorovSmirnovMetrics()
    metric.register_reference_samples(out[: out.numel() // 2])
    metric.register_predicted_samples(out[out.numel() // 2:])
    distance, _ = metric.compute_kolmogorov_smirnov_distance_and_pvalue()
    assert distance < 0.01  # two halves of the same uniform ensemble


@pytest.mark.gpu
def test_sampled_structures_match_the_data_distribution(cuda_device):
    """The reference's validation loop in miniature (models/axl_diffusion_lightning_model.py:520-640): sample with the
    analytical score network (a diamond crystal with Gaussian displacements sigma_d), then compare the samples with
    structures drawn directly from that data distribution through the reference's two sample-quality statistics --
    pair distances and Stillinger-Weber energies, each through KolmogorovSmirnovMetrics -- all computed on the GPU."""
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
        PredictorCorrectorSamplingParameters
    from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.analytical_score_network import (
        AnalyticalScoreNetwork, AnalyticalScoreNetworkParameters)
    from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL, AXL_COMPOSITION
    from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
    from diffusion_for_multi_scale_molecular_dynamics_b200.oracle.energy_oracle_factory import (
        create_energy_oracle, create_energy_oracle_parameters)
    diamond = [[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.25, .25, .25], [.25, .75, .75], [.75, .25, .75],
               [.75, .75, .25]]
    sigma_d, B, T, length = 0.01, 1024, 200, 5.43
    network = AnalyticalScoreNetwork(AnalyticalScoreNetworkParameters(
        number_of_atoms=8, num_atom_types=1, kmax=4, equilibrium_relative_coordinates=diamond, sigma_d=sigma_d)).to(cuda_device)
    sampling = PredictorCorrectorSamplingParameters(number_of_samples=B, number_of_atoms=8, num_atom_types=1,
                                                    use_fixed_lattice_parameters=True, cell_dimensions=[length] * 3)
    generator = LangevinGenerator(NoiseParameters(total_time_steps=T, sigma_min=0.001), sampling, network)
    generator.seed = 23
    samples = generator.sample(B, cuda_device)
    torch.manual_seed(5)
    data_x = torch.remainder(torch.tensor(diamond) + sigma_d * torch.randn(B, 8, 3), 1.0).to(cuda_device)
    cell = get_orthogonal_basis_vectors(B, [length] * 3).to(cuda_device)
    structure_metric = KolmogorovSmirnovMetrics()
    structure_metric.register_reference_samples(compute_distances_in_batch(data_x * length, cell, 5.0))
    structure_metric.register_predicted_samples(compute_distances_in_batch(samples.X * length, cell, 5.0))
    distance, _ = structure_metric.compute_kolmogorov_smirnov_distance_and_pvalue()
    assert distance < 0.02, f"structure KS distance {distance:.4f}"
    energy_oracle = create_energy_oracle(create_energy_oracle_parameters(dict(name="lammps", sw_coeff_filename="Si.sw"), ["Si"]))
    energy_metric = KolmogorovSmirnovMetrics()
    energy_metric.register_reference_samples(energy_oracle.compute_oracle_energies(
        {AXL_COMPOSITION: AXL(A=samples.A, X=data_x, L=samples.L)}))
    energy_metric.register_predicted_samples(energy_oracle.compute_oracle_energies({AXL_COMPOSITION: samples}))
    distance, _ = energy_metric.compute_kolmogorov_smirnov_distance_and_pvalue()
    # the energy grows with the square of the displacement width, which the 200-step sampler reproduces to a few per
    # cent (finite step, sigma_min): a visible but small shift of the energy distribution
    assert distance < 0.2, f"energy KS distance {distance:.4f}"
    # a deliberately wrong ensemble (twice the displacement width) is told apart by both statistics
    wrong_x = torch.remainder(torch.tensor(diamond) + 2 * sigma_d * torch.randn(B, 8, 3), 1.0).to(cuda_device)
    energy_metric.reset()
    energy_metric.register_reference_samples(energy_oracle.compute_oracle_energies(
        {AXL_COMPOSITION: AXL(A=samples.A, X=data_x, L=samples.L)}))
    energy_metric.register_predicted_samples(energy_oracle.compute_oracle_energies(
        {AXL_COMPOSITION: AXL(A=samples.A, X=wrong_x, L=samples.L)}))
    assert energy_metric.compute_kolmogorov_smirnov_distance_and_pvalue()[0] > 0.5

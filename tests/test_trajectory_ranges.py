"""Trajectory loop variants against the CPU oracle with the same injected noise: no corrector, several correctors,
partial index ranges (sample_from_noisy_composition(start, starting_step_index, ending_step_index)), through both the
step-by-step Python loop and the fused single-call loop (b200md_pc_sample, in-kernel noise; compared with the
step-by-step loop run with the same counter-based noise)."""
import pytest
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
from oracle import pc_oracle as oracle
from tests.egnn_variants import EGNN_VARIANTS, hyper_parameters
from tests.test_gpu_parity import TapedHooks, torus_distance

pytestmark = pytest.mark.gpu

B, N, T, LENGTH = 4, 16, 6, 7.0


def setup(cuda_device, n_corr):
    torch.manual_seed(50 + n_corr)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sampling = PredictorCorrectorSamplingParameters(number_of_samples=B, number_of_atoms=N, num_atom_types=2,
                                                    number_of_corrector_steps=n_corr, use_fixed_lattice_parameters=True,
                                                    cell_dimensions=[LENGTH] * 3)
    gen = LangevinGenerator(NoiseParameters(total_time_steps=T, sigma_min=0.005, sigma_max=0.5), sampling, net.to(cuda_device))
    x0 = torch.rand(B, N, 3)
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = LENGTH
    return gen, sd, x0, lattice


@pytest.mark.parametrize("n_corr,start,end", [(0, T, 0), (3, T, 0), (1, 4, 2), (2, 3, 0), (1, T, T - 1)])
def test_loop_variants_against_oracle(cuda_device, n_corr, start, end):
    gen, sd, x0, lattice = setup(cuda_device, n_corr)
    hp = hyper_parameters("radial_sige")
    # a partially unmasked start for the ranges that do not begin at T
    a0 = torch.full((B, N), 2) if start == T else torch.randint(0, 3, (B, N))
    tables = oracle.noise_tables(T, 3, sigma_min=0.005, sigma_max=0.5)

    def score_fn(at, x, lat, sigma):
        return oracle.egnn_score_forward(sd, hp, at, x, lat, torch.full((B, 1), sigma))

    try:  # the reference asserts that no MASK is left after the last step (langevin_generator.py:616-620)
        ref_a, ref_x, _ = oracle.pc_sample(score_fn, tables, a0.clone(), x0.clone(), lattice, oracle.NoiseTape(90),
                                           number_of_corrector_steps=n_corr, starting_step_index=start,
                                           ending_step_index=end)
    except AssertionError:
        pytest.skip("the oracle's random start leaves a MASK at the last step for this seed")
    composition = AXL(A=a0.clone().to(cuda_device), X=x0.clone().to(cuda_device), L=lattice.to(cuda_device))
    with TapedHooks(gen, oracle.NoiseTape(90)):
        out = gen.sample_from_noisy_composition(composition, start, end)
    assert torch.equal(out.A.cpu(), ref_a)
    assert torus_distance(out.X, ref_x).max() < 2e-5


@pytest.mark.parametrize("n_corr,start,end", [(0, T, 0), (3, T, 0), (2, 5, 2)])
def test_fused_loop_equals_step_by_step(cuda_device, n_corr, start, end):
    gen, _, x0, lattice = setup(cuda_device, n_corr)
    gen.seed = 77
    a0 = torch.full((B, N), 2)
    make = lambda: AXL(A=a0.clone().to(cuda_device), X=x0.clone().to(cuda_device), L=lattice.to(cuda_device))  # noqa: E731
    fused = gen.sample_from_noisy_composition(make(), start, end)
    gen.allow_fused = False
    stepped = gen.sample_from_noisy_composition(make(), start, end)
    assert torch.equal(fused.A, stepped.A)
    assert torus_distance(fused.X, stepped.X).max() < 2e-5


@pytest.mark.parametrize("n_corr,start,end", [(1, T, 0), (2, 5, 1)])
def test_graph_replay_equals_direct_launches(cuda_device, n_corr, start, end, monkeypatch):
    """b200md_pc_sample replays ONE captured step graph whose kernels read their step scalars from a device table
    (csrc/sampler.cu); B200MD_PC_DIRECT=1 enqueues the same kernels one by one with the scalars as launch arguments.
    Same kernels, same arguments: same atom types, coordinates equal up to the order of the fp32 atomics."""
    gen, _, x0, lattice = setup(cuda_device, n_corr)
    gen.seed = 31
    a0 = torch.full((B, N), 2)
    make = lambda: AXL(A=a0.clone().to(cuda_device), X=x0.clone().to(cuda_device), L=lattice.to(cuda_device))  # noqa: E731
    from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
    gen.axl_network.packed_weights(cuda_device)   # the packing kernels are launched (and counted) on first use
    before = _cabi.launch_count()
    replayed = gen.sample_from_noisy_composition(make(), start, end)
    launches_graph = _cabi.launch_count() - before
    assert _cabi.lib().b200md_pc_sample_last_mode() == 1, "the step graph was not captured / instantiated"
    monkeypatch.setenv("B200MD_PC_DIRECT", "1")
    before = _cabi.launch_count()
    direct = gen.sample_from_noisy_composition(make(), start, end)
    launches_direct = _cabi.launch_count() - before
    assert _cabi.lib().b200md_pc_sample_last_mode() == 0
    assert torch.equal(replayed.A, direct.A)
    assert torus_distance(replayed.X, direct.X).max() < 2e-5
    assert not torch.equal(replayed.X.cpu(), x0)
    # the launch counter counts kernels executed: the graph adds its table loader and one cursor kernel per step
    assert launches_graph == launches_direct + 1 + (start - end)

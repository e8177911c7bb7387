"""Lattice diffusion (use_fixed_lattice_parameters=False): predictor / corrector steps and a short trajectory against
the golden vectors of the reference's LangevinGenerator (langevin_generator.py:441-490,536-805).  The EGNN lattice
score is zero, so the lattice parameters move by the injected noise and the radius graph follows the changing box."""
from unittest import mock

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
from tests.conftest import load_golden, sub_dict
from tests.egnn_variants import EGNN_VARIANTS

pytestmark = pytest.mark.gpu


class LatticeTapedHooks:
    """The golden generator's hooks: coordinates, Gumbel, binary and LATTICE noise from one NoiseTape."""

    def __init__(self, generator, tape):
        g = generator
        self.patches = [
            mock.patch.object(g, "_draw_coordinates_gaussian_sample",
                              side_effect=lambda n: tape.normal("coordinates", n, g.number_of_atoms, 3)),
            mock.patch.object(g, "_draw_lattice_gaussian_sample",
                              side_effect=lambda n: tape.normal("lattice", n, g.num_lattice_parameters)),
            mock.patch.object(g, "_draw_gumbel_sample",
                              side_effect=lambda n: tape.gumbel(g.small_epsilon, n, g.number_of_atoms, g.num_classes)),
            mock.patch.object(g, "_draw_binary_sample", side_effect=lambda n: tape.uniform("binary", n, g.number_of_atoms)),
        ]

    def __enter__(self):
        for p in self.patches:
            p.start()
        return self

    def __exit__(self, *exc):
        for p in self.patches:
            p.stop()


def build(g, device):
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]))
    net.load_state_dict(sub_dict(g, "sd"), strict=True)
    B, N, _ = g["x0"].shape
    sampling = PredictorCorrectorSamplingParameters(
        number_of_samples=B, number_of_atoms=N, num_atom_types=2, number_of_corrector_steps=int(g["n_corr"]),
        use_fixed_lattice_parameters=False, cell_dimensions=[7.0, 7.0, 7.0])
    noise = NoiseParameters(total_time_steps=int(g["T"]), sigma_min=0.005, sigma_max=0.5)
    return LangevinGenerator(noise, sampling, net.to(device)), int(g["T"])


def torus(a, b):
    d = (a.detach().cpu() - b).abs()
    return torch.minimum(d, 1.0 - d)


def test_free_lattice_single_steps(cuda_device):
    g = load_golden("free_lattice")
    gen, _ = build(g, cuda_device)
    start = AXL(A=g["a0"].to(cuda_device), X=g["x0"].to(cuda_device), L=g["l"].to(cuda_device))
    forces = torch.zeros_like(start.X)
    with LatticeTapedHooks(gen, oracle.NoiseTape(81)):
        p = gen.predictor_step(start, 3, forces)
        c = gen.corrector_step(p, 2, forces)
    assert torch.equal(p.A.cpu(), g["pred_a"]) and torch.equal(c.A.cpu(), g["corr_a"])
    torch.testing.assert_close(p.L.cpu(), g["pred_l"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(c.L.cpu(), g["corr_l"], rtol=1e-6, atol=1e-6)
    assert not torch.equal(p.L.cpu(), g["l"])  # the lattice did move
    assert torus(p.X, g["pred_x"]).max() < 1e-5 and torus(c.X, g["corr_x"]).max() < 1e-5


def test_free_lattice_trajectory(cuda_device):
    g = load_golden("free_lattice")
    gen, T = build(g, cuda_device)
    start = AXL(A=g["a0"].clone().to(cuda_device), X=g["x0"].clone().to(cuda_device), L=g["l"].clone().to(cuda_device))
    with LatticeTapedHooks(gen, oracle.NoiseTape(80)):
        final = gen.sample_from_noisy_composition(start, T, 0)
    assert torch.equal(final.A.cpu(), g["final_a"])
    torch.testing.assert_close(final.L.cpu(), g["final_l"], rtol=1e-5, atol=1e-5)
    assert torus(final.X, g["final_x"]).max() < 2e-5


def test_free_lattice_native_noise(cuda_device):
    """In-kernel noise: the lattice noise stream is keyed like the others (independent of batch sharding)."""
    g = load_golden("free_lattice")
    gen, T = build(g, cuda_device)
    gen.seed = 7
    full = gen.sample(8, cuda_device)
    gen.sample_offset = 3
    part = gen.sample(2, cuda_device)
    assert torch.equal(part.A, full.A[3:5])
    torch.testing.assert_close(part.L, full.L[3:5], rtol=1e-5, atol=1e-5)
    assert not torch.allclose(full.L[0], full.L[1])


def test_every_draw_of_a_trajectory_has_its_own_counter_key(cuda_device):
    """The counter-based generator returns the same numbers for the same (seed, sample, step id, draw id): no two draws of
    a trajectory may share (step id, draw id).  In particular the starting lattice L_T must not reuse the key of the last
    predictor step's lattice noise (step id 0, DRAW_LATTICE_NORMAL)."""
    from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
    g = load_golden("free_lattice")
    gen, T = build(g, cuda_device)
    gen.seed = 11
    keys = [(0, _cabi.DRAW_INIT_UNIFORM), (0, _cabi.DRAW_INIT_LATTICE)]   # FullRandomTrajectoryInitializer.initialize
    original = gen._draw

    def recording_draw(kind, number_of_samples, units, unit_size, draw_id):
        keys.append((gen._step_id, draw_id))
        return original(kind, number_of_samples, units, unit_size, draw_id)

    with mock.patch.object(gen, "_draw", side_effect=recording_draw):
        final = gen.sample(4, cuda_device)
    assert len(keys) > 4 * T and len(set(keys)) == len(keys), sorted(k for k in keys if keys.count(k) > 1)
    # and the starting lattice differs from the noise the last predictor step adds
    start = gen.trajectory_initializer.initialize(4, cuda_device)
    gen._device, gen._step_id = cuda_device, 0
    last_noise = gen._draw_lattice_gaussian_sample(4)
    assert not torch.equal(start.L, last_noise) and bool(torch.isfinite(final.L).all())


def test_standalone_steps_draw_fresh_noise(cuda_device):
    """corrector_step / predictor_step called outside sample() (a user-written Langevin loop): consecutive calls at the
    same time index must not reuse their noise -- in the reference every call advances the global generator."""
    g = load_golden("free_lattice")
    gen, _ = build(g, cuda_device)
    start = AXL(A=g["a0"].to(cuda_device), X=g["x0"].to(cuda_device), L=g["l"].to(cuda_device))
    forces = torch.zeros_like(start.X)
    for seed in (None, 5):
        gen.seed = seed
        torch.manual_seed(0)
        first = gen.corrector_step(start, 2, forces)
        second = gen.corrector_step(start, 2, forces)
        assert not torch.equal(first.X, second.X) and not torch.equal(first.L, second.L)
        p1, p2 = gen.predictor_step(start, 3, forces), gen.predictor_step(start, 3, forces)
        assert not torch.equal(p1.X, p2.X)
    gen.seed = None
    torch.manual_seed(0)   # torch.manual_seed still controls reproducibility
    again = gen.corrector_step(start, 2, forces)
    torch.manual_seed(0)
    assert torch.equal(gen.corrector_step(start, 2, forces).X, again.X)

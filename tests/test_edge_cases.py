"""GPU edge cases of the path: empty and degenerate inputs, tile-boundary edge counts, and size-independent properties of
the score network at the largest BASELINE supercell (4x4x4, 512 atoms), where the CPU oracle is too slow to run."""
import pytest
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import \
    map_relative_coordinates_to_unit_cell
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.structure_utils import compute_distances_in_batch
from tests.test_gpu_parity import PRODUCTION, assert_anchored, network_batch, oracle_forward_both

pytestmark = pytest.mark.gpu

WIDE = dict(message_hidden_dimensions_size=64, coordinate_hidden_dimensions_size=64, node_hidden_dimensions_size=64,
            message_n_hidden_dimensions=1, coordinate_n_hidden_dimensions=1, node_n_hidden_dimensions=1, n_layers=2)
FULL = dict(residual=True, attention=False, normalize=False, tanh=False, coords_agg="mean", message_agg="mean")


def forward_both(cuda_device, hp, a, x, lattice, sigma, impl=_cabi.IMPL_AUTO, seed=11):
    torch.manual_seed(seed)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(cuda_device)
    net.kernel_impl = impl
    out = net(network_batch(a, x, lattice, sigma, cuda_device), conditional=False)
    return net, out, oracle_forward_both(sd, dict(FULL, **hp), a, x, lattice, sigma)


@pytest.mark.parametrize("impl", [_cabi.IMPL_FFMA, _cabi.IMPL_TCGEN05_F16X3_TS], ids=["ffma", "f16x3ts"])
@pytest.mark.parametrize("B,N,edges", [(64, 2, "fully_connected"), (3, 1, "fully_connected"), (5, 12, "radial_cutoff"),
                                       (1, 40, "radial_cutoff")],
                         ids=["128-edges-one-full-tile", "single-atom-no-edges", "sparse-some-isolated-atoms", "one-structure"])
def test_egnn_degenerate_graphs(cuda_device, B, N, edges, impl):
    """Exactly one full tile of 128 edges, no edge at all, isolated atoms (cutoff 1.2 A in a 9 A box), a single
    structure: the tile loops, the segmented reductions and the mean aggregation (division by max(deg, 1))."""
    hp = dict(num_atom_types=2, edges=edges, **WIDE)
    if edges == "radial_cutoff":
        hp["radial_cutoff"] = 1.2 if B == 5 else 3.0
    torch.manual_seed(B * 100 + N)
    x = torch.rand(B, N, 3)
    a = torch.randint(0, 3, (B, N))
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = 9.0
    sigma = torch.rand(B, 1) * 0.3 + 0.01
    net, out, (logits32, score32, logits64, score64) = forward_both(cuda_device, hp, a, x, lattice, sigma, impl)
    assert_anchored(out.A, logits32, logits64, "logits")
    if N == 1:
        # no edge: zhat = z and S = z . Gamma . z is exactly 0; what the kernel returns is the rounding residue of one
        # fused multiply-add on products of unit-size numbers (c s - s c), at most one ulp of those products
        assert float(score64.abs().max()) == 0.0 and float(out.X.abs().max()) <= 2.0 ** -23
    else:
        assert_anchored(out.X, score32, score64, "score")
    if (B, N) == (64, 2):
        assert net.last_number_of_edges() == 128
    if N == 1:
        assert net.last_number_of_edges() == 0


def test_empty_batches(cuda_device):
    """n_samples == 0 through the C ABI: nothing launched, empty outputs."""
    empty = torch.empty(0, 8, 3, device=cuda_device)
    assert map_relative_coordinates_to_unit_cell(empty).shape == (0, 8, 3)
    cell = torch.empty(0, 3, 3, device=cuda_device)
    assert compute_distances_in_batch(empty, cell, 2.0).shape == (0,)
    out = torch.empty_like(empty)
    assert _cabi.lib().b200md_pc_coordinates_update(_cabi.ptr(empty), _cabi.ptr(empty), None, 0.1, 0.1, 0.1, 0, 8, 3, 1, 1,
                                                    0, _cabi.ptr(out), _cabi.stream_ptr(cuda_device)) == 0
    assert _cabi.lib().b200md_pc_coordinates_update(None, None, None, 0.1, 0.1, 0.1, 4, 8, 3, 1, 1, 0, None,
                                                    _cabi.stream_ptr(cuda_device)) != 0  # null pointers are refused


def test_score_network_invariances_at_the_largest_supercell(cuda_device):
    """Si 4x4x4 (512 atoms, ~45k edges per structure, production widths): the score is invariant under a rigid
    translation of the structure (modulo the cell) and equivariant under a permutation of the atoms; the logits follow
    the permutation.  Exact symmetries of the reference network (models/score_networks/test_score_network_equivariance)."""
    hp = dict(num_atom_types=1, **PRODUCTION)
    torch.manual_seed(21)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp)).to(cuda_device)
    B, N, length = 4, 512, 21.72
    x = torch.rand(B, N, 3)
    a = torch.randint(0, 2, (B, N))
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = length
    sigma = torch.full((B, 1), 0.1)
    base = net(network_batch(a, x, lattice, sigma, cuda_device), conditional=False)
    assert 40000 < net.last_number_of_edges() / B < 50000
    shifted = torch.remainder(x + torch.tensor([0.31, 0.77, 0.05]), 1.0)
    moved = net(network_batch(a, shifted, lattice, sigma, cuda_device), conditional=False)
    scale = base.X.abs().max()
    # two evaluations of the same fp32 kernel on differently rounded inputs: the translation changes every rounding
    assert (moved.X - base.X).abs().max() <= 2.5e-4 * scale
    finite = torch.isfinite(base.A)
    assert (moved.A[finite] - base.A[finite]).abs().max() <= 2e-4 * base.A[finite].abs().max()
    perm = torch.randperm(N)
    permuted = net(network_batch(a[:, perm], x[:, perm], lattice, sigma, cuda_device), conditional=False)
    assert (permuted.X - base.X[:, perm.to(cuda_device)]).abs().max() <= 6e-5 * scale  # two summation orders of one fp32 kernel, each ~1.5e-5 from the exact value
    assert (permuted.A[finite[:, perm.to(cuda_device)]] - base.A[:, perm.to(cuda_device)][finite[:, perm.to(cuda_device)]]
            ).abs().max() <= 2e-5 * base.A[finite].abs().max()


def test_full_size_batch_is_consistent_with_its_parts(cuda_device):
    """BASELINE config 2 at full size (4096 structures x 64 atoms, production widths, 6.5 M edges: ~340 tiles per CTA,
    clusters of two CTAs sharing the weight stream): two predictor-corrector steps of the whole batch against the same
    steps of a 300-structure slice of it taken alone.  The slice lands on different tiles, CTAs and cluster ranks, and the
    noise is keyed by the global sample index, so the rows must agree (up to the order of the fp32 reductions)."""
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
        PredictorCorrectorSamplingParameters
    from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
    from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
    hp = dict(num_atom_types=1, **PRODUCTION)
    torch.manual_seed(33)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp)).to(cuda_device)
    B, N, T = 4096, 64, 1000
    sampling = PredictorCorrectorSamplingParameters(number_of_samples=B, number_of_atoms=N, num_atom_types=1,
                                                    use_fixed_lattice_parameters=True, cell_dimensions=[10.86] * 3)
    gen = LangevinGenerator(NoiseParameters(total_time_steps=T), sampling, net)
    gen.seed = 5
    x0 = torch.rand(B, N, 3, device=cuda_device)
    a0 = torch.full((B, N), 1, dtype=torch.long, device=cuda_device)
    lattice = torch.zeros(B, 6, device=cuda_device)
    lattice[:, :3] = 10.86
    full = gen.sample_from_noisy_composition(AXL(A=a0.clone(), X=x0.clone(), L=lattice), T, T - 2)
    lo, hi = 1717, 2017
    gen.sample_offset = lo
    part = gen.sample_from_noisy_composition(AXL(A=a0[lo:hi].clone(), X=x0[lo:hi].clone(), L=lattice[lo:hi]), T, T - 2)
    gen.sample_offset = 0
    assert torch.equal(part.A, full.A[lo:hi])
    d = (part.X - full.X[lo:hi]).abs()
    assert torch.minimum(d, 1 - d).max() < 2e-5
    assert bool(((full.X >= 0) & (full.X < 1)).all()) and not torch.equal(full.X, x0)


def test_forward_is_bitwise_reproducible_when_segments_span_two_warps(cuda_device):
    """The segmented sums end in one `red.global.add.f32` per (segment tail, column): a node with at most 32 edges falls
    into at most two warps of the edge kernel, receives at most two partial sums, and 0 + a + b does not depend on their
    arrival order -- two evaluations of the same batch are bitwise identical.  With longer segments (512 atoms: 88 edges
    per atom) three or more partials are added in arrival order, like the reference's atomic `scatter_add_` on CUDA
    (models/egnn_utils.py:37)."""
    from oracle import pc_oracle
    hp = dict(num_atom_types=1, **dict(PRODUCTION, radial_cutoff=5.0))
    torch.manual_seed(5)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp)).to(cuda_device)
    B, N = 48, 64
    x = torch.rand(B, N, 3)
    a = torch.randint(0, 2, (B, N))
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = 14.0
    sigma = torch.full((B, 1), 0.2)
    edges = pc_oracle.edges_with_radial_cutoff(x, torch.diag_embed(lattice[:, :3].clip(min=2.2 * 5.0)), 5.0)
    assert int(torch.bincount(edges[:, 0], minlength=B * N).max()) <= 32   # the premise, checked on the CPU edge list
    first = net(network_batch(a, x, lattice, sigma, cuda_device), conditional=False)
    assert net.last_number_of_edges() == edges.shape[0]
    x1, a1 = first.X.clone(), first.A.clone()
    for _ in range(3):
        again = net(network_batch(a, x, lattice, sigma, cuda_device), conditional=False)
        assert torch.equal(again.X, x1) and torch.equal(again.A, a1)

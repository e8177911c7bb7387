"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the UNMODIFIED reference on CPU.

Run here (the container that has /root/reference):   python oracle/make_golden.py
The fixtures are small (a few hundred KB), committed, and are what pins ``oracle/pc_oracle.py`` (CPU tests) and,
through it, the CUDA path (GPU tests).  Noise is injected through the reference's own ``_draw_*`` hooks from a
``NoiseTape`` (same mechanism as the reference's tests: tests/generators/test_langevin_generator.py:158-172).
"""
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shims  # noqa: E402
from oracle.pc_oracle import NoiseTape  # noqa: E402

ref_shims.install_shims()

from diffusion_for_multi_scale_molecular_dynamics.generators.constrained_langevin_generator import \
    ConstrainedLangevinGenerator  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.generators.langevin_generator import LangevinGenerator  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.generators.sampling_constraint import SamplingConstraint  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.models.egnn_utils import (get_edges_with_radial_cutoff,  # noqa: E402
                                                                             unsorted_segment_mean,
                                                                             unsorted_segment_sum)
from diffusion_for_multi_scale_molecular_dynamics.models.score_networks.analytical_score_network import (  # noqa: E402
    AnalyticalScoreNetwork, AnalyticalScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics.models.score_networks.egnn_score_network import (  # noqa: E402
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics.namespace import (AXL, CARTESIAN_FORCES, NOISE,  # noqa: E402
                                                                     NOISY_AXL_COMPOSITION, TIME)
from diffusion_for_multi_scale_molecular_dynamics.noise_schedulers.noise_parameters import NoiseParameters  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.noise_schedulers.noise_scheduler import NoiseScheduler  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.noisers.atom_types_noiser import AtomTypesNoiser  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.noisers.relative_coordinates_noiser import \
    RelativeCoordinatesNoiser  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.score.wrapped_gaussian_score import \
    get_coordinates_sigma_normalized_score  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.utils.basis_transformations import \
    map_relative_coordinates_to_unit_cell  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.utils.d3pm_utils import \
    get_probability_at_previous_time_step  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics.utils.neighbors import \
    get_periodic_adjacency_information  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from tests.egnn_variants import EGNN_VARIANTS  # noqa: E402


def to_np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **to_np(arrays))
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KB)")


def golden_noise_tables():
    out = {}
    for tag, kwargs, C in (("exp_T10_C2", dict(total_time_steps=10), 2),
                           ("lin_T7_C3", dict(total_time_steps=7, schedule_type="linear", sigma_min=1e-4, sigma_max=0.2,
                                              corrector_step_epsilon=2.5e-8), 3),
                           ("exp_T1000_C2", dict(total_time_steps=1000), 2)):
        noise, langevin = NoiseScheduler(NoiseParameters(**kwargs), num_classes=C).get_all_sampling_parameters()
        for field in ("time", "sigma", "sigma_squared", "g", "g_squared", "beta", "alpha_bar", "q_matrix", "q_bar_matrix",
                      "q_bar_tm1_matrix"):
            out[f"{tag}.{field}"] = getattr(noise, field)
        out[f"{tag}.epsilon"] = langevin.epsilon
        out[f"{tag}.sqrt_2_epsilon"] = langevin.sqrt_2_epsilon
    save("noise_tables", **out)


def golden_wrap_and_score():
    torch.manual_seed(11)
    x = torch.cat([torch.randn(200) * 3.0, torch.tensor([0.0, 1.0, -1.0, -1e-9, 1e-9, 0.99999994, -0.0, 2.5, -2.5,
                                                        -5.9604645e-08, 1.0000001, 123456.78])])
    out = dict(wrap_in=x, wrap_out=map_relative_coordinates_to_unit_cell(x.clone()))
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        u = torch.rand(400, dtype=dtype)
        u[:6] = torch.tensor([0.0, 0.5, 0.49999, 0.999, 1e-7, 0.25], dtype=dtype)
        sigma = torch.exp(torch.rand(400, dtype=dtype) * 9.0 - 7.0)  # 1e-3 .. 7
        sigma[:6] = torch.tensor([0.005, 0.3989422804014327, 0.39894, 0.5, 0.001, 2.0], dtype=dtype)
        for kmax in (1, 4):
            out[f"wg_{tag}_k{kmax}"] = get_coordinates_sigma_normalized_score(u, sigma, kmax=kmax)
        out[f"wg_{tag}_u"], out[f"wg_{tag}_sigma"] = u, sigma
    save("wrap_and_wrapped_gaussian", **out)


def golden_neighbors():
    torch.manual_seed(5)
    out = {}
    # triclinic-ish cells like tests/utils/test_neighbors.py:102-170
    B, N = 4, 14
    basis = torch.diag_embed(5.0 + 2.0 * torch.rand(B, 3)) + 0.4 * (torch.rand(B, 3, 3) - 0.5)
    rel = torch.rand(B, N, 3)
    cart = torch.matmul(rel, basis)
    for rc in (1.1, 2.2):
        adj = get_periodic_adjacency_information(cart, basis, rc)
        out[f"tri_rc{rc}.adjacency"] = adj.adjacency_matrix
        out[f"tri_rc{rc}.shifts"] = adj.shifts
        out[f"tri_rc{rc}.edge_batch"] = adj.edge_batch_indices
        out[f"tri_rc{rc}.number_of_edges"] = adj.number_of_edges
    out["tri.cart"], out["tri.basis"] = cart, basis
    # the EGNN form: relative coordinates in an orthogonal box, sorted unique pairs
    B, N = 3, 24
    lengths = torch.tensor([[7.1, 7.1, 7.1], [6.7, 8.0, 7.4], [9.0, 6.7, 6.9]])
    rel = torch.rand(B, N, 3)
    rel[0, 1] = rel[0, 0]  # coincident atoms: d2 == 0 to the home image, non-zero to the others
    cell = torch.diag_embed(lengths)
    out["egnn.rel"], out["egnn.cell"] = rel, cell
    out["egnn.edges_rc3.0"] = get_edges_with_radial_cutoff(rel, cell, 3.0)
    out["egnn.edges_rc3.3"] = get_edges_with_radial_cutoff(rel, cell, 3.3)
    save("neighbors", **out)


def golden_segment():
    torch.manual_seed(3)
    data = torch.randn(40, 5)
    ids = torch.randint(0, 9, (40,))
    save("segment", data=data, ids=ids, sum=unsorted_segment_sum(data, ids, 11), mean=unsorted_segment_mean(data, ids, 11))


def make_network(variant: str, seed: int):
    torch.manual_seed(seed)
    return EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS[variant]))


def golden_egnn():
    out = {}
    cases = dict(default_fc=(3, 8, 5.43), radial_sige=(2, 16, 7.0), all_options=(2, 12, 6.0), wide=(2, 20, 7.5))
    for variant, (B, N, a) in cases.items():
        net = make_network(variant, seed=100)
        num_atom_types = EGNN_VARIANTS[variant]["num_atom_types"]
        torch.manual_seed(200)
        x = torch.rand(B, N, 3)
        atom_types = torch.randint(0, num_atom_types + 1, (B, N))
        lattice = torch.zeros(B, 6)
        lattice[:, :3] = a
        sigma = torch.rand(B, 1) * 0.4 + 0.01
        batch = {NOISY_AXL_COMPOSITION: AXL(A=atom_types, X=x, L=lattice), TIME: torch.rand(B, 1), NOISE: sigma,
                 CARTESIAN_FORCES: torch.zeros_like(x)}
        with torch.no_grad():
            result = net(batch, conditional=False)
        for key, value in net.state_dict().items():
            out[f"{variant}.sd.{key}"] = value
        out[f"{variant}.x"], out[f"{variant}.a"], out[f"{variant}.l"], out[f"{variant}.sigma"] = x, atom_types, lattice, sigma
        out[f"{variant}.logits"], out[f"{variant}.score"] = result.A, result.X
    save("egnn_forward", **out)


def golden_atom_types():
    torch.manual_seed(9)
    out = {}
    B, N, C, T = 6, 7, 3, 8
    noise, _ = NoiseScheduler(NoiseParameters(total_time_steps=T), num_classes=C).get_all_sampling_parameters()
    sampling = PredictorCorrectorSamplingParameters(number_of_samples=B, number_of_atoms=N, num_atom_types=C - 1,
                                                    use_fixed_lattice_parameters=True, cell_dimensions=[5.0, 5.0, 5.0])
    gen = LangevinGenerator(NoiseParameters(total_time_steps=T), sampling, axl_network=None)
    logits = torch.randn(B, N, C)
    logits[..., -1] = -torch.inf
    a = torch.randint(0, C, (B, N))
    a[0] = C - 1  # an all-MASK sample
    a[1] = 0  # a fully unmasked sample
    gumbel_u = torch.rand(B, N, C)
    gumbel = -torch.log(-torch.log(gumbel_u.clip(min=1e-8)))
    binary = torch.rand(B, N)
    out.update(logits=logits, a=a, gumbel=gumbel, binary=binary)
    for idx in (0, 3, T - 1):
        q, qb, qbm = noise.q_matrix[idx], noise.q_bar_matrix[idx], noise.q_bar_tm1_matrix[idx]
        out[f"idx{idx}.q"], out[f"idx{idx}.q_bar"], out[f"idx{idx}.q_bar_tm1"] = q, qb, qbm
        expand = lambda m: m.expand(B, N, C, C)  # noqa: E731
        onehot = torch.nn.functional.one_hot(a, num_classes=C).float()
        out[f"idx{idx}.probs"] = get_probability_at_previous_time_step(logits, onehot, expand(q), expand(qb),
                                                                       expand(qbm), 1e-8, True)
        for greedy in (False, True):
            for one in (False, True):
                with mock.patch.object(gen, "_draw_gumbel_sample", return_value=gumbel.clone()), \
                        mock.patch.object(gen, "_draw_binary_sample", return_value=binary.clone()):
                    new = gen._atom_types_update(logits, a, expand(q), expand(qb), expand(qbm),
                                                 atom_type_greedy_sampling=greedy,
                                                 one_atom_type_transition_per_step=one)
                out[f"idx{idx}.g{int(greedy)}o{int(one)}"] = new
    save("atom_types", **out)


class TapedHooks:
    """Patch the reference's noise draws with a NoiseTape (per-kind streams)."""

    def __init__(self, generator, tape: NoiseTape, lattice_noise: bool = False):
        g = generator
        B_of = lambda n: n  # noqa: E731
        self.patches = [
            mock.patch.object(g, "_draw_coordinates_gaussian_sample",
                              side_effect=lambda n: tape.normal("coordinates", B_of(n), g.number_of_atoms, 3)),
            mock.patch.object(g, "_draw_lattice_gaussian_sample",
                              side_effect=(lambda n: tape.normal("lattice", n, g.num_lattice_parameters)) if lattice_noise
                              else (lambda n: torch.zeros(n, g.num_lattice_parameters))),
            mock.patch.object(g, "_draw_gumbel_sample",
                              side_effect=lambda n: tape.gumbel(g.small_epsilon, n, g.number_of_atoms, g.num_classes)),
            mock.patch.object(g, "_draw_binary_sample", side_effect=lambda n: tape.uniform("binary", n, g.number_of_atoms)),
            mock.patch.object(RelativeCoordinatesNoiser, "_get_gaussian_noise",
                              side_effect=lambda shape: tape.normal("repaint_normal", *shape)),
            mock.patch.object(AtomTypesNoiser, "_get_uniform_noise",
                              side_effect=lambda shape: tape.uniform("repaint_uniform", *shape)),
        ]

    def __enter__(self):
        for p in self.patches:
            p.start()
        return self

    def __exit__(self, *exc):
        for p in self.patches:
            p.stop()


def golden_trajectories():
    out = {}
    cases = {
        # name: (variant, B, N, cell, T, n_corr, sampling kwargs, constrained?)
        "si_fc": ("default_fc", 4, 8, 5.43, 6, 1, dict(), False),
        "sige_radial": ("radial_sige", 3, 16, 7.0, 5, 2, dict(), False),
        "sige_free": ("radial_sige", 3, 16, 7.0, 4, 1, dict(one_atom_type_transition_per_step=False,
                                                             atom_type_greedy_sampling=False,
                                                             atom_type_transition_in_corrector=True), False),
        "repaint": ("radial_sige", 3, 16, 7.0, 5, 1, dict(), True),
    }
    for name, (variant, B, N, a, T, n_corr, kwargs, constrained) in cases.items():
        net = make_network(variant, seed=300)
        num_atom_types = EGNN_VARIANTS[variant]["num_atom_types"]
        noise_parameters = NoiseParameters(total_time_steps=T, sigma_min=0.005, sigma_max=0.5)
        sampling = PredictorCorrectorSamplingParameters(
            number_of_samples=B, number_of_atoms=N, num_atom_types=num_atom_types, number_of_corrector_steps=n_corr,
            use_fixed_lattice_parameters=True, cell_dimensions=[a, a, a], **kwargs)
        torch.manual_seed(400)
        x0 = torch.rand(B, N, 3)
        a0 = torch.full((B, N), num_atom_types, dtype=torch.long)
        lattice = torch.zeros(B, 6)
        lattice[:, :3] = a
        if constrained:
            nc = 5
            constraint = SamplingConstraint(elements=["Si", "Ge"], constrained_relative_coordinates=torch.rand(nc, 3),
                                            constrained_atom_types=torch.randint(0, num_atom_types, (nc,)),
                                            constrained_indices=torch.tensor([2, 3, 7, 11, 12]))
            gen = ConstrainedLangevinGenerator(noise_parameters, sampling, net, constraint)
            out[f"{name}.c_x0"] = constraint.constrained_relative_coordinates
            out[f"{name}.c_a0"] = constraint.constrained_atom_types
            out[f"{name}.c_idx"] = constraint.constrained_indices
        else:
            gen = LangevinGenerator(noise_parameters, sampling, net)
        tape = NoiseTape(seed=77)
        with torch.no_grad(), TapedHooks(gen, tape):
            final = gen.sample_from_noisy_composition(AXL(A=a0.clone(), X=x0.clone(), L=lattice.clone()), T, 0)
            if constrained:
                final = gen._apply_constraint(final, torch.device("cpu"))
        # one isolated predictor step and one corrector step from (a0, x0) at a mid index, fresh tape
        mid = max(2, T // 2)
        tape = NoiseTape(seed=78)
        if not constrained:
            with torch.no_grad(), TapedHooks(gen, tape):
                p = gen.predictor_step(AXL(A=a0.clone(), X=x0.clone(), L=lattice.clone()), mid, torch.zeros_like(x0))
                c = gen.corrector_step(AXL(A=p.A.clone(), X=p.X.clone(), L=lattice.clone()), mid - 1, torch.zeros_like(x0))
            out[f"{name}.pred_a"], out[f"{name}.pred_x"] = p.A, p.X
            out[f"{name}.corr_a"], out[f"{name}.corr_x"] = c.A, c.X
            out[f"{name}.mid"] = mid
        for key, value in net.state_dict().items():
            out[f"{name}.sd.{key}"] = value
        out[f"{name}.x0"], out[f"{name}.a0"], out[f"{name}.l"] = x0, a0, lattice
        out[f"{name}.final_a"], out[f"{name}.final_x"] = final.A, final.X
        out[f"{name}.T"], out[f"{name}.n_corr"] = T, n_corr
    save("trajectories", **out)


DIAMOND = [[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.25, .25, .25], [.25, .75, .75], [.75, .25, .75],
           [.75, .75, .25]]


def golden_analytical():
    """AnalyticalScoreNetwork forward and a whole PC trajectory driven by it (Si 1x1x1 diamond, sigma_d = 0.01)."""
    out = {}
    params = AnalyticalScoreNetworkParameters(number_of_atoms=8, num_atom_types=1, kmax=4,
                                              equilibrium_relative_coordinates=DIAMOND, sigma_d=0.01)
    net = AnalyticalScoreNetwork(params)
    torch.manual_seed(21)
    B = 6
    x = torch.rand(B, 8, 3)
    sigma = torch.tensor([0.005, 0.02, 0.1, 0.3, 0.45, 0.5]).reshape(B, 1)
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = 5.43
    batch = {NOISY_AXL_COMPOSITION: AXL(A=torch.ones(B, 8, dtype=torch.long), X=x, L=lattice), TIME: torch.rand(B, 1),
             NOISE: sigma, CARTESIAN_FORCES: torch.zeros_like(x)}
    result = net(batch, conditional=False)
    out.update(x=x, sigma=sigma, score=result.X, logits=result.A, x0=torch.tensor(DIAMOND))
    T, n_corr = 25, 1
    sampling = PredictorCorrectorSamplingParameters(number_of_samples=B, number_of_atoms=8, num_atom_types=1,
                                                    number_of_corrector_steps=n_corr, use_fixed_lattice_parameters=True,
                                                    cell_dimensions=[5.43] * 3)
    gen = LangevinGenerator(NoiseParameters(total_time_steps=T), sampling, net)
    x_start = torch.rand(B, 8, 3)
    a_start = torch.ones(B, 8, dtype=torch.long)
    with torch.no_grad(), TapedHooks(gen, NoiseTape(seed=31)):
        final = gen.sample_from_noisy_composition(AXL(A=a_start.clone(), X=x_start.clone(), L=lattice.clone()), T, 0)
    out.update(traj_x0=x_start, traj_final_x=final.X, traj_final_a=final.A, T=T)
    save("analytical", **out)


def golden_pair_distances():
    """compute_distances_in_batch (utils/structure_utils.py:41-121) of the unmodified reference."""
    from diffusion_for_multi_scale_molecular_dynamics.utils.structure_utils import compute_distances_in_batch
    torch.manual_seed(21)
    out = {}
    B, N = 4, 14
    basis = torch.diag_embed(5.0 + 2.0 * torch.rand(B, 3)) + 0.4 * (torch.rand(B, 3, 3) - 0.5)
    cart = torch.matmul(torch.rand(B, N, 3), basis)
    out["tri.cart"], out["tri.basis"] = cart, basis
    for rc in (1.7, 2.6):
        out[f"tri.distances_rc{rc}"] = compute_distances_in_batch(cart, basis, rc)
    # thermally displaced diamond silicon, 2x2x2 conventional cells
    basis_atoms = torch.tensor([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.25, .25, .25], [.25, .75, .75],
                                [.75, .25, .75], [.75, .75, .25]])
    grid = torch.stack(torch.meshgrid(*[torch.arange(2)] * 3, indexing="ij"), dim=-1).reshape(-1, 1, 3)
    rel = ((grid + basis_atoms.unsqueeze(0)) / 2).reshape(1, -1, 3).repeat(3, 1, 1)
    rel = torch.remainder(rel + 0.01 * torch.randn_like(rel), 1.0)
    cell = torch.diag_embed(torch.full((3, 3), 10.86))
    cart = torch.matmul(rel, cell)
    out["si.cart"], out["si.basis"] = cart, cell
    out["si.distances_rc5.0"] = compute_distances_in_batch(cart, cell, 5.0)
    save("pair_distances", **out)


def golden_adaptive_corrector():
    """AdaptiveCorrectorGenerator (generators/adaptive_corrector.py:17-148) of the unmodified reference: the step-size
    formula on random inputs, and a short corrector-only trajectory with injected noise."""
    from diffusion_for_multi_scale_molecular_dynamics.generators.adaptive_corrector import AdaptiveCorrectorGenerator
    out = {}
    variant, B, N, a, T, n_corr = "radial_sige", 3, 16, 7.0, 5, 2
    net = make_network(variant, seed=310)
    num_atom_types = EGNN_VARIANTS[variant]["num_atom_types"]
    noise_parameters = NoiseParameters(total_time_steps=T, sigma_min=0.005, sigma_max=0.5, corrector_r=0.17)
    sampling = PredictorCorrectorSamplingParameters(
        algorithm="adaptive_corrector", number_of_samples=B, number_of_atoms=N, num_atom_types=num_atom_types,
        number_of_corrector_steps=n_corr, use_fixed_lattice_parameters=True, cell_dimensions=[a, a, a])
    gen = AdaptiveCorrectorGenerator(noise_parameters, sampling, net)
    torch.manual_seed(410)
    for tag, shape, n_dim, sigma in (("x", (5, 12, 3), 3, 0.07), ("l", (5, 6), 2, 0.03), ("tiny", (2, 4, 3), 3, 0.5)):
        scores = torch.randn(*shape) * (1e-12 if tag == "tiny" else 0.3)   # "tiny" exercises the small_epsilon clip
        z = torch.randn(*shape)
        out[f"step.{tag}.scores"], out[f"step.{tag}.z"], out[f"step.{tag}.sigma"] = scores, z, torch.tensor(sigma)
        out[f"step.{tag}.eps"] = gen._generic_corrector_step_size(2, torch.tensor(sigma), scores, z, n_dim=n_dim)
    x0 = torch.rand(B, N, 3)
    a0 = torch.full((B, N), num_atom_types, dtype=torch.long)
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = a
    tape = NoiseTape(seed=79)
    with torch.no_grad(), TapedHooks(gen, tape):
        final = gen.sample_from_noisy_composition(AXL(A=a0.clone(), X=x0.clone(), L=lattice.clone()), T, 0)
    for key, value in net.state_dict().items():
        out[f"traj.sd.{key}"] = value
    out["traj.x0"], out["traj.a0"], out["traj.l"] = x0, a0, lattice
    out["traj.final_a"], out["traj.final_x"] = final.A, final.X
    out["traj.T"], out["traj.n_corr"] = T, n_corr
    save("adaptive_corrector", **out)


def golden_force_field():
    """ForceFieldAugmentedScoreNetwork (models/score_networks/force_field_augmented_score_network.py) of the reference:
    the relative-coordinate pseudo force alone and on top of an EGNN score."""
    from diffusion_for_multi_scale_molecular_dynamics.models.score_networks.force_field_augmented_score_network import (
        ForceFieldAugmentedScoreNetwork, ForceFieldParameters)
    out = {}
    net = make_network("radial_sige", seed=320)
    torch.manual_seed(420)
    B, N = 4, 16
    x = torch.rand(B, N, 3)
    x[0, 1] = x[0, 0] + torch.tensor([0.01, 0.0, 0.0])  # a nearly coincident pair: large repulsion
    x[1, 3] = x[1, 2]                                    # exactly coincident atoms: d2 == 0 to the home image
    a = torch.randint(0, 3, (B, N))
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = torch.tensor([[7.0, 7.0, 7.0], [6.5, 7.5, 8.0], [7.2, 6.8, 7.0], [9.0, 7.0, 6.6]])
    sigma = torch.rand(B, 1) * 0.3 + 0.01
    batch = {NOISY_AXL_COMPOSITION: AXL(A=a, X=x, L=lattice), TIME: torch.rand(B, 1), NOISE: sigma,
             CARTESIAN_FORCES: torch.zeros(B, N, 3)}
    for rc, strength in ((1.5, 1.0), (2.5, 0.3)):
        ff = ForceFieldAugmentedScoreNetwork(net, ForceFieldParameters(radial_cutoff=rc, strength=strength))
        with torch.no_grad():
            out[f"rc{rc}.force"] = ff.get_relative_coordinates_pseudo_force(batch)
            augmented = ff(batch, conditional=False)
        out[f"rc{rc}.augmented_x"], out[f"rc{rc}.augmented_a"] = augmented.X, augmented.A
        out[f"rc{rc}.strength"] = strength
    for key, value in net.state_dict().items():
        out[f"sd.{key}"] = value
    out["x"], out["a"], out["l"], out["sigma"] = x, a, lattice, sigma
    save("force_field", **out)


def golden_free_lattice():
    """LangevinGenerator with use_fixed_lattice_parameters=False: the lattice parameters diffuse too (the EGNN lattice
    score is zero, so they move by the injected noise) and the radius graph follows the changing box."""
    out = {}
    variant, B, N, a, T, n_corr = "radial_sige", 3, 16, 7.0, 5, 1
    net = make_network(variant, seed=330)
    num_atom_types = EGNN_VARIANTS[variant]["num_atom_types"]
    noise_parameters = NoiseParameters(total_time_steps=T, sigma_min=0.005, sigma_max=0.5)
    sampling = PredictorCorrectorSamplingParameters(
        number_of_samples=B, number_of_atoms=N, num_atom_types=num_atom_types, number_of_corrector_steps=n_corr,
        use_fixed_lattice_parameters=False, cell_dimensions=[a, a, a])
    gen = LangevinGenerator(noise_parameters, sampling, net)
    torch.manual_seed(430)
    x0 = torch.rand(B, N, 3)
    a0 = torch.full((B, N), num_atom_types, dtype=torch.long)
    lattice = torch.zeros(B, 6)
    lattice[:, :3] = torch.tensor([[7.0, 7.2, 6.9], [7.4, 7.0, 7.1], [6.8, 7.3, 7.0]])
    with torch.no_grad(), TapedHooks(gen, NoiseTape(seed=80), lattice_noise=True):
        final = gen.sample_from_noisy_composition(AXL(A=a0.clone(), X=x0.clone(), L=lattice.clone()), T, 0)
    with torch.no_grad(), TapedHooks(gen, NoiseTape(seed=81), lattice_noise=True):
        p = gen.predictor_step(AXL(A=a0.clone(), X=x0.clone(), L=lattice.clone()), 3, torch.zeros_like(x0))
        c = gen.corrector_step(AXL(A=p.A.clone(), X=p.X.clone(), L=p.L.clone()), 2, torch.zeros_like(x0))
    for key, value in net.state_dict().items():
        out[f"sd.{key}"] = value
    out["x0"], out["a0"], out["l"] = x0, a0, lattice
    out["final_a"], out["final_x"], out["final_l"] = final.A, final.X, final.L
    out["pred_a"], out["pred_x"], out["pred_l"] = p.A, p.X, p.L
    out["corr_a"], out["corr_x"], out["corr_l"] = c.A, c.X, c.L
    out["T"], out["n_corr"] = T, n_corr
    save("free_lattice", **out)


def golden_checkpoint():
    """A checkpoint shaped like the reference's Lightning checkpoints: ``state_dict`` keys under ``axl_network.`` and the
    reference's own parameter dataclass pickled (by class name) inside ``hyper_parameters``."""
    net = make_network("radial_sige", seed=500)
    state = {"axl_network." + k: v for k, v in net.state_dict().items()}
    state["loss_calculator.weights"] = torch.ones(3)  # something that is not part of the score network
    hyper = {"hyper_params": {"score_network_parameters": EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]),
                              "kmax_target_score": 4}}
    path = os.path.join(OUT, "lightning_like_checkpoint.ckpt")
    torch.save({"state_dict": state, "hyper_parameters": hyper, "epoch": 3, "pytorch-lightning_version": "2.2.0"}, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KB)")


if __name__ == "__main__":
    torch.set_num_threads(4)
    ALL = [golden_noise_tables, golden_wrap_and_score, golden_neighbors, golden_segment, golden_egnn, golden_atom_types,
           golden_trajectories, golden_analytical, golden_checkpoint, golden_pair_distances, golden_adaptive_corrector, golden_force_field, golden_free_lattice]
    wanted = sys.argv[1:]   # e.g. `python oracle/make_golden.py pair_distances`; no argument = everything
    for fn in ALL:
        if not wanted or fn.__name__.replace("golden_", "") in wanted:
            fn()

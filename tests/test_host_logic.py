"""CPU: host-side mirror of the reference interface -- noise tables, dataclasses, Bloch vectors, generator plumbing."""
import dataclasses
from unittest import mock

import pytest
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.constrained_langevin_generator import \
    ConstrainedLangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.instantiate_generator import instantiate_generator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.load_sampling_parameters import \
    load_sampling_parameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.sampling_constraint import (
    SamplingConstraint, read_sampling_constraint, write_sampling_constraint)
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import (
    FullRandomTrajectoryInitializer, instantiate_trajectory_initializer)
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network_factory import (
    SCORE_NETWORKS_BY_ARCH, create_score_network, create_score_network_parameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_scheduler import NoiseScheduler
from tests.conftest import load_golden, sub_dict
from tests.egnn_variants import EGNN_VARIANTS


@pytest.mark.parametrize("tag,kwargs,C", [
    ("exp_T10_C2", dict(total_time_steps=10), 2),
    ("lin_T7_C3", dict(total_time_steps=7, schedule_type="linear", sigma_min=1e-4, sigma_max=0.2,
                       corrector_step_epsilon=2.5e-8), 3),
    ("exp_T1000_C2", dict(total_time_steps=1000), 2),
])
def test_noise_scheduler_tables_are_bit_identical_to_the_reference(tag, kwargs, C):
    golden = sub_dict(load_golden("noise_tables"), tag)
    noise, langevin = NoiseScheduler(NoiseParameters(**kwargs), num_classes=C).get_all_sampling_parameters()
    for field in ("time", "sigma", "sigma_squared", "g", "g_squared", "beta", "alpha_bar", "q_matrix", "q_bar_matrix",
                  "q_bar_tm1_matrix"):
        assert torch.equal(getattr(noise, field), golden[field]), field
    assert torch.equal(langevin.epsilon, golden["epsilon"])
    assert torch.equal(langevin.sqrt_2_epsilon, golden["sqrt_2_epsilon"])


def test_unknown_schedule_is_refused():
    with pytest.raises(AssertionError):
        NoiseParameters(total_time_steps=5, schedule_type="cosine")


@pytest.mark.parametrize("variant", list(EGNN_VARIANTS))
def test_egnn_module_tree_matches_the_reference_state_dict(variant):
    """Same parameter names, shapes, and -- for the frozen Bloch / projection tensors -- values."""
    reference = sub_dict(sub_dict(load_golden("egnn_forward"), variant), "sd")
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS[variant]))
    own = net.state_dict()
    assert list(own) == list(reference)
    for key, value in own.items():
        assert value.shape == reference[key].shape, key
    assert torch.equal(own["bloch_wave_reciprocal_lattice_vectors"], reference["bloch_wave_reciprocal_lattice_vectors"])
    assert torch.equal(own["projection_matrices"], reference["projection_matrices"])
    net.load_state_dict(reference, strict=True)


def test_same_seed_gives_the_reference_initial_weights():
    """Module construction order equals the reference's, so torch.manual_seed reproduces its random init."""
    reference = sub_dict(sub_dict(load_golden("egnn_forward"), "radial_sige"), "sd")
    torch.manual_seed(100)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"]))
    for key, value in net.state_dict().items():
        assert torch.equal(value, reference[key]), key


def test_score_network_factory():
    assert SCORE_NETWORKS_BY_ARCH["egnn"] is EGNNScoreNetwork
    params = create_score_network_parameters(dict(architecture="egnn", n_layers=2), dict(num_atom_types=2, unrelated=1))
    assert isinstance(params, EGNNScoreNetworkParameters) and params.num_atom_types == 2 and params.n_layers == 2
    assert isinstance(create_score_network(params), EGNNScoreNetwork)
    with pytest.raises(NotImplementedError):
        create_score_network_parameters(dict(architecture="mace"), dict(num_atom_types=1))
    with pytest.raises(AssertionError):
        EGNNScoreNetwork(EGNNScoreNetworkParameters(num_atom_types=1, edges="radial_cutoff"))  # cutoff missing
    with pytest.raises(AssertionError):
        EGNNScoreNetwork(EGNNScoreNetworkParameters(num_atom_types=1, edges="fully_connected", radial_cutoff=3.0))


def sampling(**kwargs):
    base = dict(number_of_samples=4, number_of_atoms=8, num_atom_types=1, use_fixed_lattice_parameters=True,
                cell_dimensions=[5.43, 5.43, 5.43])
    base.update(kwargs)
    return PredictorCorrectorSamplingParameters(**base)


def test_sampling_parameters_mirror_the_reference_defaults():
    p = sampling()
    assert p.algorithm == "predictor_corrector" and p.number_of_corrector_steps == 1 and p.small_epsilon == 1e-8
    assert p.one_atom_type_transition_per_step and p.atom_type_greedy_sampling
    assert not p.atom_type_transition_in_corrector and p.sample_batchsize is None
    assert torch.allclose(p.fixed_lattice_parameters, torch.tensor([5.43, 5.43, 5.43, 0, 0, 0]))
    with pytest.raises(AssertionError):
        sampling(cell_dimensions=None)
    loaded = load_sampling_parameters(dict(algorithm="predictor_corrector", **{
        k: v for k, v in dataclasses.asdict(p).items() if k not in ("algorithm",)}))
    assert loaded == p
    adaptive = load_sampling_parameters(dict(algorithm="adaptive_corrector", **{
        k: v for k, v in dataclasses.asdict(p).items() if k not in ("algorithm",)}))
    assert adaptive.algorithm == "adaptive_corrector"
    with pytest.raises(NotImplementedError):  # ODE / SDE solvers are outside the accelerated path
        load_sampling_parameters(dict(algorithm="ode"))
    with pytest.raises(AssertionError):
        load_sampling_parameters(dict(algorithm="euler"))


def test_generator_plumbing_without_a_device():
    noise = NoiseParameters(total_time_steps=5)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(num_atom_types=1))
    gen = instantiate_generator(sampling(), noise, net)
    assert type(gen) is LangevinGenerator and isinstance(gen.trajectory_initializer, FullRandomTrajectoryInitializer)
    assert gen.num_classes == 2 and gen.masked_atom_type_index == 1 and gen.number_of_discretization_steps == 5
    assert not gen._hooks_overridden()
    with mock.patch.object(gen, "_draw_gumbel_sample", return_value=None):
        assert gen._hooks_overridden()  # an injected hook must force the step-by-step loop
    assert not gen._hooks_overridden()

    class Tweaked(LangevinGenerator):
        def _get_coordinates_corrector_step_size(self, index_i, sigma_i, model_predictions_i, z):
            return 0.5

    assert Tweaked(noise, sampling(), net)._hooks_overridden()
    constraint = SamplingConstraint(elements=["Si"], constrained_relative_coordinates=torch.rand(3, 3),
                                    constrained_atom_types=torch.zeros(3, dtype=torch.long))
    constrained = instantiate_generator(sampling(), noise, net, sampling_constraints=constraint)
    assert type(constrained) is ConstrainedLangevinGenerator and not constrained._hooks_overridden()
    assert torch.equal(constrained.constraint_indices, torch.arange(3))
    with pytest.raises(NotImplementedError):
        instantiate_generator(dataclasses.replace(sampling(), algorithm="ode"), noise, net)
    with pytest.raises(AssertionError):
        gen.predictor_step(AXL(A=torch.zeros(1, 8), X=torch.zeros(1, 8, 3), L=torch.zeros(1, 6)), 0, None)
    with pytest.raises(AssertionError):
        gen.sample_from_noisy_composition(AXL(A=None, X=torch.zeros(1, 8, 3), L=None), 2, 2)
    with pytest.raises(_cabi.B200MDError):  # no CPU path
        gen.sample(2, torch.device("cpu"))
    with pytest.raises(AssertionError):
        LangevinGenerator(noise, sampling(record_samples_corrector_steps=True), net)


def test_sampling_constraint_validation_and_round_trip(tmp_path):
    constraint = SamplingConstraint(elements=["Si", "Ge"], constrained_relative_coordinates=torch.rand(2, 3),
                                    constrained_atom_types=torch.tensor([0, 1]),
                                    constrained_indices=torch.tensor([4, 1]))
    write_sampling_constraint(constraint, tmp_path / "c.pt")
    again = read_sampling_constraint(tmp_path / "c.pt")
    assert torch.equal(again.constrained_relative_coordinates, constraint.constrained_relative_coordinates)
    assert torch.equal(again.constrained_indices, constraint.constrained_indices) and again.elements == ["Si", "Ge"]
    with pytest.raises(AssertionError):
        SamplingConstraint(elements=["Si"], constrained_relative_coordinates=torch.rand(2, 3),
                           constrained_atom_types=torch.tensor([0, 1]))  # type 1 not in elements
    with pytest.raises(AssertionError):
        SamplingConstraint(elements=["Si"], constrained_relative_coordinates=torch.rand(2, 3).double(),
                           constrained_atom_types=torch.tensor([0, 0]))


def test_trajectory_initializer_factory(tmp_path):
    p = sampling()
    assert isinstance(instantiate_trajectory_initializer(p), FullRandomTrajectoryInitializer)
    start = AXL(A=torch.zeros(4, 8, dtype=torch.long), X=torch.rand(4, 8, 3), L=torch.zeros(4, 6))
    torch.save({"noisy_axl": start, "start_time_step_index": 3}, tmp_path / "start.pt")
    init = instantiate_trajectory_initializer(p, str(tmp_path / "start.pt"))
    assert init.create_start_time_step_index(1000) == 3 and init.create_end_time_step_index() == 0
    assert torch.equal(init.initialize(4, torch.device("cpu")).X, start.X)
    with pytest.raises(AssertionError):
        init.initialize(5, torch.device("cpu"))


def test_checkpoint_unpickler_never_executes_named_globals(tmp_path):
    """A checkpoint that names an arbitrary callable (``os.system``) loads as an inert holder: only the reference package
    (mapped to this package or to an opaque stand-in) and an allow-list of tensor / container globals are resolved."""
    import io
    import os
    import pickle

    from diffusion_for_multi_scale_molecular_dynamics_b200 import sample_diffusion

    marker = tmp_path / "executed"

    class Evil:
        def __reduce__(self):
            return (os.system, (f"touch {marker}",))

    buffer = io.BytesIO()
    pickle.dump({"payload": Evil(), "fine": {"a": 1}}, buffer)
    buffer.seek(0)
    loaded = sample_diffusion._TolerantUnpickler(buffer).load()
    assert not marker.exists() and loaded["fine"] == {"a": 1}
    assert isinstance(loaded["payload"], sample_diffusion._Opaque)
    # the Lightning-like fixture (reference classes + tensors) still loads, and a missing file is an I/O error, not a fallback
    checkpoint = sample_diffusion._load_checkpoint_file(os.path.join(os.path.dirname(__file__), "golden",
                                                                    "lightning_like_checkpoint.ckpt"))
    assert "state_dict" in checkpoint and sample_diffusion._find_score_network_parameters(checkpoint) is not None
    with pytest.raises(FileNotFoundError):
        sample_diffusion._load_checkpoint_file(tmp_path / "missing.ckpt")


def test_auto_implementation_widens_its_range_once(monkeypatch):
    """Host logic of the fp16-overflow fallback (no GPU): under IMPL_AUTO a TC_RANGE bit in the status word switches the
    network to the 3xTF32 kernels for good, clears the word (bits raised downstream of the overflow are stale) and asks
    the caller to evaluate again; a forced implementation keeps the bit, which is then raised as an AssertionError."""
    from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
    from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
        EGNNScoreNetwork, EGNNScoreNetworkParameters)
    hp = dict(num_atom_types=1, edges="radial_cutoff", radial_cutoff=3.0, message_hidden_dimensions_size=64,
              coordinate_hidden_dimensions_size=64, node_hidden_dimensions_size=64, n_layers=1)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp))
    assert net.effective_kernel_impl() == _cabi.IMPL_AUTO and "3xFP16" in net.last_edge_kernel_name()
    net._status = torch.tensor([_cabi.STATUS_TC_RANGE | _cabi.STATUS_BAD_LOGITS], dtype=torch.int32)
    assert net.widen_range_after_overflow() is True
    assert net.effective_kernel_impl() == _cabi.IMPL_TCGEN05 and "3xTF32" in net.last_edge_kernel_name()
    assert int(net._status) == 0
    net._status.fill_(_cabi.STATUS_TC_RANGE)
    assert net.widen_range_after_overflow() is False   # already on the wide-range kernels: nothing left to fall back to
    with pytest.raises(AssertionError, match="fp16 range"):
        net.raise_on_device_status()
    forced = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp))
    forced.kernel_impl = _cabi.IMPL_TCGEN05_F16X3_TS
    forced._status = torch.tensor([_cabi.STATUS_TC_RANGE], dtype=torch.int32)
    assert forced.widen_range_after_overflow() is False and int(forced._status) == _cabi.STATUS_TC_RANGE
    narrow = EGNNScoreNetwork(EGNNScoreNetworkParameters(num_atom_types=1))   # default widths: FFMA kernels, no fp16 anywhere
    narrow._status = torch.tensor([_cabi.STATUS_TC_RANGE], dtype=torch.int32)
    assert narrow.widen_range_after_overflow() is False


def test_status_word_helpers():
    from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
    assert _cabi.status_word(None) == 0
    status = torch.tensor([_cabi.STATUS_X_OUT_OF_RANGE | _cabi.STATUS_MASK_REMAINS], dtype=torch.int32)
    assert _cabi.status_word(status) == 17
    with pytest.raises(AssertionError, match=r"\[0,1\)") as info:
        _cabi.raise_for_word(17, status)
    assert "MASKED" in str(info.value) and int(status) == 0
    _cabi.raise_for_word(0, status)   # nothing set: no exception

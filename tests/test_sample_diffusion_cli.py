"""The sampling entry point (reference sample_diffusion.py): checkpoint loading on CPU, end-to-end run on the GPU."""
import os
import subprocess
import sys

import pytest
import torch
import yaml

from diffusion_for_multi_scale_molecular_dynamics_b200 import sample_diffusion
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.sampling_constraint import (
    SamplingConstraint, write_sampling_constraint)
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL_COMPOSITION, CARTESIAN_POSITIONS
from tests.conftest import GOLDEN
from tests.egnn_variants import EGNN_VARIANTS

CHECKPOINT = os.path.join(GOLDEN, "lightning_like_checkpoint.ckpt")
CONFIG = dict(
    noise=dict(total_time_steps=6, sigma_min=0.005, sigma_max=0.5),
    sampling=dict(algorithm="predictor_corrector", number_of_samples=10, sample_batchsize=4, number_of_atoms=16,
                  num_atom_types=2, number_of_corrector_steps=1, use_fixed_lattice_parameters=True,
                  cell_dimensions=[7.0, 7.0, 7.0]),
)


def test_lightning_style_checkpoint_is_read_without_the_reference_package():
    """The fixture pickles the reference's EGNNScoreNetworkParameters by class name and prefixes the weights with
    'axl_network.' like AXLDiffusionLightningModel does; neither lightning nor the reference is importable here."""
    assert "diffusion_for_multi_scale_molecular_dynamics" not in sys.modules
    network = sample_diffusion.get_axl_network(CHECKPOINT)
    assert isinstance(network, EGNNScoreNetwork)
    assert network.edges == "radial_cutoff" and network.radial_cutoff == 3.0 and network.num_atom_types == 2
    torch.manual_seed(500)
    expected = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["radial_sige"])).state_dict()
    for key, value in network.state_dict().items():
        assert torch.equal(value, expected[key]), key


def test_plain_state_dict_with_the_architecture_in_the_yaml(tmp_path):
    torch.manual_seed(1)
    net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**EGNN_VARIANTS["default_fc"]))
    torch.save(net.state_dict(), tmp_path / "weights.pt")
    hyper = dict(spatial_dimension=3, model=dict(score_network=dict(architecture="egnn", num_atom_types=1)))
    loaded = sample_diffusion.get_axl_network(tmp_path / "weights.pt", hyper)
    for key, value in loaded.state_dict().items():
        assert torch.equal(value, net.state_dict()[key])
    with pytest.raises(AssertionError):
        sample_diffusion.get_axl_network(tmp_path / "weights.pt", dict(noise={}))


def test_parameter_extraction():
    noise, sampling = sample_diffusion.extract_and_validate_parameters(CONFIG)
    assert noise.total_time_steps == 6 and sampling.number_of_samples == 10 and sampling.sample_batchsize == 4
    with pytest.raises(AssertionError):
        sample_diffusion.extract_and_validate_parameters(dict(noise=CONFIG["noise"]))


@pytest.mark.gpu
@pytest.mark.parametrize("constrained", [False, True])
def test_cli_end_to_end(tmp_path, constrained):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    config_path = tmp_path / "config.yaml"
    config = dict(CONFIG)
    config["sampling"] = dict(CONFIG["sampling"], record_samples=not constrained)
    config_path.write_text(yaml.dump(config))
    args = ["--config", str(config_path), "--checkpoint", CHECKPOINT, "--output", str(tmp_path / "out")]
    if constrained:
        constraint = SamplingConstraint(elements=["Si", "Ge"], constrained_relative_coordinates=torch.rand(3, 3),
                                        constrained_atom_types=torch.tensor([0, 1, 0]),
                                        constrained_indices=torch.tensor([1, 5, 9]))
        write_sampling_constraint(constraint, tmp_path / "constraint.pt")
        args += ["--path_to_sampling_constraint_data_pickle", str(tmp_path / "constraint.pt")]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-m", "diffusion_for_multi_scale_molecular_dynamics_b200.sample_diffusion"] + args,
                   check=True, cwd=root, timeout=300)
    samples = torch.load(tmp_path / "out" / "samples.pt", weights_only=False)
    axl = samples[AXL_COMPOSITION]
    assert axl.X.shape == (10, 16, 3) and axl.A.shape == (10, 16) and axl.L.shape == (10, 6)
    assert bool(((axl.X >= 0) & (axl.X < 1)).all()) and bool((axl.A < 2).all())
    torch.testing.assert_close(samples[CARTESIAN_POSITIONS], axl.X * 7.0)
    plain = torch.load(tmp_path / "out" / "samples_plain.pt", weights_only=True)
    assert torch.equal(plain["relative_coordinates"], axl.X)
    assert os.path.isfile(tmp_path / "out" / "config_backup.yaml")
    if constrained:
        assert torch.equal(axl.X[:, [1, 5, 9]], constraint.constrained_relative_coordinates.expand(10, 3, 3))
        assert torch.equal(axl.A[:, [1, 5, 9]], constraint.constrained_atom_types.expand(10, 3))
    else:
        trajectory = torch.load(tmp_path / "out" / "trajectories.pt", weights_only=False)
        assert len(trajectory["predictor_step"]) == 6 * 3  # 6 steps for each of the 3 chunks (4 + 4 + 2 samples)

"""Sampling entry point (reference: sample_diffusion.py:52-270): draw samples from a trained score network.

Same command line as the reference script --

    python -m diffusion_for_multi_scale_molecular_dynamics_b200.sample_diffusion \
        --config config.yaml --checkpoint last.ckpt --output out_dir [--device cuda]
        [--path_to_starting_configuration_data_pickle start.pt] [--path_to_sampling_constraint_data_pickle constraint.pt]

-- and the same outputs: ``samples.pt`` (``{cartesian_positions, original_axl}``), ``trajectories.pt`` when
``record_samples`` is set, plus ``samples_plain.pt`` (the same tensors in a plain dict that needs no package to load).
The yaml needs the reference's ``noise`` and ``sampling`` blocks; the network architecture comes from the checkpoint's
hyper-parameters or, failing that, from a ``model: score_network:`` block of the yaml (as in the training configs).
Under ``torchrun`` (one process per GPU) the samples are sharded across ranks and rank 0 writes the files.

A ``force_field`` block wraps the network in ``ForceFieldAugmentedScoreNetwork`` as in the reference.  Not carried over:
the LAMMPS energy oracle (``energies.pt``) and ORION reporting -- they are outside the accelerated path; an ``oracle``
block is reported and ignored.
"""
import argparse
import dataclasses
import logging
import os
import pickle
import shutil
import socket
from pathlib import Path
from typing import Any, AnyStr, Dict, Optional, Union

import torch
import torch.distributed as dist
import yaml

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.axl_generator import SamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.instantiate_generator import instantiate_generator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.load_sampling_parameters import \
    load_sampling_parameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.sampling_constraint import read_sampling_constraint
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import \
    instantiate_trajectory_initializer
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.force_field_augmented_score_network import (
    ForceFieldAugmentedScoreNetwork, ForceFieldParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import ScoreNetwork
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network_factory import (
    SCORE_NETWORK_PARAMETERS_BY_ARCH, create_score_network, create_score_network_parameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL_COMPOSITION, CARTESIAN_POSITIONS
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.sampling.diffusion_sampling import create_batch_of_samples

logger = logging.getLogger(__name__)

REFERENCE_PACKAGE = "diffusion_for_multi_scale_molecular_dynamics"
AXL_NETWORK_PREFIX = "axl_network."


class _Opaque:
    """Stand-in for classes of the reference that a Lightning checkpoint pickles by name (its hyper-parameter
    dataclasses).  Keeps the constructor arguments / state as attributes."""

    def __init__(self, *args, **kwargs):
        self.__dict__.update(kwargs)
        self._args = args

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)


# Globals a checkpoint may legitimately reference besides the reference package: tensor storage / rebuild helpers,
# containers, numpy scalars and paths.  Anything else (os.system, builtins.eval, ...) becomes an inert _Opaque holder:
# loading a checkpoint must never execute code it names.
_ALLOWED_GLOBALS = {
    "collections": {"OrderedDict", "defaultdict"},
    "builtins": {"set", "frozenset", "dict", "list", "tuple", "int", "float", "bool", "str", "bytes", "complex", "slice",
                 "range", "bytearray", "object"},
    "pathlib": {"Path", "PosixPath", "PurePath", "PurePosixPath", "WindowsPath", "PureWindowsPath"},
    "torch": {"Size", "device", "dtype", "Tensor", "FloatStorage", "DoubleStorage", "LongStorage", "IntStorage",
              "HalfStorage", "BFloat16Storage", "BoolStorage", "ByteStorage", "CharStorage", "ShortStorage",
              "float32", "float64", "float16", "bfloat16", "int64", "int32", "int16", "int8", "uint8", "bool"},
    "torch._utils": {"_rebuild_tensor_v2", "_rebuild_tensor", "_rebuild_parameter", "_rebuild_parameter_with_state"},
    "torch.storage": {"_load_from_bytes", "TypedStorage", "UntypedStorage"},
    "torch.serialization": {"_get_layout"},
    "torch.nn.parameter": {"Parameter"},
    "numpy": {"ndarray", "dtype", "float32", "float64", "int64", "int32", "bool_"},
    "numpy.core.multiarray": {"scalar", "_reconstruct"},
    "numpy._core.multiarray": {"scalar", "_reconstruct"},
    "dataclasses": {"_HAS_DEFAULT_FACTORY_CLASS"},
}


class _TolerantUnpickler(pickle.Unpickler):
    """Resolves reference classes to this package's equivalents when they exist, allow-listed library globals to
    themselves, and everything else to an inert opaque holder."""

    def find_class(self, module, name):
        if module == REFERENCE_PACKAGE or module.startswith(REFERENCE_PACKAGE + "."):
            mirrored = module.replace(REFERENCE_PACKAGE, REFERENCE_PACKAGE + "_b200", 1)
            try:
                return super().find_class(mirrored, name)
            except (ImportError, AttributeError):
                return type(name, (_Opaque,), {})
        if name in _ALLOWED_GLOBALS.get(module, ()):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                pass
        return type(name, (_Opaque,), {})


class _TolerantPickle:
    Unpickler = _TolerantUnpickler
    __name__ = "tolerant_pickle"

    @staticmethod
    def load(file, **kwargs):
        return _TolerantUnpickler(file, **kwargs).load()


def _load_checkpoint_file(checkpoint_path: Union[str, Path]) -> Any:
    try:
        return torch.load(checkpoint_path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError:  # a Lightning checkpoint with pickled hyper-parameter objects (I/O errors propagate)
        return torch.load(checkpoint_path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)


def _find_score_network_parameters(node: Any, depth: int = 0):
    """Depth-first search of checkpoint hyper-parameters for the score-network parameter object / dict."""
    known = tuple(SCORE_NETWORK_PARAMETERS_BY_ARCH.values())
    if isinstance(node, known):
        return node
    if depth > 6:
        return None
    if isinstance(node, dict):
        if node.get("architecture") in SCORE_NETWORK_PARAMETERS_BY_ARCH and "num_atom_types" in node:
            return node
        children = list(node.values())
    elif isinstance(node, _Opaque) or dataclasses.is_dataclass(node):
        state = vars(node)
        if state.get("architecture") in SCORE_NETWORK_PARAMETERS_BY_ARCH and "num_atom_types" in state:
            return {k: v for k, v in state.items() if not k.startswith("_")}
        children = list(state.values())
    else:
        return None
    for child in children:
        found = _find_score_network_parameters(child, depth + 1)
        if found is not None:
            return found
    return None


def _parameters_from_dict(raw: Dict[str, Any]):
    data_class = SCORE_NETWORK_PARAMETERS_BY_ARCH[raw["architecture"]]
    names = {f.name for f in dataclasses.fields(data_class)}
    return data_class(**{k: v for k, v in raw.items() if k in names})


def get_axl_network(checkpoint_path: Union[str, Path], hyper_params: Optional[Dict[AnyStr, Any]] = None) -> ScoreNetwork:
    """Score network with the weights of ``checkpoint_path`` (reference :191-205).

    Accepts a Lightning checkpoint of the reference's ``AXLDiffusionLightningModel`` (``state_dict`` keys prefixed by
    ``axl_network.``, hyper-parameters pickled inside) or a plain ``state_dict`` file of the score network.
    """
    logger.info("Loading checkpoint...")
    checkpoint = _load_checkpoint_file(checkpoint_path)
    state_dict = checkpoint.get("state_dict", checkpoint) if isinstance(checkpoint, dict) else checkpoint
    state_dict = {(k[len(AXL_NETWORK_PREFIX):] if k.startswith(AXL_NETWORK_PREFIX) else k): v
                  for k, v in state_dict.items() if torch.is_tensor(v)}
    parameters = None
    if isinstance(checkpoint, dict):
        parameters = _find_score_network_parameters(checkpoint.get("hyper_parameters"))
    if parameters is None and hyper_params is not None and "model" in hyper_params:
        model_block = hyper_params["model"]
        globals_block = {k: v for k, v in hyper_params.items() if not isinstance(v, dict)}
        parameters = create_score_network_parameters(model_block["score_network"], globals_block)
    assert parameters is not None, ("The score-network architecture is neither in the checkpoint's hyper-parameters nor "
                                    "in a 'model: score_network:' block of the configuration file.")
    if isinstance(parameters, dict):
        parameters = _parameters_from_dict(parameters)
    axl_network = create_score_network(parameters)
    own_keys = set(axl_network.state_dict())
    axl_network.load_state_dict({k: v for k, v in state_dict.items() if k in own_keys}, strict=True)
    axl_network.eval()
    return axl_network


def load_and_backup_hyperparameters(config_file_path: Union[str, Path], output_directory: Union[str, Path]):
    """yaml -> dict, with a copy of the file next to the outputs (reference utils/main_utils.py:145-175)."""
    with open(config_file_path, "r") as stream:
        hyper_params = yaml.load(stream, Loader=yaml.FullLoader)
    destination = os.path.join(output_directory, "config_backup.yaml")
    if os.path.abspath(str(config_file_path)) != os.path.abspath(destination):
        shutil.copyfile(config_file_path, destination)
    return hyper_params


def extract_and_validate_parameters(hyper_params: Dict[AnyStr, Any]):
    """``(NoiseParameters, SamplingParameters)`` from the ``noise`` and ``sampling`` blocks (reference :168-188)."""
    assert "noise" in hyper_params, "The noise parameters must be defined to draw samples."
    noise_parameters = NoiseParameters(**hyper_params["noise"])
    assert "sampling" in hyper_params, "The sampling parameters must be defined to draw samples."
    sampling_parameters = load_sampling_parameters(hyper_params["sampling"])
    return noise_parameters, sampling_parameters


def create_samples_and_write_to_disk(generator, sampling_parameters: SamplingParameters, oracle_parameters,
                                     device: torch.device, output_path: Union[str, Path]):
    """Draw the samples and write ``samples.pt`` [+ ``trajectories.pt``] (reference :208-270)."""
    logger.info("Generating samples...")
    with torch.no_grad():
        samples_batch = create_batch_of_samples(generator=generator, sampling_parameters=sampling_parameters,
                                                device=device)
    logger.info("Done Generating Samples.")
    if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
        return samples_batch
    output_directory = Path(output_path)
    on_host = {CARTESIAN_POSITIONS: samples_batch[CARTESIAN_POSITIONS].cpu(),
               AXL_COMPOSITION: type(samples_batch[AXL_COMPOSITION])(*[t.cpu() for t in samples_batch[AXL_COMPOSITION]])}
    with open(output_directory / "samples.pt", "wb") as fd:
        torch.save(on_host, fd)
    axl = on_host[AXL_COMPOSITION]
    with open(output_directory / "samples_plain.pt", "wb") as fd:
        torch.save({"cartesian_positions": on_host[CARTESIAN_POSITIONS], "atom_types": axl.A,
                    "relative_coordinates": axl.X, "lattice_parameters": axl.L}, fd)
    if oracle_parameters:
        logger.warning("An energy oracle is configured; LAMMPS is outside this sampling path, energies.pt is not written.")
    if sampling_parameters.record_samples:
        logger.info("Writing sampling trajectories to disk...")
        generator.sample_trajectory_recorder.write_to_pickle(output_directory / "trajectories.pt")
    logger.info("Done!")
    return samples_batch


def main(args: Optional[Any] = None, axl_network: Optional[ScoreNetwork] = None) -> None:
    """Load a score network and draw samples (reference :52-165)."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", required=True, help="config file with sampling parameters in yaml format.")
    parser.add_argument("--checkpoint", default=None, help="path to checkpoint model to be loaded.")
    parser.add_argument("--output", required=True, help="path to outputs - will store files here")
    parser.add_argument("--path_to_starting_configuration_data_pickle", default=None,
                        help="Path to a pickle that specifies starting compositions and starting time index.")
    parser.add_argument("--path_to_sampling_constraint_data_pickle", default=None,
                        help="Path to a pickle that specifies the REPAINT sampling constraints.")
    parser.add_argument("--device", default="cuda", help="Device to use. Defaults to cuda.")
    args = parser.parse_args(args)

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl" if args.device.startswith("cuda") else "gloo")
    device = torch.device(f"cuda:{local_rank}") if args.device == "cuda" else torch.device(args.device)
    if device.type == "cuda":
        torch.cuda.set_device(device)

    os.makedirs(args.output, exist_ok=True)
    logging.basicConfig(level=logging.INFO)
    if axl_network is None:
        assert args.checkpoint is not None and os.path.exists(args.checkpoint), \
            f"The path {args.checkpoint} does not exist. Cannot go on."
    logger.info("Sampling Experiment info:")
    logger.info(f"  Hostname : {socket.gethostname()}")
    logger.info(f"  Checkpoint : {args.checkpoint}")
    logger.info(f"  Device   : {device}")

    hyper_params = load_and_backup_hyperparameters(config_file_path=args.config, output_directory=args.output)
    noise_parameters, sampling_parameters = extract_and_validate_parameters(hyper_params)
    if axl_network is None:
        axl_network = get_axl_network(args.checkpoint, hyper_params)
    axl_network = axl_network.to(device)
    if "force_field" in hyper_params:  # reference :132-139
        force_field_parameters = ForceFieldParameters(**hyper_params["force_field"])
        logger.info("Augmenting the AXL_network with an excluding Force Field.")
        axl_network = ForceFieldAugmentedScoreNetwork(axl_network, force_field_parameters)

    trajectory_initializer = instantiate_trajectory_initializer(
        sampling_parameters=sampling_parameters,
        path_to_starting_configuration_data_pickle=args.path_to_starting_configuration_data_pickle)
    sampling_constraints = None
    if args.path_to_sampling_constraint_data_pickle is not None:
        sampling_constraints = read_sampling_constraint(args.path_to_sampling_constraint_data_pickle)
    generator = instantiate_generator(sampling_parameters=sampling_parameters, noise_parameters=noise_parameters,
                                      axl_network=axl_network, trajectory_initializer=trajectory_initializer,
                                      sampling_constraints=sampling_constraints)
    create_samples_and_write_to_disk(generator=generator, sampling_parameters=sampling_parameters,
                                     oracle_parameters=hyper_params.get("oracle"), device=device,
                                     output_path=args.output)
    if world_size > 1 and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

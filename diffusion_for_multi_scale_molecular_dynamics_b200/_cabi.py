"""ctypes binding of ``csrc/libb200md.so`` (C ABI: ``include/b200md.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/build.py``.  Loading is lazy so that
the host-side modules (noise tables, dataclasses, sharding logic) import on a machine without the
extension; every *compute* entry point goes through :func:`lib`, which raises :class:`B200MDError` when
the shared object is missing -- there is deliberately no fallback implementation.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint32, c_uint64, c_void_p
from typing import Optional

import torch

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIBRARY_PATH = os.path.join(_CSRC, "libb200md.so")


class B200MDError(RuntimeError):
    """Raised when the sm_100a extension is unavailable or a C-ABI call fails."""


# status bits (include/b200md.h)
STATUS_X_OUT_OF_RANGE = 1
STATUS_A_OUT_OF_RANGE = 2
STATUS_CUTOFF_TOO_LARGE = 4
STATUS_EDGE_OVERFLOW = 8
STATUS_MASK_REMAINS = 16
STATUS_BAD_LOGITS = 32
STATUS_SIGMA_OUT_OF_RANGE = 64
STATUS_TC_RANGE = 128

STATUS_MESSAGES = {
    STATUS_X_OUT_OF_RANGE: "All components of the relative coordinates are expected to be in [0,1).",
    STATUS_A_OUT_OF_RANGE: "All atom types are expected to be in [0, num_atom_types].",
    STATUS_CUTOFF_TOO_LARGE: "The radial cutoff is so large that neighbors could be located beyond the first "
    "shell of periodic unit cell images.",
    STATUS_EDGE_OVERFLOW: "The edge workspace is too small for this batch; raise edge_capacity.",
    STATUS_MASK_REMAINS: "There remains MASKED atoms at the last time step: review code, there must be a bug or "
    "invalid input.",
    STATUS_BAD_LOGITS: "Logits are pathological: the probabilities do not sum to one.",
    STATUS_SIGMA_OUT_OF_RANGE: "All values of sigma should be larger than zero.",
    STATUS_TC_RANGE: "An activation exceeded the fp16 range of the 3xFP16 tensor-core kernels; set "
    "kernel_impl = IMPL_TCGEN05 (3xTF32) for this network (kernel_impl = IMPL_AUTO does so by itself).",
}

DRAW_COORD_NORMAL, DRAW_GUMBEL, DRAW_BINARY, DRAW_LATTICE_NORMAL = 0, 1, 2, 3
DRAW_INIT_UNIFORM, DRAW_REPAINT_NORMAL, DRAW_REPAINT_UNIFORM, DRAW_REPAINT_INIT = 4, 5, 6, 7
DRAW_LATTICE_NORMAL_SECOND = 8  # the corrector's second lattice draw (reference langevin_generator.py:771-777)
DRAW_INIT_LATTICE = 9  # the starting lattice L_T: its own key, (step 0, DRAW_LATTICE_NORMAL) is the last predictor's draw

ATOM_GREEDY, ATOM_ONE_TRANSITION, ATOM_LAST_STEP = 1, 2, 4
GRAPH_UNIQUE_PAIRS, GRAPH_PER_IMAGE = 0, 1
IMPL_AUTO, IMPL_FFMA, IMPL_TCGEN05, IMPL_TCGEN05_F16X3, IMPL_TCGEN05_F16X3_TS = 0, 1, 2, 3, 4

EGNN_RESIDUAL, EGNN_ATTENTION, EGNN_NORMALIZE, EGNN_TANH = 1, 2, 4, 8
EGNN_COORDS_MEAN, EGNN_MESSAGE_MEAN, EGNN_RADIAL_CUTOFF, EGNN_KEEP_DUPLICATES = 16, 32, 64, 128


class EgnnConfig(ctypes.Structure):
    """Mirror of ``b200md_egnn_config``."""

    _fields_ = [
        ("spatial_dimension", c_int32),
        ("num_atom_types", c_int32),
        ("n_layers", c_int32),
        ("n_bloch", c_int32),
        ("node_hidden", c_int32),
        ("node_n_hidden", c_int32),
        ("message_hidden", c_int32),
        ("message_n_hidden", c_int32),
        ("coord_hidden", c_int32),
        ("coord_n_hidden", c_int32),
        ("flags", c_int32),
        ("radial_cutoff", c_float),
        ("lattice_clip_min", c_float),
    ]


class PcSchedule(ctypes.Structure):
    """Mirror of ``b200md_pc_schedule``."""

    _fields_ = [
        ("total_time_steps", c_int32),
        ("number_of_corrector_steps", c_int32),
        ("start_step_index", c_int32),
        ("end_step_index", c_int32),
        ("atom_flags", c_int32),
        ("atom_type_transition_in_corrector", c_int32),
        ("small_epsilon", c_float),
        ("sigma_min", c_float),
        ("sigma_host", c_void_p),
        ("g_host", c_void_p),
        ("g2_host", c_void_p),
        ("epsilon_host", c_void_p),
        ("sqrt_2_epsilon_host", c_void_p),
        ("q", c_void_p),
        ("q_bar", c_void_p),
        ("q_bar_tm1", c_void_p),
        ("n_constrained", c_int32),
        ("x0_known", c_void_p),
        ("a0_known", c_void_p),
        ("constrained_indices", c_void_p),
    ]


_P = c_void_p
_CFG = POINTER(EgnnConfig)

# name -> (restype, argtypes); must list EVERY symbol declared in include/b200md.h
SIGNATURES = {
    "b200md_error_string": (c_char_p, [c_int32]),
    "b200md_abi_version": (c_int32, []),
    "b200md_launch_count": (c_uint64, []),
    "b200md_profile_enable": (c_int32, [c_int32]),
    "b200md_profile_collect": (c_int32, [POINTER(ctypes.c_double), POINTER(c_int64), c_int32]),
    "b200md_draw_normal": (c_int32, [_P, c_int64, c_int32, c_int32, c_uint64, c_uint32, c_uint32, c_uint64, _P]),
    "b200md_draw_uniform": (c_int32, [_P, c_int64, c_int32, c_int32, c_uint64, c_uint32, c_uint32, c_uint64, _P]),
    "b200md_draw_gumbel": (c_int32, [_P, c_int64, c_int32, c_int32, c_float, c_uint64, c_uint32, c_uint32, c_uint64, _P]),
    "b200md_wrap_to_unit_cell": (c_int32, [_P, _P, c_int64, _P]),
    "b200md_pc_coordinates_update": (
        c_int32,
        [_P, _P, _P, c_float, c_float, c_float, c_int64, c_int32, c_int32, c_uint64, c_uint32, c_uint64, _P, _P],
    ),
    "b200md_pc_lattice_update": (
        c_int32,
        [_P, _P, _P, c_float, c_float, c_float, c_int64, c_int32, c_uint64, c_uint32, c_uint64, _P, _P],
    ),
    "b200md_adaptive_corrector_step_size": (
        c_int32, [_P, _P, c_int64, c_int32, c_int32, c_float, c_float, c_float, _P, _P, _P]),
    "b200md_pc_update_device_weights": (c_int32, [_P, _P, _P, _P, c_float, c_int64, c_int32, _P, _P]),
    "b200md_atom_types_update": (
        c_int32,
        [_P, _P, _P, _P, _P, _P, _P, c_int32, c_float, c_int64, c_int32, c_int32, c_uint64, c_uint32, c_uint64,
         _P, _P, _P, _P],
    ),
    "b200md_repaint": (
        c_int32,
        [_P, _P, _P, _P, _P, c_int32, c_float, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_uint64, c_uint32,
         c_uint64, _P],
    ),
    "b200md_wrapped_gaussian_score_f32": (c_int32, [_P, _P, c_int64, c_int32, _P, _P]),
    "b200md_wrapped_gaussian_score_f64": (c_int32, [_P, _P, c_int64, c_int32, _P, _P]),
    "b200md_analytical_score_f32": (c_int32, [_P, _P, _P, c_float, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "b200md_segment_reduce": (c_int32, [_P, _P, c_int64, c_int32, c_int64, c_int32, _P, _P, _P]),
    "b200md_radius_graph_count": (c_int32, [_P, _P, c_int32, c_float, c_int64, c_int32, c_int32, _P, _P, _P, _P]),
    "b200md_radius_graph_fill": (
        c_int32, [_P, _P, c_int32, c_float, c_int64, c_int32, c_int32, _P, c_int64, _P, _P, _P, _P, _P]),
    "b200md_radius_graph_count_masked": (c_int32, [_P, _P, c_int32, c_float, c_int64, c_int32, _P, _P, _P, _P, _P]),
    "b200md_radius_graph_fill_from_mask": (c_int32, [_P, c_int64, c_int32, _P, c_int64, _P, _P, _P, _P]),
    "b200md_pair_distances_count": (c_int32, [_P, _P, c_int32, c_float, c_int64, c_int32, _P, _P, _P]),
    "b200md_pair_distances_fill": (c_int32, [_P, _P, c_int32, c_float, c_int64, c_int32, _P, _P, _P]),
    "b200md_repulsive_pseudo_force": (c_int32, [_P, _P, c_float, c_float, c_float, c_int64, c_int32, _P, _P, _P, _P]),
    "b200md_stillinger_weber_energy": (
        c_int32, [_P, _P, _P, _P, c_int32, c_float, c_float, c_int64, c_int32, _P, _P, _P]),
    "b200md_egnn_param_count": (c_int32, [_CFG]),
    "b200md_egnn_packed_floats": (c_int64, [_CFG]),
    "b200md_egnn_pack_weights": (c_int32, [_CFG, POINTER(c_void_p), c_int32, _P, _P]),
    "b200md_egnn_default_edge_capacity": (c_int64, [_CFG, c_int64, c_int32, POINTER(c_float)]),
    "b200md_egnn_workspace_bytes": (c_int64, [_CFG, c_int64, c_int32, c_int64]),
    "b200md_egnn_resolve_impl": (c_int32, [_CFG, c_int32]),
    "b200md_egnn_forward": (
        c_int32,
        [_CFG, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P, _P, _P, c_int64, c_int64, _P, _P, c_int32, _P],
    ),
    "b200md_pc_sample": (
        c_int32,
        [_CFG, _P, POINTER(PcSchedule), _P, _P, _P, c_int64, c_int32, c_uint64, c_uint64, _P, c_int64, c_int64, _P,
         c_int32, _P],
    ),
    "b200md_pc_sample_workspace_bytes": (c_int64, [_CFG, c_int64, c_int32, c_int64]),
    "b200md_pc_sample_last_mode": (c_int32, []),
}

_LIB: Optional[ctypes.CDLL] = None


def load_library(path: str = LIBRARY_PATH) -> ctypes.CDLL:
    """dlopen the extension and bind every symbol of include/b200md.h (no GPU needed)."""
    if not os.path.isfile(path):
        raise B200MDError(
            f"sm_100a extension not found at {path}. Build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (nvcc, -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback."
        )
    library = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(library, name)  # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    return library


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load_library()
        if _LIB.b200md_abi_version() != 1:
            raise B200MDError("libb200md.so ABI version mismatch; rebuild the extension.")
    return _LIB


def check(code: int) -> None:
    """Turn a negative C-ABI return code into an exception."""
    if code < 0:
        message = lib().b200md_error_string(int(code)).decode()
        raise B200MDError(f"b200md call failed ({code}): {message}")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors must live on the same CUDA device; anything else is a hard error (no CPU path)."""
    device = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise B200MDError(
                "this operation runs only on a CUDA device (sm_100a kernels); got a tensor on "
                f"'{t.device}'. There is no CPU fallback in diffusion_for_multi_scale_molecular_dynamics_b200."
            )
        if device is None:
            device = t.device
        elif t.device != device:
            raise B200MDError("all tensors must be on the same CUDA device")
    return device


def ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def stream_ptr(device: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def as_f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def as_i64(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.int64).contiguous()


def status_word(status: Optional[torch.Tensor]) -> int:
    """The device status word as a host integer (one host sync)."""
    return 0 if status is None else int(status.item())


def raise_for_word(word: int, status: Optional[torch.Tensor], error=AssertionError) -> None:
    """Turn a non-zero status word into the reference's assertion text; the device word is cleared first."""
    if word:
        messages = [text for bit, text in STATUS_MESSAGES.items() if word & bit]
        if status is not None:
            status.zero_()
        raise error(" ".join(messages) if messages else f"device status {word}")


def raise_on_status(status: torch.Tensor, error=AssertionError) -> None:
    """Read the device status word (one host sync) and raise like the reference's asserts would."""
    raise_for_word(status_word(status), status, error)


KERNEL_CLASSES = ("graph", "node", "edge", "update", "other")


def profile_enable(on: bool) -> None:
    check(lib().b200md_profile_enable(int(on)))


def profile_collect() -> dict:
    """{class: (milliseconds, launches)} accumulated since the last collect (synchronises the recorded events)."""
    n = len(KERNEL_CLASSES)
    ms = (ctypes.c_double * n)()
    launches = (c_int64 * n)()
    check(lib().b200md_profile_collect(ms, launches, n))
    return {name: (float(ms[i]), int(launches[i])) for i, name in enumerate(KERNEL_CLASSES)}


def launch_count() -> int:
    return int(lib().b200md_launch_count())

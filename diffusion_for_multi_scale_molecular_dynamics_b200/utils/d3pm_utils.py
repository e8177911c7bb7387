"""Discrete-diffusion helpers (reference: utils/d3pm_utils.py:7-150).

On the sampling path the posterior p(a_{t-1} | a_t) is evaluated inside the fused atom-type kernel
(``b200md_atom_types_update``); :func:`get_probability_at_previous_time_step` exposes that evaluation.
"""
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi


def class_index_to_onehot(index: torch.Tensor, num_classes: int) -> torch.Tensor:
    """float one-hot encoding with a trailing class dimension (reference :7-20)."""
    return torch.nn.functional.one_hot(index.long(), num_classes=num_classes).to(device=index.device, dtype=torch.float)


def _single_matrix(m: torch.Tensor) -> torch.Tensor:
    """The sampler broadcasts one [C, C] matrix to [B, N, C, C]; the kernel wants the [C, C] original."""
    while m.dim() > 2:
        m = m[0]
    return _cabi.as_f32(m)


def get_probability_at_previous_time_step(
    probability_at_zeroth_timestep: torch.Tensor,
    one_hot_probability_at_current_timestep: torch.Tensor,
    q_matrices: torch.Tensor,
    q_bar_matrices: torch.Tensor,
    q_bar_tm1_matrices: torch.Tensor,
    small_epsilon: float,
    probability_at_zeroth_timestep_are_logits: bool = False,
) -> torch.Tensor:
    """p(a_{t-1} | a_t, gamma_0) for a common time index (reference :64-124), logits input only."""
    if not probability_at_zeroth_timestep_are_logits:
        raise _cabi.B200MDError("the sm_100a atom-type kernel takes logits (the sampling path's call signature)")
    logits = _cabi.as_f32(probability_at_zeroth_timestep)
    device = _cabi.require_cuda(logits, one_hot_probability_at_current_timestep)
    batch, natoms, num_classes = logits.shape
    atom_types = one_hot_probability_at_current_timestep.argmax(dim=-1).to(torch.int64).contiguous()
    out_types = torch.empty_like(atom_types)
    probs = torch.empty_like(logits)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    zeros = torch.zeros(batch, natoms, num_classes, dtype=torch.float32, device=device)
    q, qb, qbm = (_single_matrix(m).to(device) for m in (q_matrices, q_bar_matrices, q_bar_tm1_matrices))
    _cabi.check(_cabi.lib().b200md_atom_types_update(
        _cabi.ptr(logits), _cabi.ptr(atom_types), _cabi.ptr(q), _cabi.ptr(qb), _cabi.ptr(qbm), _cabi.ptr(zeros), None,
        0, float(small_epsilon), batch, natoms, num_classes, 0, 0, 0, _cabi.ptr(out_types), _cabi.ptr(probs),
        _cabi.ptr(status), _cabi.stream_ptr(device)))
    _cabi.raise_on_status(status, AssertionError)
    return probs

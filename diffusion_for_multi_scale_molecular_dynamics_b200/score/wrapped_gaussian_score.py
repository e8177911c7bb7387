"""Sigma-normalised score of the wrapped Gaussian kernel (reference: score/wrapped_gaussian_score.py).

One kernel (``b200md_wrapped_gaussian_score_f32/_f64``) evaluates the lattice-image sum k in [-kmax, kmax]
with the reference's three numerically stable branches, selected per element in registers instead of by
boolean-mask gather/scatter.
"""
import numpy as np
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi

SIGMA_THRESHOLD = torch.Tensor([1.0 / np.sqrt(2.0 * np.pi)])
U_THRESHOLD = torch.Tensor([0.5])


def get_coordinates_sigma_normalized_score(
    relative_coordinates: torch.Tensor, sigmas: torch.Tensor, kmax: int, coordinates_bounded: bool = True
) -> torch.Tensor:
    """sigma * d/du log K_sigma(u) for the wrapped Gaussian K (reference :131-198)."""
    assert kmax >= 0, "kmax must be a non negative integer"
    assert sigmas.shape == relative_coordinates.shape, \
        "The relative_coordinates and sigmas inputs should have the same shape"
    assert sigmas.device == relative_coordinates.device, "relative_coordinates and sigmas should be on the same device."
    device = _cabi.require_cuda(relative_coordinates, sigmas)
    # the reference's two value asserts (:170-178) as ONE host synchronisation: a 2-bit code reduced on the device
    code = (sigmas <= 0).any().to(torch.int32)
    if coordinates_bounded:
        outside = torch.logical_or(relative_coordinates < 0, relative_coordinates >= 1).any().to(torch.int32)
        code = code + 2 * outside
    code = int(code.item())
    assert not code & 1, "All values of sigma should be larger than zero."
    assert not code & 2, "the relative coordinates should all be in [0, 1)"
    dtype = relative_coordinates.dtype
    if dtype not in (torch.float32, torch.float64):
        raise _cabi.B200MDError("wrapped Gaussian score supports float32 and float64")
    u = relative_coordinates.detach().contiguous()
    s = sigmas.detach().to(dtype).contiguous()
    out = torch.empty_like(u)
    fn = _cabi.lib().b200md_wrapped_gaussian_score_f32 if dtype == torch.float32 else \
        _cabi.lib().b200md_wrapped_gaussian_score_f64
    _cabi.check(fn(_cabi.ptr(u), _cabi.ptr(s), u.numel(), int(kmax), _cabi.ptr(out), _cabi.stream_ptr(device)))
    return out

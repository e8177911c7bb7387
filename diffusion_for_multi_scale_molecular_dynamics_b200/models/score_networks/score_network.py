"""ScoreNetwork base class (reference: models/score_networks/score_network.py:26-242).

Same contract as the reference: ``forward(batch, conditional)`` validates the batch, calls
``_forward_unchecked`` and forces the MASK logit to -inf.  The reference validates tensor *values* with
``assert tensor.all()`` (one host synchronisation each, >= 6 per call); here shapes are checked on the host and
values are checked by the kernels, which raise bits in a device-side status word.  The word is read once per
``forward`` (or once per ``sample()`` when a generator sets ``defer_status_check``) and turned into the same
``AssertionError``.
"""
from dataclasses import dataclass
from typing import AnyStr, Dict, Optional

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, CARTESIAN_FORCES, NOISE,
                                                                       NOISY_AXL_COMPOSITION, TIME)
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import \
    get_number_of_lattice_parameters


@dataclass(kw_only=True)
class ScoreNetworkParameters:
    """Base hyper-parameters of a score network (reference :26-45)."""

    architecture: str
    spatial_dimension: int = 3
    num_atom_types: int  # atomic species, not counting the MASK class
    conditional_prob: float = 0.0
    conditional_gamma: float = 2.0

    def __post_init__(self):
        self.num_lattice_parameters = get_number_of_lattice_parameters(self.spatial_dimension)


class ScoreNetwork(torch.nn.Module):
    """Polymorphic interface of all score networks."""

    def __init__(self, hyper_params: ScoreNetworkParameters):
        super().__init__()
        self._hyper_params = hyper_params
        self.spatial_dimension = hyper_params.spatial_dimension
        self.num_atom_types = hyper_params.num_atom_types
        self.conditional_prob = hyper_params.conditional_prob
        self.conditional_gamma = hyper_params.conditional_gamma
        # When True the caller (a generator) takes responsibility for reading `device_status()` later.
        self.defer_status_check = False
        self._status: Optional[torch.Tensor] = None

    # ---- device-side status word ---------------------------------------------------------------------
    def device_status(self, device: torch.device) -> torch.Tensor:
        if self._status is None or self._status.device != device:
            self._status = torch.zeros(1, dtype=torch.int32, device=device)
        return self._status

    def raise_on_device_status(self) -> None:
        if self._status is not None:
            _cabi.raise_on_status(self._status, AssertionError)

    # ---- validation --------------------------------------------------------------------------------------
    def _check_batch(self, batch: Dict[AnyStr, torch.Tensor]):
        """Host-side structural checks (reference :68-181); value checks are done on the device."""
        assert NOISY_AXL_COMPOSITION in batch, (
            f"The noisy coordinates, atomic types and lattice vectors should be present in the batch dictionary "
            f"with key '{NOISY_AXL_COMPOSITION}'")
        composition = batch[NOISY_AXL_COMPOSITION]
        x_shape = composition.X.shape
        assert len(x_shape) == 3 and x_shape[2] == self.spatial_dimension, (
            "The relative coordinates are expected to be in a tensor of shape [batch_size, number_of_atoms, "
            "spatial_dimension]")
        batch_size = x_shape[0]
        assert TIME in batch, f"The time step should be present in the batch dictionary with key '{TIME}'"
        t_shape = batch[TIME].shape
        assert t_shape[0] == batch_size, "the batch size dimension is inconsistent between positions and time steps."
        assert len(t_shape) == 2 and t_shape[1] == 1, \
            "The time steps are expected to be in a tensor of shape [batch_size, 1]"
        assert NOISE in batch, "There should be a 'noise' parameter in the batch dictionary."
        assert batch[NOISE].shape == t_shape, "the 'noise' parameter should have the same shape as the 'time'."
        l_shape = composition.L.shape
        assert l_shape[0] == batch_size, "the batch size dimension is inconsistent between positions and unit cell."
        assert len(l_shape) == 2 and l_shape[1] == int(self.spatial_dimension * (self.spatial_dimension + 1) / 2), (
            "The lattice parameters are expected to be in a tensor of shape [batch_size, spatial_dimension * "
            "(spatial_dimension + 1) / 2].}")
        a_shape = composition.A.shape
        assert a_shape[0] == batch_size, "the batch size dimension is inconsistent between positions and atom types."
        assert len(a_shape) == 2, "The atoms type are expected to be in a tensor of shape [batch_size, number of atoms]."
        if self.conditional_prob > 0:
            assert CARTESIAN_FORCES in batch, (
                f"The cartesian forces should be present in the batch dictionary with key '{CARTESIAN_FORCES}'")
            f_shape = batch[CARTESIAN_FORCES].shape
            assert len(f_shape) == 3 and f_shape[2] == self.spatial_dimension, (
                "The cartesian forces are expected to be in a tensor of shape [batch_size, number_of_atoms,"
                f"{self.spatial_dimension}]")
        self._check_batch_values(batch)

    def _check_batch_values(self, batch: Dict[AnyStr, torch.Tensor]):
        """Value checks for subclasses whose kernels do not validate their inputs (reference :110-165)."""
        composition = batch[NOISY_AXL_COMPOSITION]
        x, a, t = composition.X, composition.A, batch[TIME]
        assert torch.logical_and(x >= 0.0, x < 1.0).all(), \
            "All components of the relative coordinates are expected to be in [0,1)."
        assert torch.logical_and(t >= 0.0, t <= 1.0).all(), "The times are expected to be normalized between 0 and 1."
        assert torch.logical_and(a >= 0, a < self.num_atom_types + 1).all(), \
            f"All atom types are expected to be in [0, {self.num_atom_types}]."

    def _impose_non_mask_atomic_type_prediction(self, output: AXL):
        """The model may never predict MASK: last logit -> -inf (reference :183-185)."""
        output.A[..., self.num_atom_types] = -torch.inf

    def forward(self, batch: Dict[AnyStr, torch.Tensor], conditional: Optional[bool] = None) -> AXL:
        """Validate, evaluate, forbid MASK (reference :187-223)."""
        self._check_batch(batch)
        if conditional is None:
            conditional = bool(torch.rand(1,) < self.conditional_prob)
        if conditional:
            # The reference's conditional branch multiplies AXL namedtuples by floats and cannot run (:213-219).
            raise NotImplementedError("conditional score evaluation is not available on the sampling path")
        output = self._forward_unchecked(batch, conditional=False)
        if not self.defer_status_check:
            word = _cabi.status_word(self._status)  # the one host synchronisation of this call
            if word and self.widen_range_after_overflow(word):
                output = self._forward_unchecked(batch, conditional=False)  # once more, with the wider-range kernels
                word = _cabi.status_word(self._status)
            _cabi.raise_for_word(word, self._status, AssertionError)
        self._impose_non_mask_atomic_type_prediction(output)
        return output

    def widen_range_after_overflow(self, word: Optional[int] = None) -> bool:
        """True when the last evaluation overflowed a reduced-range kernel AND the network has now switched to a
        wider-range implementation, i.e. the caller should evaluate again (see EGNNScoreNetwork).  ``word`` is the
        status word if the caller has read it already."""
        return False

    def _forward_unchecked(self, batch: Dict[AnyStr, torch.Tensor], conditional: bool = False) -> AXL:
        raise NotImplementedError

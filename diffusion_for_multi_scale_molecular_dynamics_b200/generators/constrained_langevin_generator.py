"""RePaint-style constrained sampling on sm_100a (reference: generators/constrained_langevin_generator.py:24-182).

After every predictor step the known atoms are re-noised from their t=0 values to the current time index and
written over the sampled composition (``b200md_repaint``: x <- wrap(x0 + sigma_i z), a <- argmax(log(onehot(a0)
Qbar_i) + gumbel)); after the trajectory the constraint is applied once more without noise.
"""
from typing import Optional

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.sampling_constraint import SamplingConstraint
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import TrajectoryInitializer
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import ScoreNetwork
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters


class ConstrainedLangevinGenerator(LangevinGenerator):
    """Annealed Langevin dynamics with inpainting of a fixed set of atoms."""

    _HOOKS = LangevinGenerator._HOOKS + ("_apply_constraint", "_repaint_composition", "_draw_repaint_gaussian_sample",
                                         "_draw_repaint_uniform_sample")

    def __init__(
        self,
        noise_parameters: NoiseParameters,
        sampling_parameters: PredictorCorrectorSamplingParameters,
        axl_network: ScoreNetwork,
        sampling_constraints: SamplingConstraint,
        trajectory_initializer: Optional[TrajectoryInitializer] = None,
    ):
        super().__init__(noise_parameters=noise_parameters, sampling_parameters=sampling_parameters,
                         axl_network=axl_network, trajectory_initializer=trajectory_initializer)
        self.sampling_constraints = sampling_constraints
        number_of_constraints, spatial_dimension = sampling_constraints.constrained_relative_coordinates.shape
        assert len(sampling_constraints.elements) == sampling_parameters.num_atom_types, \
            "Inconsistent number of atom types vs. elements list"
        assert number_of_constraints <= self.number_of_atoms, "There are more constrained positions than atoms!"
        assert spatial_dimension <= self.spatial_dimension, \
            "The spatial dimension of the constrained relative coordinates is inconsistent"
        if sampling_constraints.constrained_indices is None:
            self.constraint_indices = torch.arange(number_of_constraints)  # the first atoms are the known ones
        else:
            self.constraint_indices = sampling_constraints.constrained_indices
        self._constraint_cache = {}

    @classmethod
    def _fused_classes(cls):
        return (LangevinGenerator, ConstrainedLangevinGenerator)

    def _constraint_on(self, device: torch.device):
        key = str(device)
        if key not in self._constraint_cache:
            c = self.sampling_constraints
            self._constraint_cache[key] = (
                c.constrained_relative_coordinates.to(device=device, dtype=torch.float32).contiguous(),
                c.constrained_atom_types.to(device=device, dtype=torch.int64).contiguous(),
                self.constraint_indices.to(device=device, dtype=torch.int64).contiguous())
        return self._constraint_cache[key]

    def _constraint_arguments(self, device: torch.device):
        x0, a0, idx = self._constraint_on(device)
        return int(idx.shape[0]), x0, a0, idx

    # noise hooks of the re-noising step, shapes [B, Nc, d] and [B, Nc, C]
    def _draw_repaint_gaussian_sample(self, number_of_samples):
        return None  # None -> drawn inside b200md_repaint (draw id REPAINT_NORMAL)

    def _draw_repaint_uniform_sample(self, number_of_samples):
        return None  # None -> drawn inside b200md_repaint (draw id REPAINT_UNIFORM)

    def _repaint_call(self, composition: AXL, sigma_i: float, q_bar_i: Optional[torch.Tensor]) -> AXL:
        device = _cabi.require_cuda(composition.X, composition.A)
        x0, a0, idx = self._constraint_on(device)
        x, a = composition.X, composition.A
        if x.dtype != torch.float32 or not x.is_contiguous() or a.dtype != torch.int64 or not a.is_contiguous():
            raise _cabi.B200MDError("repaint works in place on contiguous float32 X and int64 A tensors")
        number_of_samples, number_of_atoms, d = x.shape
        z = u = None
        if sigma_i >= 0.0:
            z = self._draw_repaint_gaussian_sample(number_of_samples)
            u = self._draw_repaint_uniform_sample(number_of_samples)
            z = None if z is None else _cabi.as_f32(z.to(device))
            u = None if u is None else _cabi.as_f32(u.to(device))
        _cabi.check(_cabi.lib().b200md_repaint(
            _cabi.ptr(x), _cabi.ptr(a), _cabi.ptr(x0), _cabi.ptr(a0), _cabi.ptr(idx), int(idx.shape[0]), float(sigma_i),
            _cabi.ptr(q_bar_i), _cabi.ptr(z), _cabi.ptr(u), number_of_samples, number_of_atoms, d, self.num_classes,
            self._active_seed, self._step_id, int(self.sample_offset), _cabi.stream_ptr(device)))
        return AXL(A=a, X=x, L=composition.L)

    def _apply_constraint(self, composition: AXL, device: torch.device) -> AXL:
        """Overwrite the constrained atoms with their known values, in place (reference :74-82)."""
        return self._repaint_call(composition, -1.0, None)

    def _repaint_composition(self, raw_composition_i: AXL, index_i: int) -> AXL:
        """Known atoms noised from t=0 to time index ``index_i`` replace the sampled ones, in place (:136-163)."""
        if index_i == 0:  # already at t=0: no noise (reference :122-124)
            return self._repaint_call(raw_composition_i, -1.0, None)
        idx = index_i - 1
        device = raw_composition_i.X.device
        q_bar = self._tables_on(device)[1][idx]
        return self._repaint_call(raw_composition_i, float(self._host["sigma"][idx]), q_bar)

    def predictor_step(self, composition_i: AXL, index_i: int, cartesian_forces: torch.Tensor) -> AXL:
        raw = super().predictor_step(composition_i=composition_i, index_i=index_i, cartesian_forces=cartesian_forces)
        return self._repaint_composition(raw_composition_i=raw, index_i=index_i - 1)

    def sample(self, number_of_samples: int, device: torch.device) -> AXL:
        composition = super().sample(number_of_samples=number_of_samples, device=device)
        return self._apply_constraint(composition, device)  # one last time, without noise

"""Annealed-Langevin predictor-corrector generator on sm_100a (reference: generators/langevin_generator.py:27-831).

Same constructor, public methods and overridable hooks as the reference.  Differences in mechanism only:

* every arithmetic step runs in a CUDA kernel behind the C ABI (``include/b200md.h``): the coordinate update
  fused with the unit-cell wrap, the D3PM atom-type update fused with softmax / posterior / greedy / Gumbel
  arg-max / one-transition, the lattice update;
* noise is drawn on the device by a counter-based generator keyed by (seed, GLOBAL sample index, atom, step,
  draw kind) instead of on the CPU with a per-step host-to-device copy -- the ``_draw_*`` hooks return device
  tensors and can be patched exactly like the reference's;
* ``sample()`` / ``sample_from_noisy_composition()`` hand the whole trajectory to ``b200md_pc_sample`` when the
  score network is the EGNN and no hook is overridden; otherwise they run the reference's step-by-step loop
  over ``predictor_step`` / ``corrector_step``;
* the reference's data-dependent asserts (host synchronisations) are recorded in a device status word that is
  read once per trajectory (or once per step when the step API is called directly).
"""
import ctypes
import dataclasses
from typing import Optional, Tuple

import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import (
    PredictorCorrectorAXLGenerator, PredictorCorrectorSamplingParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import TrajectoryInitializer
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.score_network import ScoreNetwork
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, CARTESIAN_FORCES, NOISE,
                                                                       NOISY_AXL_COMPOSITION, TIME)
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_scheduler import NoiseScheduler
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.sample_trajectory import SampleTrajectory


def _as_float(value) -> float:
    return float(value.item()) if isinstance(value, torch.Tensor) else float(value)


def _on_device(value) -> bool:
    return torch.is_tensor(value) and value.is_cuda


def _device_weights(score_weight, noise_weight, device: torch.device) -> torch.Tensor:
    """(score weight, noise weight) as a 2-float device tensor, without a host round trip for device inputs."""
    parts = [torch.as_tensor(w, dtype=torch.float32, device=device).reshape(-1)[:1] for w in (score_weight, noise_weight)]
    return torch.cat(parts).contiguous()


class LangevinGenerator(PredictorCorrectorAXLGenerator):
    """Annealed Langevin dynamics (Song & Ermon 2019) with D3PM atom types, as in the reference."""

    # hooks whose override forces the step-by-step loop
    _HOOKS = ("_draw_coordinates_gaussian_sample", "_draw_lattice_gaussian_sample", "_draw_gumbel_sample",
              "_draw_binary_sample", "_get_model_predictions", "_relative_coordinates_update",
              "_relative_coordinates_update_predictor_step", "_relative_coordinates_update_corrector_step",
              "_atom_types_update", "_lattice_parameters_update",
              "_lattice_parameters_update_predictor_step", "_lattice_parameters_update_corrector_step",
              "_get_coordinates_corrector_step_size", "_get_lattice_parameters_corrector_step_size",
              "predictor_step", "corrector_step")

    def __init__(
        self,
        noise_parameters: NoiseParameters,
        sampling_parameters: PredictorCorrectorSamplingParameters,
        axl_network: ScoreNetwork,
        trajectory_initializer: Optional[TrajectoryInitializer] = None,
    ):
        super().__init__(
            number_of_discretization_steps=noise_parameters.total_time_steps,
            number_of_corrector_steps=sampling_parameters.number_of_corrector_steps,
            spatial_dimension=sampling_parameters.spatial_dimension,
            num_atom_types=sampling_parameters.num_atom_types,
            number_of_atoms=sampling_parameters.number_of_atoms,
            use_fixed_lattice_parameters=sampling_parameters.use_fixed_lattice_parameters,
            fixed_lattice_parameters=sampling_parameters.fixed_lattice_parameters,
            trajectory_initializer=trajectory_initializer,
        )
        self.noise_parameters = noise_parameters
        sampler = NoiseScheduler(noise_parameters, num_classes=self.num_classes)
        self.noise, self.langevin_dynamics = sampler.get_all_sampling_parameters()
        self.number_of_atoms = sampling_parameters.number_of_atoms
        self.masked_atom_type_index = self.num_classes - 1
        self.axl_network = axl_network
        self.small_epsilon = sampling_parameters.small_epsilon
        self.one_atom_type_transition_per_step = sampling_parameters.one_atom_type_transition_per_step
        self.atom_type_greedy_sampling = sampling_parameters.atom_type_greedy_sampling
        self.atom_type_transition_in_corrector = sampling_parameters.atom_type_transition_in_corrector
        self.use_fixed_lattice_parameters = sampling_parameters.use_fixed_lattice_parameters
        self.fixed_lattice_parameters = sampling_parameters.fixed_lattice_parameters
        self.record = sampling_parameters.record_samples
        self.record_corrector = sampling_parameters.record_samples_corrector_steps
        self.record_atom_type_update = sampling_parameters.record_atom_type_update
        if self.record_corrector or self.record_atom_type_update:
            assert self.record, "Corrector steps or atom_type_update can only be recorded if record_samples is True."
        if self.record:
            self.sample_trajectory_recorder = SampleTrajectory()
            self.sample_trajectory_recorder.record(key="noise", entry=self.noise)
            self.sample_trajectory_recorder.record(key="noise_parameters", entry=dataclasses.asdict(noise_parameters))
            self.sample_trajectory_recorder.record(key="sampling_parameters",
                                                   entry=dataclasses.asdict(sampling_parameters))

        # ---- sm_100a state ----
        # Host copies of the per-step scalars (kernel launch arguments) and device copies of the [T,C,C] tables.
        self._host = {name: getattr(self.noise, name).detach().to(torch.float32).cpu().contiguous()
                      for name in ("time", "sigma", "g", "g_squared")}
        self._host["epsilon"] = self.langevin_dynamics.epsilon.detach().to(torch.float32).cpu().contiguous()
        # sqrt(2 * eps_i) recomputed per step like the reference does (:749), not read from the stored table
        self._host["sqrt_2_epsilon"] = torch.sqrt(2 * self._host["epsilon"]).contiguous()
        self._device_tables = {}
        self.seed: Optional[int] = None  # None -> drawn from torch's global CPU generator at each sample()
        self.sample_offset = 0  # global index of this generator's first sample (multi-GPU sharding)
        self.allow_fused = True
        self._active_seed = 0
        self._step_id = 0
        self._device: Optional[torch.device] = None
        self._status: Optional[torch.Tensor] = None
        self._in_trajectory = False
        self._last_time_step = False
        self._corrector_phase = 0
        self._seed_is_set = False
        self._standalone_calls = 0

    # ---- device-side bookkeeping ---------------------------------------------------------------------------
    def _tables_on(self, device: torch.device):
        key = str(device)
        if key not in self._device_tables:
            self._device_tables[key] = tuple(
                getattr(self.noise, name).detach().to(device=device, dtype=torch.float32).contiguous()
                for name in ("q_matrix", "q_bar_matrix", "q_bar_tm1_matrix"))
        return self._device_tables[key]

    def _status_on(self, device: torch.device) -> torch.Tensor:
        if self._status is None or self._status.device != device:
            self._status = torch.zeros(1, dtype=torch.int32, device=device)
        return self._status

    def _raise_on_status(self) -> None:
        if self._status is not None:
            _cabi.raise_on_status(self._status, AssertionError)
        if hasattr(self.axl_network, "raise_on_device_status"):
            self.axl_network.raise_on_device_status()

    def _begin_trajectory(self, device: torch.device) -> None:
        self._device = device
        if self.seed is None:
            # consume the global CPU generator, so torch.manual_seed() controls reproducibility as in the reference
            self._active_seed = int(torch.randint(0, 2**62, (1,)).item())
        else:
            self._active_seed = int(self.seed)
        self.trajectory_initializer.seed = self._active_seed
        self.trajectory_initializer.sample_offset = int(self.sample_offset)

    def _begin_standalone_step(self) -> None:
        """``predictor_step`` / ``corrector_step`` called outside ``sample()``: every call draws fresh noise, as in
        the reference (whose draws advance the global generator).  The draws are keyed by (seed, step index), so the
        seed itself has to change from call to call: taken from torch's global CPU generator (``torch.manual_seed``
        controls it) or, when ``self.seed`` is set, derived from it and a call counter."""
        self._standalone_calls += 1
        if self.seed is None:
            self._active_seed = int(torch.randint(0, 2**62, (1,)).item())
        else:
            self._active_seed = (int(self.seed) + 0x9E3779B97F4A7C15 * self._standalone_calls) % 2**64

    def _phase_id(self, index_i: int, corrector_phase: int) -> int:
        """Step key of the counter-based draws: i * (1 + n_corr) + phase (0 = predictor, k+1 = k-th corrector)."""
        return int(index_i) * (1 + self.number_of_corrector_steps) + corrector_phase

    # ---- noise hooks (device-side, counter-based) ----------------------------------------------------------------
    def _draw(self, kind: str, number_of_samples: int, units: int, unit_size: int, draw_id: int) -> torch.Tensor:
        device = self._device
        if device is None or device.type != "cuda":
            raise _cabi.B200MDError("noise is drawn by the sm_100a extension; call the generator with CUDA tensors")
        shape = (number_of_samples, units, unit_size) if units > 0 else (number_of_samples, unit_size)
        out = torch.empty(shape, dtype=torch.float32, device=device)
        lib = _cabi.lib()
        args = (_cabi.ptr(out), number_of_samples, max(units, 1), unit_size)
        tail = (self._active_seed, self._step_id, draw_id, int(self.sample_offset), _cabi.stream_ptr(device))
        if kind == "normal":
            _cabi.check(lib.b200md_draw_normal(*args, *tail))
        elif kind == "uniform":
            _cabi.check(lib.b200md_draw_uniform(*args, *tail))
        else:
            _cabi.check(lib.b200md_draw_gumbel(*args, float(self.small_epsilon), *tail))
        return out

    def _draw_coordinates_gaussian_sample(self, number_of_samples):
        return self._draw("normal", number_of_samples, self.number_of_atoms, self.spatial_dimension,
                          _cabi.DRAW_COORD_NORMAL)

    def _draw_lattice_gaussian_sample(self, number_of_samples):
        draw_id = _cabi.DRAW_LATTICE_NORMAL_SECOND if getattr(self, "_second_lattice_draw", False) \
            else _cabi.DRAW_LATTICE_NORMAL
        return self._draw("normal", number_of_samples, 0, self.num_lattice_parameters, draw_id)

    def _draw_gumbel_sample(self, number_of_samples):
        return self._draw("gumbel", number_of_samples, self.number_of_atoms, self.num_classes, _cabi.DRAW_GUMBEL)

    def _draw_binary_sample(self, number_of_samples):
        return self._draw("uniform", number_of_samples, self.number_of_atoms, 1,
                          _cabi.DRAW_BINARY).reshape(number_of_samples, self.number_of_atoms)

    # ---- score network ------------------------------------------------------------------------------------------
    def _get_model_predictions(self, composition: AXL, time: float, sigma_noise: float,
                               cartesian_forces: torch.Tensor) -> AXL:
        """Evaluate the score network at a common (time, sigma) (reference :113-153)."""
        number_of_samples = composition.X.shape[0]
        ones = torch.ones(number_of_samples, 1, dtype=composition.X.dtype, device=composition.X.device)
        batch = {NOISY_AXL_COMPOSITION: composition, TIME: _as_float(time) * ones,
                 NOISE: _as_float(sigma_noise) * ones, CARTESIAN_FORCES: cartesian_forces}
        return self.axl_network(batch, conditional=False)

    # ---- coordinates ------------------------------------------------------------------------------------------------
    def _relative_coordinates_update(self, relative_coordinates: torch.Tensor, sigma_normalized_scores: torch.Tensor,
                                     sigma_i, score_weight, gaussian_noise_weight, z: torch.Tensor) -> torch.Tensor:
        """x <- wrap(x + score_weight * s / sigma_i + noise_weight * z), one kernel (reference :155-201)."""
        device = _cabi.require_cuda(relative_coordinates, sigma_normalized_scores)
        self._device = device
        number_of_samples, number_of_atoms, d = relative_coordinates.shape
        if z is None:
            z = self._draw_coordinates_gaussian_sample(number_of_samples)
        x = _cabi.as_f32(relative_coordinates)
        s = _cabi.as_f32(sigma_normalized_scores)
        zz = _cabi.as_f32(z.to(device))
        out = torch.empty_like(x)
        if _on_device(score_weight) or _on_device(gaussian_noise_weight):  # e.g. the adaptive corrector's step size
            weights = _device_weights(score_weight, gaussian_noise_weight, device)
            _cabi.check(_cabi.lib().b200md_pc_update_device_weights(
                _cabi.ptr(x), _cabi.ptr(s), _cabi.ptr(zz), _cabi.ptr(weights), _as_float(sigma_i), x.numel(), 1,
                _cabi.ptr(out), _cabi.stream_ptr(device)))
            return out
        _cabi.check(_cabi.lib().b200md_pc_coordinates_update(
            _cabi.ptr(x), _cabi.ptr(s), _cabi.ptr(zz), _as_float(score_weight), _as_float(sigma_i),
            _as_float(gaussian_noise_weight), number_of_samples, number_of_atoms, d, self._active_seed, self._step_id,
            int(self.sample_offset), _cabi.ptr(out), _cabi.stream_ptr(device)))
        return out

    def _relative_coordinates_update_predictor_step(self, relative_coordinates, sigma_normalized_scores, sigma_i,
                                                    score_weight, gaussian_noise_weight, z):
        return self._relative_coordinates_update(relative_coordinates, sigma_normalized_scores, sigma_i, score_weight,
                                                 gaussian_noise_weight, z)

    def _relative_coordinates_update_corrector_step(self, relative_coordinates, sigma_normalized_scores, sigma_i,
                                                    score_weight, gaussian_noise_weight, z):
        return self._relative_coordinates_update(relative_coordinates, sigma_normalized_scores, sigma_i, score_weight,
                                                 gaussian_noise_weight, z)

    # ---- atom types ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _single_matrix(m: torch.Tensor, device: torch.device) -> torch.Tensor:
        """The reference broadcasts one [C,C] matrix to [B,N,C,C] (:572-591); the kernel reads the [C,C] original."""
        while m.dim() > 2:
            m = m[0]
        return _cabi.as_f32(m).to(device)

    def _atom_types_update(self, predicted_logits: torch.Tensor, atom_types_i: torch.LongTensor,
                           q_matrices_i: torch.Tensor, q_bar_matrices_i: torch.Tensor,
                           q_bar_tm1_matrices_i: torch.Tensor, atom_type_greedy_sampling: bool,
                           one_atom_type_transition_per_step: bool) -> torch.LongTensor:
        """a_{i-1} ~ p(a_{i-1} | a_i, logits): one kernel, one CTA per sample (reference :247-439)."""
        device = _cabi.require_cuda(predicted_logits, atom_types_i)
        self._device = device
        number_of_samples, number_of_atoms, num_classes = predicted_logits.shape
        gumbel = _cabi.as_f32(self._draw_gumbel_sample(number_of_samples).to(device))
        binary = None
        if atom_type_greedy_sampling:
            binary = _cabi.as_f32(self._draw_binary_sample(number_of_samples).to(device))
        logits = _cabi.as_f32(predicted_logits)
        a_in = _cabi.as_i64(atom_types_i)
        a_out = torch.empty_like(a_in)
        probs = torch.empty_like(logits) if self.record_atom_type_update else None
        flags = (_cabi.ATOM_GREEDY if atom_type_greedy_sampling else 0) | \
                (_cabi.ATOM_ONE_TRANSITION if one_atom_type_transition_per_step else 0) | \
                (_cabi.ATOM_LAST_STEP if self._last_time_step else 0)
        q, qb, qbm = (self._single_matrix(m, device) for m in (q_matrices_i, q_bar_matrices_i, q_bar_tm1_matrices_i))
        _cabi.check(_cabi.lib().b200md_atom_types_update(
            _cabi.ptr(logits), _cabi.ptr(a_in), _cabi.ptr(q), _cabi.ptr(qb), _cabi.ptr(qbm), _cabi.ptr(gumbel),
            _cabi.ptr(binary), flags, float(self.small_epsilon), number_of_samples, number_of_atoms, num_classes,
            self._active_seed, self._step_id, int(self.sample_offset), _cabi.ptr(a_out), _cabi.ptr(probs),
            _cabi.ptr(self._status_on(device)), _cabi.stream_ptr(device)))
        if self.record_atom_type_update:
            entry = dict(predicted_logits=predicted_logits.detach().cpu(),
                         one_step_transition_probabilities=probs.cpu(), gumbel_sample=gumbel.cpu(),
                         a_i=atom_types_i.cpu(), a_im1=a_out.cpu())
            self.sample_trajectory_recorder.record(key="atom_type_update", entry=entry)
        return a_out

    # ---- lattice ------------------------------------------------------------------------------------------------------
    def _lattice_parameters_update(self, lattice_parameters: torch.Tensor, sigma_normalized_scores: torch.Tensor,
                                   sigma_n_i, score_weight, gaussian_noise_weight,
                                   z: Optional[torch.Tensor] = None) -> torch.Tensor:
        """l <- l + score_weight * s / sigma_n + noise_weight * z; identity for a fixed lattice (reference :441-490)."""
        if self.use_fixed_lattice_parameters:
            return lattice_parameters
        device = _cabi.require_cuda(lattice_parameters, sigma_normalized_scores)
        self._device = device
        number_of_samples, n_lp = lattice_parameters.shape
        if z is None:
            z = self._draw_lattice_gaussian_sample(number_of_samples)
        lattice = _cabi.as_f32(lattice_parameters)
        s = _cabi.as_f32(sigma_normalized_scores)
        zz = _cabi.as_f32(z.to(device))
        out = torch.empty_like(lattice)
        if _on_device(score_weight) or _on_device(gaussian_noise_weight):
            weights = _device_weights(score_weight, gaussian_noise_weight, device)
            _cabi.check(_cabi.lib().b200md_pc_update_device_weights(
                _cabi.ptr(lattice), _cabi.ptr(s), _cabi.ptr(zz), _cabi.ptr(weights), _as_float(sigma_n_i),
                lattice.numel(), 0, _cabi.ptr(out), _cabi.stream_ptr(device)))
            return out
        _cabi.check(_cabi.lib().b200md_pc_lattice_update(
            _cabi.ptr(lattice), _cabi.ptr(s), _cabi.ptr(zz), _as_float(score_weight), _as_float(sigma_n_i),
            _as_float(gaussian_noise_weight), number_of_samples, n_lp, self._active_seed, self._step_id,
            int(self.sample_offset), _cabi.ptr(out), _cabi.stream_ptr(device)))
        return out

    def _lattice_parameters_update_predictor_step(self, lattice_parameters, sigma_normalized_scores, sigma_n_i,
                                                  score_weight, gaussian_noise_weight, z=None):
        return self._lattice_parameters_update(lattice_parameters, sigma_normalized_scores, sigma_n_i, score_weight,
                                               gaussian_noise_weight, z)

    def _lattice_parameters_update_corrector_step(self, lattice_parameters, sigma_normalized_scores, sigma_n_i,
                                                  score_weight, gaussian_noise_weight, z=None):
        return self._lattice_parameters_update(lattice_parameters, sigma_normalized_scores, sigma_n_i, score_weight,
                                               gaussian_noise_weight, z)

    # ---- predictor / corrector -----------------------------------------------------------------------------------
    def _defer_network_checks(self, defer: bool) -> None:
        if hasattr(self.axl_network, "defer_status_check"):
            self.axl_network.defer_status_check = defer

    def predictor_step(self, composition_i: AXL, index_i: int, cartesian_forces: torch.Tensor) -> AXL:
        """From time index ``index_i`` to ``index_i - 1`` (reference :536-645)."""
        assert 1 <= index_i <= self.number_of_discretization_steps, \
            "The predictor step can only be invoked for index_i between 1 and the total number of discretization steps."
        device = composition_i.X.device
        self._device = device
        if not self._in_trajectory:
            self._begin_standalone_step()
        number_of_samples, number_of_atoms, d = composition_i.X.shape
        idx = index_i - 1
        self._step_id = self._phase_id(idx, 0)
        t_i = float(self._host["time"][idx])
        g_i = float(self._host["g"][idx])
        g2_i = float(self._host["g_squared"][idx])
        sigma_i = float(self._host["sigma"][idx])
        sigma_n_i = float(self._host["sigma"][idx] / (number_of_atoms ** (1 / d)))
        q, q_bar, q_bar_tm1 = (table[idx] for table in self._tables_on(device))

        self._defer_network_checks(True)
        model_predictions_i = self._get_model_predictions(composition_i, t_i, sigma_i, cartesian_forces)

        this_is_last_time_step = idx == 0
        one_transition = self.one_atom_type_transition_per_step and not this_is_last_time_step
        # reference :616-620 asserts on the host that no MASK is left after the last step; here the kernel raises
        # STATUS_MASK_REMAINS in the device status word.
        self._last_time_step = this_is_last_time_step
        try:
            a_im1 = self._atom_types_update(model_predictions_i.A, composition_i.A, q, q_bar, q_bar_tm1,
                                            atom_type_greedy_sampling=self.atom_type_greedy_sampling,
                                            one_atom_type_transition_per_step=one_transition)
        finally:
            self._last_time_step = False

        z_coordinates = self._draw_coordinates_gaussian_sample(number_of_samples).to(composition_i.X)
        x_im1 = self._relative_coordinates_update_predictor_step(composition_i.X, model_predictions_i.X, sigma_i,
                                                                 g2_i, g_i, z_coordinates)
        if self.use_fixed_lattice_parameters:
            lp_im1 = composition_i.L
        else:
            z_lattice = self._draw_lattice_gaussian_sample(number_of_samples).to(composition_i.L)
            lp_im1 = self._lattice_parameters_update_predictor_step(composition_i.L, model_predictions_i.L, sigma_n_i,
                                                                    g2_i, g_i, z_lattice)
        composition_im1 = AXL(A=a_im1, X=x_im1, L=lp_im1)
        if self.record:
            self._record_predictor_step(composition_i, composition_im1, index_i, model_predictions_i)
        if not self._in_trajectory:
            self._defer_network_checks(False)
            self._raise_on_status()
        return composition_im1

    def _get_coordinates_corrector_step_size(self, index_i: int, sigma_i, model_predictions_i: torch.Tensor,
                                             z: torch.Tensor):
        """eps_i from the tabulated Langevin dynamics, indexed by i itself (reference :669-679)."""
        return float(self._host["epsilon"][index_i])

    def _get_lattice_parameters_corrector_step_size(self, index_i: int, sigma_n_i, model_predictions_i: torch.Tensor,
                                                    z: torch.Tensor):
        return float(self._host["epsilon"][index_i])

    def corrector_step(self, composition_i: AXL, index_i: int, cartesian_forces: torch.Tensor) -> AXL:
        """Langevin corrector at time index ``index_i`` (reference :693-805)."""
        assert 0 <= index_i <= self.number_of_discretization_steps - 1, (
            "The corrector step can only be invoked for index_i between 0 and "
            "the total number of discretization steps minus 1.")
        device = composition_i.X.device
        self._device = device
        if not self._in_trajectory:
            self._begin_standalone_step()
        number_of_samples, number_of_atoms, d = composition_i.X.shape
        phase = self._corrector_phase
        self._step_id = self._phase_id(index_i, 1 + phase)
        if index_i == 0:
            sigma_i = float(torch.tensor(self.noise_parameters.sigma_min, dtype=torch.float32))
            t_i = 0.0
            idx = index_i
        else:
            idx = index_i - 1
            sigma_i = float(self._host["sigma"][idx])
            t_i = float(self._host["time"][idx])
        sigma_n_i = float(torch.tensor(sigma_i, dtype=torch.float32) / (number_of_atoms ** (1 / d)))

        self._defer_network_checks(True)
        model_predictions_i = self._get_model_predictions(composition_i, t_i, sigma_i, cartesian_forces)
        z_coordinates = self._draw_coordinates_gaussian_sample(number_of_samples).to(composition_i.X)
        eps_i = self._get_coordinates_corrector_step_size(index_i, sigma_i, model_predictions_i.X, z_coordinates)
        sqrt_2eps_i = torch.sqrt(2 * torch.as_tensor(eps_i, dtype=torch.float32))
        corrected_x_i = self._relative_coordinates_update_corrector_step(
            composition_i.X, model_predictions_i.X, sigma_i, eps_i, sqrt_2eps_i, z_coordinates)

        if self.use_fixed_lattice_parameters:
            corrected_lp_i = composition_i.L
        else:
            z_lattice = self._draw_lattice_gaussian_sample(number_of_samples).to(composition_i.L)
            eps_l = self._get_lattice_parameters_corrector_step_size(index_i, sigma_n_i, model_predictions_i.L,
                                                                     z_lattice)
            sqrt_2eps_l = torch.sqrt(2 * torch.as_tensor(eps_l, dtype=torch.float32))
            # As in the reference (:758-777) the update is NOT given z_lattice: that draw only feeds the step size (the
            # adaptive corrector uses it) and the update draws its own noise -- a second, independent lattice draw.
            self._second_lattice_draw = True
            try:
                corrected_lp_i = self._lattice_parameters_update(composition_i.L, model_predictions_i.L, sigma_n_i,
                                                                 eps_l, sqrt_2eps_l)
            finally:
                self._second_lattice_draw = False

        if self.atom_type_transition_in_corrector:
            q, q_bar, q_bar_tm1 = (table[idx] for table in self._tables_on(device))
            corrected_a_i = self._atom_types_update(
                model_predictions_i.A, composition_i.A, q, q_bar, q_bar_tm1,
                atom_type_greedy_sampling=self.atom_type_greedy_sampling,
                one_atom_type_transition_per_step=self.one_atom_type_transition_per_step)
        else:
            corrected_a_i = composition_i.A
        corrected = AXL(A=corrected_a_i, X=corrected_x_i, L=corrected_lp_i)
        if self.record_corrector:
            self._record_corrector_step(composition_i, corrected, index_i, model_predictions_i)
        if not self._in_trajectory:
            self._defer_network_checks(False)
            self._raise_on_status()
        return corrected

    # ---- trajectory drivers ---------------------------------------------------------------------------------------
    def sample(self, number_of_samples: int, device: torch.device) -> AXL:
        self._begin_trajectory(torch.device(device))
        self._seed_is_set = True
        try:
            return super().sample(number_of_samples, device)
        finally:
            self._seed_is_set = False

    def sample_from_noisy_composition(self, starting_noisy_composition: AXL, starting_step_index: int,
                                      ending_step_index: int) -> AXL:
        if not self._seed_is_set:
            self._begin_trajectory(starting_noisy_composition.X.device)
        self._in_trajectory = True
        try:
            return self._sample_loop(starting_noisy_composition, starting_step_index, ending_step_index)
        finally:
            self._in_trajectory = False
            self._defer_network_checks(False)

    def _sample_loop(self, start: AXL, starting_step_index: int, ending_step_index: int) -> AXL:
        assert starting_step_index > ending_step_index, \
            "It is nonsensical for starting_step_index to be smaller or equal to ending_step_index."
        assert starting_step_index > 0, "Starting step should be larger than zero."
        assert ending_step_index >= 0, "ending step should be larger or equal to zero."
        fused = self._run_fused(start, starting_step_index, ending_step_index)
        if fused is not None:
            return fused
        for attempt in range(2):
            composition_ip1 = start
            forces = torch.zeros_like(composition_ip1.X)
            composition_i = composition_ip1
            for i in range(starting_step_index - 1, max(ending_step_index, 0) - 1, -1):
                composition_i = self.predictor_step(composition_ip1, i + 1, forces)
                for k in range(self.number_of_corrector_steps):
                    self._corrector_phase = k
                    composition_i = self.corrector_step(composition_i, i, forces)
                self._corrector_phase = 0
                composition_ip1 = composition_i
            # an fp16 overflow inside the network (IMPL_AUTO): the trajectory is run again from `start` with the
            # wider-range kernels; the counter-based draws make it the trajectory those kernels give from the outset
            if attempt == 0 and self._network_widened_its_range():
                continue
            break
        self._raise_on_status()
        return composition_i

    def _network_widened_its_range(self) -> bool:
        widen = getattr(self.axl_network, "widen_range_after_overflow", None)
        if widen is None or not widen():
            return False
        if self._status is not None:
            self._status.zero_()  # bits raised downstream of the overflow (NaN logits, MASK left): the re-run decides
        return True

    @classmethod
    def _fused_classes(cls):
        """Classes whose implementations of the hooks are the ones ``b200md_pc_sample`` reproduces."""
        return (LangevinGenerator,)

    def _hooks_overridden(self) -> bool:
        """True when a subclass or an instance-level patch replaces any step-level method."""
        known = self._fused_classes()
        for name in self._HOOKS:
            if name in self.__dict__:  # instance-level patch (e.g. unittest.mock.patch.object)
                return True
            implementation = getattr(type(self), name, None)
            if implementation is None or not any(c.__dict__.get(name) is implementation for c in known):
                return True
        return False

    def _constraint_arguments(self, device: torch.device):
        """(n_constrained, x0, a0, indices) for b200md_pc_sample; none for the unconstrained generator."""
        return 0, None, None, None

    def _run_fused(self, start: AXL, starting_step_index: int, ending_step_index: int) -> Optional[AXL]:
        """Whole trajectory in one C-ABI call (``b200md_pc_sample``) when nothing requires the step-by-step loop."""
        from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import \
            EGNNScoreNetwork
        network = self.axl_network
        if not (self.allow_fused and type(network) is EGNNScoreNetwork and self.use_fixed_lattice_parameters
                and not self.record and self.spatial_dimension == 3 and start.X.is_cuda
                and not self._hooks_overridden()):
            return None
        device = _cabi.require_cuda(start.X, start.A, start.L)
        self._device = device
        lib = _cabi.lib()
        number_of_samples, number_of_atoms, _ = start.X.shape
        lattice = _cabi.as_f32(start.L)
        cfg = network._config()
        packed = network.packed_weights(device)
        capacity = network.resolve_edge_capacity(number_of_samples, number_of_atoms)
        workspace = network.workspace(device, number_of_samples, number_of_atoms, capacity, for_sampler=True)
        q, q_bar, q_bar_tm1 = self._tables_on(device)
        n_constrained, x0_known, a0_known, c_indices = self._constraint_arguments(device)
        sigma_min32 = float(torch.tensor(self.noise_parameters.sigma_min, dtype=torch.float32))
        schedule = _cabi.PcSchedule(
            total_time_steps=self.number_of_discretization_steps,
            number_of_corrector_steps=self.number_of_corrector_steps,
            start_step_index=int(starting_step_index), end_step_index=int(max(ending_step_index, 0)),
            atom_flags=(_cabi.ATOM_GREEDY if self.atom_type_greedy_sampling else 0)
            | (_cabi.ATOM_ONE_TRANSITION if self.one_atom_type_transition_per_step else 0),
            atom_type_transition_in_corrector=int(self.atom_type_transition_in_corrector),
            small_epsilon=float(self.small_epsilon), sigma_min=sigma_min32,
            sigma_host=self._host["sigma"].data_ptr(), g_host=self._host["g"].data_ptr(),
            g2_host=self._host["g_squared"].data_ptr(), epsilon_host=self._host["epsilon"].data_ptr(),
            sqrt_2_epsilon_host=self._host["sqrt_2_epsilon"].data_ptr(),
            q=q.data_ptr(), q_bar=q_bar.data_ptr(), q_bar_tm1=q_bar_tm1.data_ptr(),
            n_constrained=n_constrained,
            x0_known=x0_known.data_ptr() if x0_known is not None else None,
            a0_known=a0_known.data_ptr() if a0_known is not None else None,
            constrained_indices=c_indices.data_ptr() if c_indices is not None else None)
        # b200md_pc_sample and the network's own forward share ONE status word (the network's), so that an fp16
        # overflow is seen by widen_range_after_overflow() whichever path raised it
        status = network.device_status(device)
        for attempt in range(2):
            x = _cabi.as_f32(start.X).clone()
            a = _cabi.as_i64(start.A).clone()
            _cabi.check(lib.b200md_pc_sample(
                ctypes.byref(cfg), _cabi.ptr(packed), ctypes.byref(schedule), _cabi.ptr(a), _cabi.ptr(x),
                _cabi.ptr(lattice), number_of_samples, number_of_atoms, self._active_seed, int(self.sample_offset),
                _cabi.ptr(workspace), workspace.numel(), capacity, _cabi.ptr(status), network.effective_kernel_impl(),
                _cabi.stream_ptr(device)))
            # IMPL_AUTO and an activation beyond fp16: same trajectory (same counter-based noise) on the 3xTF32 kernels
            if attempt == 0 and self._network_widened_its_range():
                continue
            break
        self._raise_on_status()
        return AXL(A=a, X=x, L=start.L)

    # ---- recording ----------------------------------------------------------------------------------------------------
    def _record_step(self, key: str, names: Tuple[str, str, str], axls, index_i: int) -> None:
        entry = dict(time_step_index=index_i)
        for name, axl in zip(names, axls):
            entry[name] = AXL(A=axl.A.detach().cpu(), X=axl.X.detach().cpu(), L=axl.L.detach().cpu())
        self.sample_trajectory_recorder.record(key=key, entry=entry)

    def _record_predictor_step(self, composition_i, composition_im1, index_i, model_predictions_i):
        self._record_step("predictor_step", ("composition_i", "composition_im1", "model_predictions_i"),
                          (composition_i, composition_im1, model_predictions_i), index_i)

    def _record_corrector_step(self, composition_i, corrected_composition_i, index_i, model_predictions_i):
        self._record_step("corrector_step", ("composition_i", "corrected_composition_i", "model_predictions_i"),
                          (composition_i, corrected_composition_i, model_predictions_i), index_i)

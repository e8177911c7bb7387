"""Stillinger-Weber energy oracle on the GPU.

Drop-in for the reference's ``LammpsEnergyOracle`` (oracle/lammps_energy_oracle.py:30-193) where only energies are
needed (the energy Kolmogorov-Smirnov metric of sampled structures): same parameter dataclass (``name: lammps``,
``sw_coeff_filename``), same LAMMPS ``pair_style sw`` coefficient files, same functional form -- evaluated for the whole
batch by one kernel (``b200md_stillinger_weber_energy``, csrc/structure.cu) instead of one LAMMPS run per structure.
Forces are not produced (nothing on the sampling path consumes them).
"""
import os
import warnings
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.oracle import SW_COEFFICIENTS_DIR
from diffusion_for_multi_scale_molecular_dynamics_b200.oracle.energy_oracle import EnergyOracle, OracleParameters

MIN_BOX_SIZE = 3.0  # the reference skips (energy 0) boxes with a side below 3 Angstrom (lammps_energy_oracle.py:121-127)
SW_FIELDS = ("epsilon", "sigma", "a", "lambda", "gamma", "costheta0", "A", "B", "p", "q", "tol")


@dataclass(kw_only=True)
class LammpsOracleParameters(OracleParameters):
    """Same fields as the reference's LammpsOracleParameters."""

    name: str = "lammps"
    sw_coeff_filename: str  # Stillinger-Weber potential filename


def read_stillinger_weber_file(path: str) -> Dict[Tuple[str, str, str], Dict[str, float]]:
    """Parse a LAMMPS ``pair_style sw`` coefficient file: entries may span several lines, '#' starts a comment."""
    tokens: List[str] = []
    with open(path, "r") as fd:
        for line in fd:
            tokens += line.split("#", 1)[0].split()
    entry_size = 3 + len(SW_FIELDS)
    assert len(tokens) % entry_size == 0, f"The SW file '{path}' is malformed: incomplete entry."
    entries = {}
    for start in range(0, len(tokens), entry_size):
        chunk = tokens[start:start + entry_size]
        entries[tuple(chunk[:3])] = {name: float(value) for name, value in zip(SW_FIELDS, chunk[3:])}
    return entries


def build_parameter_table(entries: Dict[Tuple[str, str, str], Dict[str, float]], elements: List[str]) -> np.ndarray:
    """[n, n, n, 10] float32 table in the layout of b200md_stillinger_weber_energy (derived constants as in LAMMPS
    pair_sw::setup_params: cut = a sigma, sigma_gamma, lambda_epsilon, c5 = A eps B sigma^p, c6 = A eps sigma^q)."""
    n = len(elements)
    table = np.zeros((n, n, n, 10), dtype=np.float64)
    for i, ei in enumerate(elements):
        for j, ej in enumerate(elements):
            for k, ek in enumerate(elements):
                assert (ei, ej, ek) in entries, f"The SW file has no entry for the triplet {ei} {ej} {ek}."
                e = entries[(ei, ej, ek)]
                cut = e["a"] * e["sigma"]
                table[i, j, k] = [cut, cut * cut, e["sigma"], e["sigma"] * e["gamma"], e["lambda"] * e["epsilon"],
                                  e["costheta0"], e["A"] * e["epsilon"] * e["B"] * e["sigma"] ** e["p"],
                                  e["A"] * e["epsilon"] * e["sigma"] ** e["q"], e["p"], e["q"]]
    return table.astype(np.float32)


class StillingerWeberEnergyOracle(EnergyOracle):
    """Batched Stillinger-Weber energies on the GPU."""

    def __init__(self, lammps_oracle_parameters: LammpsOracleParameters, sw_coefficients_dir: Path = SW_COEFFICIENTS_DIR):
        super().__init__(lammps_oracle_parameters)
        self.sw_coefficients_file_path = str(Path(sw_coefficients_dir) / lammps_oracle_parameters.sw_coeff_filename)
        assert os.path.isfile(self.sw_coefficients_file_path), \
            f"The SW file '{self.sw_coefficients_file_path}' does not exist."
        entries = read_stillinger_weber_file(self.sw_coefficients_file_path)
        self._table = torch.from_numpy(build_parameter_table(entries, self._element_types.elements))
        self._max_cutoff = float(self._table[..., 0].max())
        self._device_table = None

    def _compute_batch_energies(self, relative_coordinates: torch.Tensor, lattice_parameters: torch.Tensor,
                                atom_types: torch.Tensor) -> torch.Tensor:
        device = _cabi.require_cuda(relative_coordinates, lattice_parameters, atom_types)
        x = _cabi.as_f32(relative_coordinates)
        lattice = _cabi.as_f32(lattice_parameters).clone()
        assert x.dim() == 3 and x.shape[-1] == 3 and lattice.shape == (x.shape[0], 6), \
            "the sm_100a Stillinger-Weber kernel supports spatial_dimension == 3 only"
        if bool((lattice[:, :3] < 0).any()):  # reference energy_oracle.py:104-110
            warnings.warn("Got a negative lattice parameter. Clipping to 1.0 Angstrom")
            negative = (lattice[:, :3] < 0).any(dim=1)
            lattice[negative, :3] = lattice[negative, :3].clip(min=1.0)
        types = atom_types.to(device=device, dtype=torch.int64).contiguous()
        if self._device_table is None or self._device_table.device != device:
            self._device_table = self._table.to(device).contiguous()
        energies = torch.empty(x.shape[0], dtype=torch.float32, device=device)
        status = torch.zeros(1, dtype=torch.int32, device=device)
        _cabi.check(_cabi.lib().b200md_stillinger_weber_energy(
            _cabi.ptr(x), _cabi.ptr(lattice), _cabi.ptr(types), _cabi.ptr(self._device_table),
            self._element_types.number_of_atom_types, self._max_cutoff, MIN_BOX_SIZE, x.shape[0], x.shape[1],
            _cabi.ptr(energies), _cabi.ptr(status), _cabi.stream_ptr(device)))
        _cabi.raise_on_status(status, AssertionError)
        return energies

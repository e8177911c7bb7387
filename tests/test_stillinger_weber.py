"""Stillinger-Weber energies.  The reference obtains them from LAMMPS, which is not available
here: the CPU oracle (oracle/sw_oracle.py) is pinned by known answers, the CUDA kernel is checked against the oracle."""
import itertools

import numpy as np
import pytest
import torch

from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (ATOM_TYPES, AXL, AXL_COMPOSITION,
                                                                       LATTICE_PARAMETERS, RELATIVE_COORDINATES)
from diffusion_for_multi_scale_molecular_dynamics_b200.oracle import SW_COEFFICIENTS_DIR
from diffusion_for_multi_scale_molecular_dynamics_b200.oracle.energy_oracle_factory import (create_energy_oracle,
                                                                                          create_energy_oracle_parameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.oracle.stillinger_weber_energy_oracle import (
    StillingerWeberEnergyOracle, build_parameter_table, read_stillinger_weber_file)
from oracle import sw_oracle

DIAMOND = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.25, .25, .25], [.25, .75, .75], [.75, .25, .75],
                    [.75, .75, .25]])


def diamond(cells: int) -> np.ndarray:
    grid = np.array(list(itertools.product(range(cells), repeat=3)))[:, None, :]
    return ((grid + DIAMOND[None]) / cells).reshape(-1, 3)


def si_entries():
    return read_stillinger_weber_file(str(SW_COEFFICIENTS_DIR / "Si.sw"))


def test_coefficient_files_parse():
    si = si_entries()
    assert list(si) == [("Si", "Si", "Si")] and si[("Si", "Si", "Si")]["epsilon"] == 2.1683
    sige = read_stillinger_weber_file(str(SW_COEFFICIENTS_DIR / "SiGe.sw"))
    assert len(sige) == 8 and sige[("Si", "Ge", "Si")]["lambda"] == 23.1
    table = build_parameter_table(sige, ["Ge", "Si"])  # ElementTypes sorts: Ge = 0, Si = 1
    assert table.shape == (2, 2, 2, 10)
    np.testing.assert_allclose(table[1, 0, 0, 0], 1.80 * 2.138)          # cut of the Si-Ge pair
    np.testing.assert_allclose(table[0, 1, 0, 4], 28.1 * 3.178, rtol=1e-6)   # lambda * epsilon of Ge Si Ge
    with pytest.raises(AssertionError):
        build_parameter_table(si, ["Ge", "Si"])


def test_oracle_known_answers():
    """Perfect diamond Si: every bond sits at the pair minimum (-eps) and every angle is tetrahedral, so E = -2 eps per
    atom.  Dimer and trimer: closed forms."""
    entries, elements = si_entries(), ["Si"]
    e = entries[("Si", "Si", "Si")]
    a = 2.0 ** (1.0 / 6.0) * e["sigma"] * 4.0 / np.sqrt(3.0)   # nearest neighbour at the pair minimum 2^(1/6) sigma
    positions = diamond(2) * 2 * a
    energy = sw_oracle.energy(positions, np.full(3, 2 * a), np.zeros(64, dtype=int), elements, entries)
    np.testing.assert_allclose(energy / 64, -2.0 * e["epsilon"], rtol=1e-6)
    r = 2.5
    box = np.full(3, 50.0)
    phi2 = e["A"] * e["epsilon"] * (e["B"] * (e["sigma"] / r) ** 4 - 1.0) * np.exp(e["sigma"] / (r - e["a"] * e["sigma"]))
    dimer = sw_oracle.energy(np.array([[10.0, 10, 10], [10 + r, 10, 10]]), box, np.zeros(2, dtype=int), elements, entries)
    np.testing.assert_allclose(dimer, phi2, rtol=1e-12)
    # right-angle trimer: two bonds of length r, third distance r sqrt(2) = 3.54 < cutoff 3.77: three pairs, three angles
    trimer = np.array([[10.0, 10, 10], [10 + r, 10, 10], [10, 10 + r, 10]])
    rr = r * np.sqrt(2.0)
    phi2_long = e["A"] * e["epsilon"] * (e["B"] * (e["sigma"] / rr) ** 4 - 1.0) * np.exp(e["sigma"] / (rr - e["a"] * e["sigma"]))
    g = lambda x: np.exp(e["gamma"] * e["sigma"] / (x - e["a"] * e["sigma"]))  # noqa: E731
    le = e["lambda"] * e["epsilon"]
    phi3 = le * (0.0 + 1 / 3) ** 2 * g(r) * g(r) + 2 * le * (np.cos(np.pi / 4) + 1 / 3) ** 2 * g(r) * g(rr)
    np.testing.assert_allclose(sw_oracle.energy(trimer, box, np.zeros(3, dtype=int), elements, entries),
                               2 * phi2 + phi2_long + phi3, rtol=1e-12)


def test_factory_and_no_cpu_fallback():
    parameters = create_energy_oracle_parameters(dict(name="lammps", sw_coeff_filename="Si.sw"), ["Si"])
    oracle_ = create_energy_oracle(parameters)
    assert isinstance(oracle_, StillingerWeberEnergyOracle)
    with pytest.raises(AssertionError):
        create_energy_oracle_parameters(dict(name="dft"), ["Si"])
    samples = {AXL_COMPOSITION: AXL(A=torch.zeros(1, 8, dtype=torch.long), X=torch.rand(1, 8, 3),
                                    L=torch.tensor([[5.43, 5.43, 5.43, 0, 0, 0]]))}
    with pytest.raises(_cabi.B200MDError):
        oracle_.compute_oracle_energies(samples)
    with pytest.raises(AssertionError):
        oracle_.compute_oracle_energies({RELATIVE_COORDINATES: torch.rand(1, 8, 3)})


@pytest.mark.gpu
def test_gpu_silicon_energies_against_oracle(cuda_device):
    """Diamond Si (the -2 eps known answer), thermally displaced crystals and uniform random structures."""
    entries = si_entries()
    oracle_ = create_energy_oracle(create_energy_oracle_parameters(dict(name="lammps", sw_coeff_filename="Si.sw"), ["Si"]))
    rng = np.random.default_rng(4)
    length = 10.86
    x = np.stack([diamond(2)] + [np.mod(diamond(2) + rng.normal(scale=s, size=(64, 3)), 1.0) for s in (0.01, 0.03)]
                 + [rng.random((64, 3)) for _ in range(3)])
    lattice = torch.zeros(len(x), 6)
    lattice[:, :3] = length
    energies = oracle_.compute_oracle_energies({RELATIVE_COORDINATES: torch.tensor(x, dtype=torch.float32).to(cuda_device),
                                                LATTICE_PARAMETERS: lattice.to(cuda_device),
                                                ATOM_TYPES: torch.zeros(len(x), 64, dtype=torch.long, device=cuda_device)})
    x32 = x.astype(np.float32).astype(np.float64)  # the kernel sees the fp32-rounded coordinates
    expected = np.array([sw_oracle.energy(xi * np.float64(np.float32(length)), np.full(3, np.float32(length)),
                                          np.zeros(64, dtype=int), ["Si"], entries) for xi in x32])
    np.testing.assert_allclose(energies.cpu().numpy(), expected, rtol=2e-5, atol=2e-4)
    a0 = 2.0 ** (1.0 / 6.0) * 2.0951 * 4.0 / np.sqrt(3.0)
    lattice[:, :3] = 2 * a0
    ideal = oracle_.compute_oracle_energies({AXL_COMPOSITION: AXL(
        A=torch.zeros(1, 64, dtype=torch.long, device=cuda_device),
        X=torch.tensor(diamond(2)[None], dtype=torch.float32, device=cuda_device), L=lattice[:1].to(cuda_device))})
    np.testing.assert_allclose(float(ideal[0]) / 64, -2 * 2.1683, rtol=1e-5)


@pytest.mark.gpu
def test_gpu_silicon_germanium_energies_against_oracle(cuda_device):
    entries = read_stillinger_weber_file(str(SW_COEFFICIENTS_DIR / "SiGe.sw"))
    oracle_ = create_energy_oracle(create_energy_oracle_parameters(dict(name="lammps", sw_coeff_filename="SiGe.sw"),
                                                                   ["Si", "Ge"]))
    rng = np.random.default_rng(5)
    length = 11.084
    x = np.stack([np.mod(diamond(2) + rng.normal(scale=0.02, size=(64, 3)), 1.0) for _ in range(3)] + [rng.random((64, 3))])
    types = rng.integers(0, 2, size=(len(x), 64))
    lattice = torch.zeros(len(x), 6)
    lattice[:, :3] = length
    energies = oracle_.compute_oracle_energies({AXL_COMPOSITION: AXL(
        A=torch.tensor(types, device=cuda_device), X=torch.tensor(x, dtype=torch.float32, device=cuda_device),
        L=lattice.to(cuda_device))})
    x32 = x.astype(np.float32).astype(np.float64)
    expected = np.array([sw_oracle.energy(xi * np.float64(np.float32(length)), np.full(3, np.float32(length)), ti,
                                          ["Ge", "Si"], entries) for xi, ti in zip(x32, types)])
    np.testing.assert_allclose(energies.cpu().numpy(), expected, rtol=2e-5, atol=2e-4)


@pytest.mark.gpu
def test_gpu_small_boxes(cuda_device):
    """A side below 3 Angstrom gives energy 0 like the reference's LAMMPS wrapper; a side inside the cutoff raises."""
    oracle_ = create_energy_oracle(create_energy_oracle_parameters(dict(name="lammps", sw_coeff_filename="Si.sw"), ["Si"]))
    x = torch.rand(2, 8, 3, device=cuda_device)
    a = torch.zeros(2, 8, dtype=torch.long, device=cuda_device)
    lattice = torch.tensor([[2.5, 5.43, 5.43, 0, 0, 0], [5.43, 5.43, 5.43, 0, 0, 0]], device=cuda_device)
    energies = oracle_.compute_oracle_energies({AXL_COMPOSITION: AXL(A=a, X=x, L=lattice)})
    assert float(energies[0]) == 0.0 and float(energies[1]) != 0.0
    lattice[0, 0] = 3.5
    with pytest.raises(AssertionError):
        oracle_.compute_oracle_energies({AXL_COMPOSITION: AXL(A=a, X=x, L=lattice)})

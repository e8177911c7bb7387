"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of the Stillinger-Weber energy the reference obtains
from LAMMPS `pair_style sw` (oracle/lammps_energy_oracle.py:60-193 builds the LAMMPS input; data/
stillinger_weber_coefficients/*.sw hold the parameters).

LAMMPS (`lammps` Python module, the reference's dependency) is NOT installed in this image and its source is not under
/root/reference, so this oracle is **parity unpinned against LAMMPS itself**.  It restates the published functional form
(F. H. Stillinger, T. A. Weber, PRB 31, 5262 (1985); LAMMPS documentation of pair_style sw):

    E = sum_{i<j} phi2(r_ij) + sum_i sum_{j<k} phi3(r_ij, r_ik, theta_jik)
    phi2(r)  = A eps [B (sigma/r)^p - (sigma/r)^q] exp(sigma / (r - a sigma))            parameters of entry (i, j, j)
    phi3     = lambda_ijk eps_ijk (cos theta - cos theta0_ijk)^2
               exp(gamma_ij sigma_ij / (r_ij - a_ij sigma_ij)) exp(gamma_ik sigma_ik / (r_ik - a_ik sigma_ik))

and is pinned by known answers in tests/test_stillinger_weber.py: the cohesive energy of perfect diamond silicon
(-2 eps per atom: the SW parameters were built so), an isolated dimer and an isolated trimer evaluated by hand.
Only tests/ may import this file.
"""
import itertools
from typing import Dict, List, Tuple

import numpy as np


def energy(cartesian_positions: np.ndarray, box: np.ndarray, atom_types: np.ndarray, elements: List[str],
           entries: Dict[Tuple[str, str, str], Dict[str, float]]) -> float:
    """Potential energy (eV) of one configuration in an orthogonal periodic box (images n in {-1,0,1}^3)."""
    positions = np.asarray(cartesian_positions, dtype=np.float64)
    lengths = np.asarray(box, dtype=np.float64)
    n_atoms = len(positions)
    images = np.array(list(itertools.product((-1, 0, 1), repeat=3)), dtype=np.float64) * lengths
    total = 0.0
    for i in range(n_atoms):
        ei = elements[atom_types[i]]
        neighbours = []  # (displacement, r, element)
        for j in range(n_atoms):
            ej = elements[atom_types[j]]
            pair = entries[(ei, ej, ej)]
            cut = pair["a"] * pair["sigma"]
            for shift in images:
                d = positions[j] + shift - positions[i]
                r = np.sqrt(d @ d)
                if 0.0 < r < cut:
                    neighbours.append((d, r, ej))
        for d, r, ej in neighbours:   # pair terms: each bond is seen from both ends
            e = entries[(ei, ej, ej)]
            s = e["sigma"]
            phi2 = e["A"] * e["epsilon"] * (e["B"] * (s / r) ** e["p"] - (s / r) ** e["q"]) * np.exp(s / (r - e["a"] * s))
            total += 0.5 * phi2
        for (d1, r1, ej), (d2, r2, ek) in itertools.combinations(neighbours, 2):
            eij, eik, eijk = entries[(ei, ej, ej)], entries[(ei, ek, ek)], entries[(ei, ej, ek)]
            cos_theta = (d1 @ d2) / (r1 * r2)
            total += (eijk["lambda"] * eijk["epsilon"] * (cos_theta - eijk["costheta0"]) ** 2
                      * np.exp(eij["gamma"] * eij["sigma"] / (r1 - eij["a"] * eij["sigma"]))
                      * np.exp(eik["gamma"] * eik["sigma"] / (r2 - eik["a"] * eik["sigma"])))
    return float(total)

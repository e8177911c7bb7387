"""Integer lattice / Bloch-wave vectors (reference: utils/lattice_utils.py:10-177).

Host-side, evaluated once at model construction.  The Bloch vectors are stored in the score network's
``bloch_wave_reciprocal_lattice_vectors`` parameter (state_dict compatible with the reference).
"""
import itertools
from typing import List

import torch


def get_relative_coordinates_lattice_vectors(number_of_shells: int = 1, spatial_dimension: int = 3) -> torch.Tensor:
    """All integer vectors in [-shells, shells]^d in itertools.product order (reference :10-29)."""
    axis = range(-number_of_shells, number_of_shells + 1)
    return 1.0 * torch.tensor(list(itertools.product(axis, repeat=spatial_dimension)))


def _cubic_point_group(spatial_dimension: int) -> List[torch.Tensor]:
    """Signed permutation matrices (the cubic point group), as integer matrices."""
    ops = []
    for perm in itertools.permutations(range(spatial_dimension)):
        for signs in itertools.product([1, -1], repeat=spatial_dimension):
            m = torch.zeros(spatial_dimension, spatial_dimension, dtype=torch.int64)
            for row, (col, sign) in enumerate(zip(perm, signs)):
                m[row, col] = sign
            ops.append(m)
    return ops


def _order_shell(shell: List[tuple]) -> List[tuple]:
    """Most-positive-first ordering of a symmetric shell (same ordering rule as reference :32-62).

    The rule ranks each component among the distinct values of its column and sorts by the resulting
    mixed-radix number, first component most significant, in decreasing order.
    """
    d = len(shell[0])
    n = len(shell)
    ranks = []
    for comp in range(d):
        values = sorted(set(v[comp] for v in shell))
        ranks.append({value: rank for rank, value in enumerate(values)})
    def key(v):
        return sum(ranks[c][v[c]] * n ** (d - c - 1) for c in range(d))
    keyed = sorted(((key(v), idx) for idx, v in enumerate(shell)))
    # the reference flips an ascending stable argsort, so ties would come out in reverse input order
    return [shell[idx] for _, idx in reversed(keyed)]


def get_cubic_point_group_complete_lattice_shells(
    number_of_complete_shells: int, spatial_dimension: int = 3
) -> List[torch.Tensor]:
    """Symmetric shells of integer vectors by increasing norm, degenerate shells included (reference :65-126)."""
    trial = 2 * number_of_complete_shells
    vectors = [tuple(int(c) for c in v) for v in itertools.product(range(-trial, trial + 1), repeat=spatial_dimension)]
    norms = torch.tensor([sum(c * c for c in v) for v in vectors])
    order = norms.argsort()  # same (default) argsort as the reference on the same key order
    ops = _cubic_point_group(spatial_dimension)
    seen = set()
    shells: List[torch.Tensor] = []
    previous_norm = 0
    for idx in order[1:].tolist():  # skip the zero vector
        v = vectors[idx]
        if v in seen:
            continue
        vt = torch.tensor(v, dtype=torch.int64)
        members = set(tuple(int(c) for c in (op @ vt).tolist()) for op in ops)
        seen.update(members)
        shells.append(torch.tensor(_order_shell(sorted(members)), dtype=torch.int32))
        norm = sum(c * c for c in v)
        if len(shells) >= number_of_complete_shells and norm > previous_norm:
            break
        previous_norm = norm
    return shells


def get_cubic_point_group_positive_normalized_bloch_wave_vectors(
    number_of_complete_shells: int, spatial_dimension: int = 3
) -> torch.Tensor:
    """Half of each shell (one of every +-K pair), stacked (reference :129-177)."""
    halves = []
    for shell in get_cubic_point_group_complete_lattice_shells(number_of_complete_shells, spatial_dimension):
        kept, seen = [], set()
        for vector in shell:
            key = tuple(vector.tolist())
            if key in seen:
                continue
            kept.append(vector)
            seen.add(key)
            seen.add(tuple((-vector).tolist()))
        halves.append(torch.stack(kept))
    return torch.vstack(halves)

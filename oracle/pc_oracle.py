"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's predictor-corrector sampling path.

This file is the *oracle*: a plain torch-on-CPU (fp32 unless the caller passes fp64) restatement of the
algorithm of mila-iqia/diffusion_for_multi_scale_molecular_dynamics for the one path this repository
accelerates.  It is imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs.  Nothing under ``diffusion_for_multi_scale_molecular_dynamics_b200/`` imports it and
it imports nothing from that package.

Parity status: **pinned**.  ``oracle/make_golden.py`` imports the unmodified reference from /root/reference
(through ``oracle/ref_shims.py``) and stores its outputs on seeded inputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those vectors.  The only third-party
arithmetic at the boundary is pykeops 2.2.3 (``uv.lock:2047-2049``), absent from /root/reference; the golden
vectors for the neighbour search were produced with the dense stand-in of ``ref_shims.py``, and the tie
behaviour at ``d^2 == rc^2`` follows that dense fp32 chain (documented in DESIGN.md).

Paths in the citations are relative to ``src/diffusion_for_multi_scale_molecular_dynamics/`` of the reference.
"""
import itertools
import math
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------------
# Noise schedule tables -- noise_schedulers/noise_scheduler.py:112-267, sigma_calculator.py:48-108
# ---------------------------------------------------------------------------------------------------------


def noise_tables(
    total_time_steps: int,
    num_classes: int,
    schedule_type: str = "exponential",
    time_delta: float = 1e-5,
    sigma_min: float = 0.005,
    sigma_max: float = 0.5,
    corrector_step_epsilon: float = 2e-5,
) -> Dict[str, torch.Tensor]:
    """All arrays the sampler reads, indexed exactly like the reference's ``Noise`` / ``LangevinDynamics``."""
    T = total_time_steps
    time = torch.linspace(time_delta, 1.0, T)  # noise_scheduler.py:185-189
    s_min, s_max = torch.tensor(sigma_min), torch.tensor(sigma_max)
    if schedule_type == "exponential":  # sigma_calculator.py:48-78
        sigma = s_min * (s_max / s_min) ** time
    elif schedule_type == "linear":  # sigma_calculator.py:81-108
        sigma = s_min + (s_max - s_min) * time
    else:
        raise NotImplementedError(schedule_type)
    sigma2 = sigma**2
    # g^2_i = sigma^2_i - sigma^2_{i-1}, sigma_0 = sigma_min   (noise_scheduler.py:192-199)
    g2 = torch.cat([torch.tensor([sigma2[0] - sigma_min**2]), sigma2[1:] - sigma2[:-1]])
    g = torch.sqrt(g2)
    # epsilon array, shifted by one (noise_scheduler.py:202-218)
    eps0 = torch.tensor([0.5 * corrector_step_epsilon * sigma_min**2 / sigma2[0]])
    epsilon = torch.cat([eps0, 0.5 * corrector_step_epsilon * sigma2[:-1] / sigma2[0]])
    sqrt_2_epsilon = torch.sqrt(2.0 * epsilon)
    beta = 1.0 / (T - torch.arange(1, T + 1) + 1)  # noise_scheduler.py:220-222
    alpha_bar = torch.cumprod(1 - beta, 0)
    b = beta.unsqueeze(-1).unsqueeze(-1)
    mask_column = F.one_hot(torch.LongTensor([num_classes - 1]), num_classes=num_classes).squeeze(0)
    q = (1 - b) * torch.eye(num_classes)
    q = q + b * torch.outer(torch.ones(num_classes), mask_column)  # noise_scheduler.py:229-242
    q_bar = torch.empty_like(q)
    q_bar[0] = q[0]
    for i in range(1, T):  # noise_scheduler.py:245-253
        q_bar[i] = torch.matmul(q_bar[i - 1], q[i])
    q_bar_tm1 = torch.cat((torch.eye(num_classes).unsqueeze(0), q_bar[:-1]), dim=0)
    return dict(time=time, sigma=sigma, sigma_squared=sigma2, g=g, g_squared=g2, epsilon=epsilon,
                sqrt_2_epsilon=sqrt_2_epsilon, beta=beta, alpha_bar=alpha_bar, q_matrix=q, q_bar_matrix=q_bar,
                q_bar_tm1_matrix=q_bar_tm1)


# ---------------------------------------------------------------------------------------------------------
# Elementwise pieces of the PC step
# ---------------------------------------------------------------------------------------------------------


def wrap_to_unit_cell(x: torch.Tensor) -> torch.Tensor:
    """utils/basis_transformations.py:95-119."""
    r = torch.remainder(x, 1.0)
    r[r == 1.0] = 0.0
    return r


def coordinates_update(x, score, sigma_i, score_weight, noise_weight, z) -> torch.Tensor:
    """generators/langevin_generator.py:194-200: wrap((x + w*s/sigma) + nw*z), evaluated left to right."""
    return wrap_to_unit_cell(x + score_weight * score / sigma_i + noise_weight * z)


def lattice_update(lattice, score, sigma_n_i, score_weight, noise_weight, z) -> torch.Tensor:
    """generators/langevin_generator.py:485-490 (no wrap)."""
    return lattice + score_weight * score / sigma_n_i + noise_weight * z


def probability_from_logits(logits: torch.Tensor, lowest: float) -> torch.Tensor:
    """utils/d3pm_utils.py:127-150."""
    raw = F.softmax(logits, dim=-1)
    total = raw.sum(dim=-1)
    torch.testing.assert_close(total, torch.ones_like(total),
                               msg="Logits are pathological: the probabilities do not sum to one.")
    clipped = raw.clip(min=lowest)
    return clipped / clipped.sum(dim=-1).unsqueeze(-1)


def probability_at_previous_time_step(logits, atom_types, q, q_bar, q_bar_tm1, small_epsilon) -> torch.Tensor:
    """p(a_{t-1} | a_t, gamma_0), utils/d3pm_utils.py:64-124, for one time index ([C,C] matrices)."""
    C = logits.shape[-1]
    gamma = probability_from_logits(logits, small_epsilon)
    onehot = F.one_hot(atom_types.long(), num_classes=C).to(gamma)
    numerator1 = torch.einsum("...j,ji->...i", gamma, q_bar_tm1)
    numerator2 = torch.einsum("ij,...j->...i", q, onehot)
    den1 = torch.einsum("ij,...j->...i", q_bar, onehot)
    den2 = torch.einsum("...j,...j->...", gamma, den1)
    return numerator1 * numerator2 / den2.unsqueeze(-1)


def atom_types_update(logits, atom_types, q, q_bar, q_bar_tm1, gumbel, binary, greedy: bool,
                      one_transition: bool, small_epsilon: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """generators/langevin_generator.py:247-439.  Returns (a_{i-1}, adjusted transition probabilities)."""
    C = logits.shape[-1]
    mask_index = C - 1
    probs = probability_at_previous_time_step(logits, atom_types, q, q_bar, q_bar_tm1, small_epsilon)
    gumbel = gumbel.clone()
    if greedy:  # :382-439
        all_masked = torch.all(atom_types == mask_index, dim=-1)
        atom_is_masked = atom_types == mask_index
        unmask = binary > probs[:, :, -1]
        do_greedy = torch.logical_and(~all_masked.view(-1, 1), unmask)
        do_greedy = torch.logical_and(do_greedy, atom_is_masked)
        probs[:, :, -1] = torch.where(do_greedy, 0, probs[:, :, -1])
        gumbel = torch.where(all_masked.view(-1, 1, 1), gumbel, 0.0)
    max_values, sampled = torch.max(torch.log(probs + small_epsilon) + gumbel, dim=-1)  # :311-315
    if one_transition:  # :339-380
        n = atom_types.shape[0]
        changed = sampled != atom_types
        restricted = torch.where(changed, max_values, -torch.inf)
        winner = torch.argmax(restricted, dim=-1)
        updated = atom_types.clone()
        rows = torch.arange(n)
        updated[rows, winner] = sampled[rows, winner]
    else:
        updated = sampled
    return updated, probs


# ---------------------------------------------------------------------------------------------------------
# Periodic radius graph -- utils/neighbors.py:36-224, models/graph_utils.py:10-50, models/egnn_utils.py:73-144
# ---------------------------------------------------------------------------------------------------------


def relative_lattice_vectors() -> torch.Tensor:
    """The 27 images n in {-1,0,1}^3 in itertools.product order (utils/lattice_utils.py:10-29)."""
    shifts = range(-1, 2)
    return torch.tensor(list(itertools.product(shifts, repeat=3)), dtype=torch.float32)


def shortest_cell_crossing_distance(basis: torch.Tensor) -> torch.Tensor:
    """utils/neighbors.py:310-351 (3D): volume / area of each pair of cell vectors, minimum."""
    a1, a2, a3 = basis[:, 0], basis[:, 1], basis[:, 2]
    c12, c13, c23 = torch.linalg.cross(a1, a2), torch.linalg.cross(a1, a3), torch.linalg.cross(a2, a3)
    volume = (c12 * a3).sum(dim=1).abs()
    distances = torch.stack([volume / torch.linalg.norm(c, dim=1) for c in (c12, c13, c23)], dim=1)
    return distances.min(dim=1).values


def _fp32_chain_matmul3(x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """x[..., 3] @ m[3, 3] as ((x0 m0 + x1 m1) + x2 m2): the explicit, contraction-free fp32 chain."""
    return (x[..., 0:1] * m[..., 0, :] + x[..., 1:2] * m[..., 1, :]) + x[..., 2:3] * m[..., 2, :]


def periodic_adjacency(cartesian_positions: torch.Tensor, basis_vectors: torch.Tensor, radial_cutoff: float):
    """get_periodic_adjacency_information (utils/neighbors.py:36-224), dense semantics.

    Returns (adjacency [2,E] local indices, shifts [E,3], edge_batch [E], number_of_edges [B]) in the
    reference's native order: (batch, image, source, ascending d^2 -- ties by destination index).
    """
    B, N, _ = cartesian_positions.shape
    rc = torch.tensor(radial_cutoff, dtype=cartesian_positions.dtype)
    assert torch.all(shortest_cell_crossing_distance(basis_vectors) > rc), (
        "The radial cutoff is so large that neighbors could be located beyond the first shell of periodic unit "
        "cell images.")
    images = relative_lattice_vectors().to(cartesian_positions.dtype)  # [27,3]
    lattice_vectors = torch.stack([_fp32_chain_matmul3(images, basis_vectors[b]) for b in range(B)])  # [B,27,3]
    shifted = cartesian_positions[:, None, :, :] + lattice_vectors[:, :, None, :]  # [B,27,N(j),3]  :242-244
    diff = cartesian_positions[:, None, :, None, :] - shifted[:, :, None, :, :]  # [B,27,N(i),N(j),3]
    sq = diff * diff
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]  # :170
    valid = torch.logical_and(d2 > 0.0, d2 <= rc**2)  # :188-190
    src_l, dst_l, shift_l, batch_l = [], [], [], []
    counts = torch.zeros(B, dtype=torch.int64)
    for b in range(B):
        for img in range(27):
            for i in range(N):
                js = torch.nonzero(valid[b, img, i]).flatten()
                if js.numel() == 0:
                    continue
                order = sorted(js.tolist(), key=lambda j: (float(d2[b, img, i, j]), j))
                for j in order:
                    src_l.append(i); dst_l.append(j); batch_l.append(b)
                    shift_l.append(lattice_vectors[b, img])
                counts[b] += len(order)
    adjacency = torch.tensor([src_l, dst_l], dtype=torch.int64).reshape(2, -1)
    shifts = torch.stack(shift_l) if shift_l else torch.zeros(0, 3, dtype=cartesian_positions.dtype)
    return adjacency, shifts, torch.tensor(batch_l, dtype=torch.int64), counts


def compute_distances_in_batch(cartesian_positions: torch.Tensor, unit_cell: torch.Tensor, max_distance: float,
                               ieee_sqrt: bool = False):
    """compute_distances_in_batch (utils/structure_utils.py:41-121), dense semantics: all distances 0 < d <= max_distance
    over the 27 images, flattened in (batch, image, source, ascending d^2 -- ties by destination) order.

    ``torch.sqrt`` on CPU is a vectorised approximation that differs from the correctly rounded square root in ~0.6 % of
    fp32 inputs (by one ulp); ``ieee_sqrt=True`` takes numpy's correctly rounded one instead, which is what CUDA's
    ``sqrtf`` (and therefore the reference run on a GPU) returns."""
    B, N, _ = cartesian_positions.shape
    rc = torch.tensor(max_distance, dtype=cartesian_positions.dtype)
    images = relative_lattice_vectors().to(cartesian_positions.dtype)
    lattice_vectors = torch.stack([_fp32_chain_matmul3(images, unit_cell[b]) for b in range(B)])  # [B,27,3]
    shifted = cartesian_positions[:, None, :, :] + lattice_vectors[:, :, None, :]
    diff = cartesian_positions[:, None, :, None, :] - shifted[:, :, None, :, :]  # [B,27,N(i),N(j),3]
    sq = diff * diff
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    order = torch.argsort(d2, dim=-1, stable=True)  # Kmin_argKmin: ascending d^2
    sorted_d2 = torch.gather(d2, -1, order)
    distances = (torch.from_numpy(np.sqrt(sorted_d2.numpy())) if ieee_sqrt else torch.sqrt(sorted_d2)).flatten()  # :110
    return distances[torch.logical_and(distances > 0.0, distances <= rc)]  # :113-121


def repulsive_pseudo_force(x_rel: torch.Tensor, lattice: torch.Tensor, radial_cutoff: float, strength: float,
                           min_box_size: float = 1.0) -> torch.Tensor:
    """ForceFieldAugmentedScoreNetwork.get_relative_coordinates_pseudo_force
    (models/score_networks/force_field_augmented_score_network.py:120-236), dense semantics."""
    B, N, _ = x_rel.shape
    lengths = lattice.clip(min=min_box_size)[:, :3]          # basis_transformations.py:253-256
    basis = torch.diag_embed(lengths)
    cart = torch.matmul(x_rel, basis)
    adjacency, shifts, edge_batch, _ = periodic_adjacency(cart, basis, radial_cutoff)
    src, dst = adjacency
    displacements = cart[edge_batch, dst] - cart[edge_batch, src] + shifts           # :150-154
    r = torch.linalg.norm(displacements, dim=1)
    prefactors = 2.0 * strength * (r - radial_cutoff) / (r + torch.tensor(1.0e-8))  # :104-108
    contributions = prefactors[:, None] * displacements
    forces = torch.zeros(B * N, 3)
    forces.index_add_(0, N * edge_batch + src, contributions)                        # :183-199
    return torch.matmul(forces.reshape(B, N, 3), torch.inverse(basis))              # :226-234


def periodic_pair_mask(relative_coordinates: torch.Tensor, unit_cell: torch.Tensor, radial_cutoff: float):
    """[B,N,N] bool: at least one of the 27 images of j lies within the cutoff of i (vectorised, for big cases)."""
    cart = torch.stack([_fp32_chain_matmul3(relative_coordinates[b], unit_cell[b])
                        for b in range(relative_coordinates.shape[0])])
    rc = torch.tensor(radial_cutoff, dtype=cart.dtype)
    assert torch.all(shortest_cell_crossing_distance(unit_cell) > rc), (
        "The radial cutoff is so large that neighbors could be located beyond the first shell of periodic unit "
        "cell images.")
    images = relative_lattice_vectors().to(cart.dtype)
    hit = torch.zeros(cart.shape[0], cart.shape[1], cart.shape[1], dtype=torch.bool)
    for img in range(27):
        lat = torch.stack([_fp32_chain_matmul3(images[img], unit_cell[b]) for b in range(cart.shape[0])])  # [B,3]
        shifted = cart + lat[:, None, :]
        diff = cart[:, :, None, :] - shifted[:, None, :, :]
        sq = diff * diff
        d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        hit |= torch.logical_and(d2 > 0.0, d2 <= rc**2)
    return hit


def edges_with_radial_cutoff(relative_coordinates, unit_cell, radial_cutoff: float) -> torch.Tensor:
    """get_edges_with_radial_cutoff with drop_duplicate_edges=True (models/egnn_utils.py:107-144):
    graph-batched (src, dst) pairs, sorted lexicographically, duplicates across images removed -> [E,2]."""
    B, N, _ = relative_coordinates.shape
    hit = periodic_pair_mask(relative_coordinates, unit_cell, radial_cutoff)
    idx = torch.nonzero(hit)  # sorted by (b, i, j)
    offset = idx[:, 0] * N
    return torch.stack([idx[:, 1] + offset, idx[:, 2] + offset], dim=1)


def fully_connected_edges(n_nodes: int, batch_size: int) -> torch.Tensor:
    """get_edges_batch (models/egnn_utils.py:73-104): order (b, x, y != x)."""
    pairs = [(x, y) for x in range(n_nodes) for y in range(n_nodes) if x != y]
    base = torch.tensor(pairs, dtype=torch.int64).reshape(-1, 2)   # (0, 2) for a single atom
    return torch.cat([base + n_nodes * b for b in range(batch_size)])


# ---------------------------------------------------------------------------------------------------------
# EGNN score network -- models/score_networks/egnn_score_network.py:131-303, models/egnn.py, egnn_utils.py
# ---------------------------------------------------------------------------------------------------------


def segment_sum(data: torch.Tensor, ids: torch.Tensor, n: int) -> torch.Tensor:
    """models/egnn_utils.py:11-38."""
    out = data.new_zeros(n, data.shape[1])
    return out.scatter_add_(0, ids.unsqueeze(-1).expand(-1, data.shape[1]), data)


def segment_mean(data: torch.Tensor, ids: torch.Tensor, n: int) -> torch.Tensor:
    """models/egnn_utils.py:41-70 (count clamped to 1)."""
    total = segment_sum(data, ids, n)
    count = segment_sum(torch.ones_like(data), ids, n)
    return total / count.clamp(min=1)


def _mlp(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, n_linear: int, last_activation: bool,
         last_bias: bool = True) -> torch.Tensor:
    """Sequential(Linear, SiLU, Linear, SiLU, ...): Linear modules sit at even indices (models/egnn.py:78-134)."""
    for k in range(n_linear):
        w = sd[f"{prefix}.{2 * k}.weight"]
        bias = sd.get(f"{prefix}.{2 * k}.bias") if (k < n_linear - 1 or last_bias) else None
        x = F.linear(x, w, bias)
        if k < n_linear - 1 or last_activation:
            x = F.silu(x)
    return x


def egnn_score_forward(sd: Dict[str, torch.Tensor], hp: Dict, atom_types, x_rel, lattice, sigma,
                       edges: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """ScoreNetwork.forward of EGNNScoreNetwork: (logits [B,N,C] with MASK = -inf, sigma-normalised score [B,N,3]).

    ``sd`` is a reference ``state_dict``; ``hp`` holds the EGNNScoreNetworkParameters fields that are not
    recoverable from tensor shapes (n_layers, *_n_hidden_dimensions, residual, attention, normalize, tanh,
    coords_agg, message_agg, edges, radial_cutoff, num_atom_types).
    """
    B, N, d = x_rel.shape
    num_atom_types = hp["num_atom_types"]
    C = num_atom_types + 1
    if edges is None:
        if hp.get("edges", "fully_connected") == "fully_connected":
            edges = fully_connected_edges(N, B)
        else:  # egnn_score_network.py:236-247
            rc = hp["radial_cutoff"]
            lp = lattice.clip(min=2.2 * rc)
            lp[:, d:] = 0
            unit_cell = torch.diag_embed(lp[:, :d])
            edges = edges_with_radial_cutoff(x_rel, unit_cell, rc)
    flat_x = x_rel.reshape(B * N, d)
    # uplift (egnn_score_network.py:196-224)
    two_pi_x = 2.0 * torch.pi * flat_x
    bloch = sd["bloch_wave_reciprocal_lattice_vectors"].to(two_pi_x)
    kr = torch.einsum("ks,ns->nk", bloch, two_pi_x)
    z = torch.stack([kr.cos(), kr.sin()], dim=-1).reshape(B * N, -1)  # nodes (nbloch two)
    # node attributes (egnn_score_network.py:165-194)
    sig = sigma.reshape(B, 1).to(flat_x).repeat_interleave(N, dim=0)
    h = torch.cat([sig, F.one_hot(atom_types.reshape(-1).long(), num_classes=C).to(flat_x)], dim=1)
    # EGNN (models/egnn.py:362-385)
    h = F.linear(h, sd["egnn.embedding_in.weight"], sd["egnn.embedding_in.bias"])
    coord = z.clone()
    row, col = edges[:, 0], edges[:, 1]
    n_nodes = B * N
    coords_agg = segment_sum if hp.get("coords_agg", "mean") == "sum" else segment_mean
    msg_agg = segment_sum if hp.get("message_agg", "mean") == "sum" else segment_mean
    for layer in range(hp.get("n_layers", 4)):
        p = f"egnn.graph_layers.{layer}"
        diff = coord[row] - coord[col]  # models/egnn.py:230-238
        radial = torch.sum(diff**2, 1).unsqueeze(1)
        if hp.get("normalize", False):  # :240-264, epsilon = 1e-8
            diff = torch.tanh(radial) / torch.sqrt(radial + 1e-8**2) * diff
        m = _mlp(sd, f"{p}.message_mlp", torch.cat([h[row], h[col], radial], dim=1),
                 1 + hp.get("message_n_hidden_dimensions", 1), True)  # :136-160
        if hp.get("attention", False):
            m = m * torch.sigmoid(F.linear(m, sd[f"{p}.att_mlp.0.weight"], sd[f"{p}.att_mlp.0.bias"]))
        phi_x = _mlp(sd, f"{p}.coord_mlp", m, 2 + hp.get("coordinate_n_hidden_dimensions", 1), False,
                     last_bias=False)  # :110-129
        if hp.get("tanh", False):
            phi_x = torch.tanh(phi_x)
        coord = coord + coords_agg(diff * phi_x, row, n_nodes)  # :189-215
        agg = torch.cat([h, msg_agg(m, row, n_nodes)], dim=1)  # :162-187
        out = _mlp(sd, f"{p}.node_mlp", agg, 2 + hp.get("node_n_hidden_dimensions", 1), False)
        h = h + out if hp.get("residual", True) else out
    logits = F.linear(h, sd["egnn.node_classification_layer.weight"], sd["egnn.node_classification_layer.bias"])
    # projection S^alpha = z . Gamma^alpha . zhat (egnn_score_network.py:276-281)
    gamma = sd["projection_matrices"].to(z)
    score = torch.einsum("ni,aij,nj->na", z, gamma, coord)
    logits = logits.reshape(B, N, C).clone()
    logits[..., num_atom_types] = -torch.inf  # models/score_networks/score_network.py:183-185
    return logits, score.reshape(B, N, d)


# ---------------------------------------------------------------------------------------------------------
# Wrapped-Gaussian sigma-normalised score -- score/wrapped_gaussian_score.py:131-419
# ---------------------------------------------------------------------------------------------------------

SIGMA_THRESHOLD = torch.Tensor([1.0 / math.sqrt(2.0 * math.pi)])  # float32 tensor, as in the reference (:37)
U_THRESHOLD = torch.Tensor([0.5])


def wrapped_gaussian_score(u: torch.Tensor, sigma: torch.Tensor, kmax: int) -> torch.Tensor:
    """get_coordinates_sigma_normalized_score, all three branches, any dtype."""
    flat_u, flat_s = u.reshape(-1), sigma.reshape(-1)
    k = torch.arange(-kmax, kmax + 1)
    out = torch.zeros_like(flat_u)
    small_sigma = flat_s <= SIGMA_THRESHOLD
    m1a = torch.logical_and(small_sigma, flat_u < U_THRESHOLD)
    m1b = torch.logical_and(small_sigma, flat_u >= U_THRESHOLD)
    m2 = flat_s > SIGMA_THRESHOLD
    for mask, branch in ((m1a, 0), (m1b, 1), (m2, 2)):
        if not mask.any():
            continue
        cu, cs = flat_u[mask].unsqueeze(1), flat_s[mask].unsqueeze(1)
        if branch == 0:  # :255-274
            e = torch.exp(-0.5 * (k**2 + 2.0 * cu * k) / cs**2)
        elif branch == 1:  # :277-298
            e = torch.exp(-0.5 * ((k**2 - 1.0) + 2.0 * cu * (k + 1.0)) / cs**2)
        if branch < 2:  # :301-367
            num = ((torch.ones_like(cu) * k) * e).sum(dim=1)
            out[mask] = (-flat_u[mask] - num / e.sum(dim=1)) / flat_s[mask]
        else:  # :369-419
            pi = torch.tensor(math.pi, dtype=u.dtype)
            upk, gu, sg = cu + k, cu * k, cs * k
            g = torch.ones_like(cu) * k
            sig = cs * torch.ones_like(k)
            exp_upk = (-pi * upk**2).exp()
            comb = torch.sqrt(2.0 * pi) * sig * (-2.0 * pi**2 * sg**2).exp() - (-pi * g**2).exp()
            z2 = exp_upk.sum(dim=1) + (comb * torch.cos(2.0 * pi * gu)).sum(dim=1)
            dz2 = -2.0 * pi * ((upk * exp_upk).sum(dim=1) + (g * comb * torch.sin(2.0 * pi * gu)).sum(dim=1))
            out[mask] = flat_s[mask] * dz2 / z2
    return out.reshape(u.shape)


def analytical_score(x: torch.Tensor, x0: torch.Tensor, sigma_t: torch.Tensor, sigma_d: float, kmax: int) -> torch.Tensor:
    """AnalyticalScoreNetwork without permutation symmetry (models/score_networks/analytical_score_network.py:134-298):
    sigma_t * S_wrapped(wrap(x - x0), sigma_eff) / sigma_eff with sigma_eff = sqrt(sigma_d^2 + sigma_t^2)."""
    sig = sigma_t.reshape(-1, 1, 1).to(x).expand_as(x)
    effective = torch.sqrt(sigma_d**2 + sig**2)
    u = wrap_to_unit_cell(x - x0.unsqueeze(0))
    return sig * (wrapped_gaussian_score(u, effective, kmax) / effective)


# ---------------------------------------------------------------------------------------------------------
# Noise tape: deterministic noise shared by the reference, the oracle and the CUDA path in parity tests
# ---------------------------------------------------------------------------------------------------------


class NoiseTape:
    """One independent, seeded stream per draw kind; each call returns the next tensor of that stream.

    The reference draws from the global CPU RNG in a fixed order (generators/langevin_generator.py:92-111);
    parity tests patch the ``_draw_*`` hooks of every implementation with one tape each, built from the
    same seed, exactly as the reference's own tests inject noise (tests/generators/test_langevin_generator.py).
    """

    KINDS = ("coordinates", "lattice", "gumbel_uniform", "binary", "init_uniform", "repaint_normal",
             "repaint_uniform")

    def __init__(self, seed: int):
        self.generators = {}
        for i, kind in enumerate(self.KINDS):
            gen = torch.Generator(device="cpu")
            gen.manual_seed(seed * 1000 + i)
            self.generators[kind] = gen

    def normal(self, kind: str, *shape) -> torch.Tensor:
        return torch.randn(*shape, generator=self.generators[kind])

    def uniform(self, kind: str, *shape) -> torch.Tensor:
        return torch.rand(*shape, generator=self.generators[kind])

    def gumbel(self, small_epsilon: float, *shape) -> torch.Tensor:
        """generators/langevin_generator.py:100-107."""
        return -torch.log(-torch.log(self.uniform("gumbel_uniform", *shape).clip(min=small_epsilon)))


# ---------------------------------------------------------------------------------------------------------
# The PC loop -- generators/predictor_corrector_axl_generator.py:81-161, langevin_generator.py:536-805,
# constrained_langevin_generator.py:74-182
# ---------------------------------------------------------------------------------------------------------

ScoreFn = Callable[[torch.Tensor, torch.Tensor, torch.Tensor, float], Tuple[torch.Tensor, torch.Tensor]]


def pc_sample(
    score_fn: ScoreFn,
    tables: Dict[str, torch.Tensor],
    atom_types: torch.Tensor,
    x: torch.Tensor,
    lattice: torch.Tensor,
    tape: NoiseTape,
    number_of_corrector_steps: int = 1,
    sigma_min: float = 0.005,
    small_epsilon: float = 1e-8,
    greedy: bool = True,
    one_transition: bool = True,
    atom_type_transition_in_corrector: bool = False,
    starting_step_index: Optional[int] = None,
    ending_step_index: int = 0,
    constraint: Optional[Dict[str, torch.Tensor]] = None,
    record: Optional[List] = None,
    free_lattice: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Predictor-corrector trajectory.  ``score_fn(A, X, L, sigma) -> (logits, score)``.

    ``constraint`` = dict(x0 [Nc,3], a0 [Nc], indices [Nc]) switches on RePaint.  Noise comes from ``tape`` in the
    reference's per-kind order.  With a fixed lattice the lattice draws are skipped: they do not influence the result
    (langevin_generator.py:475-477).  ``free_lattice`` (use_fixed_lattice_parameters=False) diffuses the lattice
    parameters too; the lattice score of the EGNN is identically zero (egnn_score_network.py:296-300), so
    l <- l + g z in the predictor (:485-490,636-641) and l <- l + sqrt(2 eps) z' in the corrector, where z' is the
    SECOND lattice draw of the step: the first one only feeds the step size (:758-777).
    """
    T = tables["sigma"].shape[0]
    B, N, d = x.shape
    C = tables["q_matrix"].shape[-1]
    start = T if starting_step_index is None else starting_step_index
    a_cur, x_cur = atom_types.clone(), x.clone()
    lattice = lattice.clone() if free_lattice else lattice

    def repaint(a_in, x_in, index_i):  # constrained_langevin_generator.py:136-163
        idx = constraint["indices"]
        nc = idx.shape[0]
        tape.uniform("init_uniform", B, N, d)  # the reference re-initialises a throw-away composition (:84-92)
        if index_i == 0:
            x_known = constraint["x0"].expand(B, nc, d)
            a_known = constraint["a0"].expand(B, nc)
        else:  # data/diffusion/noising_transform.py:98-200 at time index index_i (idx = index_i - 1)
            sig = tables["sigma"][index_i - 1]
            z = tape.normal("repaint_normal", B, N, d)[:, idx]
            x_known = wrap_to_unit_cell(constraint["x0"].expand(B, nc, d) + sig * z)
            u = tape.uniform("repaint_uniform", B, N, C)[:, idx]
            onehot = F.one_hot(constraint["a0"].long(), num_classes=C).to(x_in)
            post = torch.einsum("nj,ji->ni", onehot, tables["q_bar_matrix"][index_i - 1])  # q(a_t | a_0)
            noisy = torch.log(post).expand(B, nc, C) + (-torch.log(-torch.log(u)))
            a_known = torch.argmax(noisy, dim=-1)
        x_in[:, idx] = x_known
        a_in[:, idx] = a_known
        return a_in, x_in

    for i in range(start - 1, max(ending_step_index, 0) - 1, -1):
        # predictor step, index_i = i + 1, idx = i   (langevin_generator.py:536-645)
        idx = i
        sigma_i = tables["sigma"][idx]
        logits, score = score_fn(a_cur, x_cur, lattice, float(sigma_i))
        last = idx == 0
        gumbel = tape.gumbel(small_epsilon, B, N, C)
        binary = tape.uniform("binary", B, N) if greedy else None
        a_new, _ = atom_types_update(logits, a_cur, tables["q_matrix"][idx], tables["q_bar_matrix"][idx],
                                     tables["q_bar_tm1_matrix"][idx], gumbel, binary, greedy,
                                     one_transition and not last, small_epsilon)
        if last:
            assert (a_new != C - 1).all(), "There remains MASKED atoms at the last time step"
        z = tape.normal("coordinates", B, N, d)
        x_cur = coordinates_update(x_cur, score, sigma_i, tables["g_squared"][idx], tables["g"][idx], z)
        if free_lattice:
            lattice = lattice + tables["g"][idx] * tape.normal("lattice", B, lattice.shape[1])
        a_cur = a_new
        if constraint is not None:
            a_cur, x_cur = repaint(a_cur, x_cur, i)
        if record is not None:
            record.append(("predictor", i + 1, a_cur.clone(), x_cur.clone()))
        for _ in range(number_of_corrector_steps):  # corrector step, index_i = i   (:693-805)
            if i == 0:
                sigma_c, idx_c = torch.tensor(sigma_min), 0
            else:
                idx_c = i - 1
                sigma_c = tables["sigma"][idx_c]
            logits, score = score_fn(a_cur, x_cur, lattice, float(sigma_c))
            z = tape.normal("coordinates", B, N, d)
            eps_i = tables["epsilon"][i]
            x_cur = coordinates_update(x_cur, score, sigma_c, eps_i, torch.sqrt(2 * eps_i), z)
            if free_lattice:
                tape.normal("lattice", B, lattice.shape[1])  # z_lattice: consumed by the step-size hook only
                lattice = lattice + torch.sqrt(2 * eps_i) * tape.normal("lattice", B, lattice.shape[1])
            if atom_type_transition_in_corrector:
                gumbel = tape.gumbel(small_epsilon, B, N, C)
                binary = tape.uniform("binary", B, N) if greedy else None
                a_cur, _ = atom_types_update(logits, a_cur, tables["q_matrix"][idx_c], tables["q_bar_matrix"][idx_c],
                                             tables["q_bar_tm1_matrix"][idx_c], gumbel, binary, greedy, one_transition,
                                             small_epsilon)
            if record is not None:
                record.append(("corrector", i, a_cur.clone(), x_cur.clone()))
    if constraint is not None:  # final hard constraint (constrained_langevin_generator.py:179-181)
        x_cur[:, constraint["indices"]] = constraint["x0"]
        a_cur[:, constraint["indices"]] = constraint["a0"]
    return a_cur, x_cur, lattice


def pair_distance_histogram(x_rel: torch.Tensor, cell_lengths: torch.Tensor, r_max: float, bins: int) -> torch.Tensor:
    """g(r)-style statistic: histogram of all pair distances 0 < d <= r_max over the 27 images
    (the definition of utils/structure_utils.py:41-121), orthogonal cells."""
    cart = x_rel * cell_lengths[:, None, :]
    images = relative_lattice_vectors().to(cart)
    hist = torch.zeros(bins)
    for img in range(27):
        lat = images[img][None, :] * cell_lengths  # [B,3]
        diff = cart[:, :, None, :] - (cart[:, None, :, :] + lat[:, None, None, :])
        dist = torch.sqrt((diff**2).sum(-1))
        sel = dist[(dist > 0) & (dist <= r_max)]
        hist += torch.histc(sel, bins=bins, min=0.0, max=r_max)
    return hist


def stillinger_weber_two_body_energy(x_rel: torch.Tensor, cell_lengths: torch.Tensor) -> torch.Tensor:
    """Energy proxy per structure: the two-body term of the Stillinger-Weber Si potential
    (data/stillinger_weber_coefficients/Si.sw: epsilon 2.1683 eV, sigma 2.0951 A, a 1.80, A 7.049556277, B 0.6022245584,
    p 4, q 0) summed over all periodic pairs within a*sigma.  The reference evaluates energies through LAMMPS
    (oracle/lammps.py:19-101), which is not available here; the pair term alone is a sufficient *statistic* to compare two
    ensembles of sampled structures."""
    eps, sig, a, A, B = 2.1683, 2.0951, 1.80, 7.049556277, 0.6022245584
    cart = x_rel * cell_lengths[:, None, :]
    images = relative_lattice_vectors().to(cart)
    energy = torch.zeros(x_rel.shape[0], dtype=cart.dtype)
    for img in range(27):
        lat = images[img][None, :] * cell_lengths
        diff = cart[:, :, None, :] - (cart[:, None, :, :] + lat[:, None, None, :])
        r = torch.sqrt((diff**2).sum(-1)) / sig
        inside = (r > 0.05) & (r < a - 1e-6)   # r > 0.05: skip self pairs / numerically coincident atoms
        rr = torch.where(inside, r, torch.ones_like(r))
        phi = A * eps * (B * rr**-4 - 1.0) * torch.exp(1.0 / (rr - a).clamp(max=-1e-6))
        energy += 0.5 * torch.where(inside, phi, torch.zeros_like(phi)).sum(dim=(1, 2))
    return energy

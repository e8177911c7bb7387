#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels of the path (graph / score / update / statistics), each launched
alone at a size larger than L2 and timed with CUDA events (best of 10 after 3 warm-ups), against the measured copy
bandwidth of MEASURED_PEAKS.json.  Bytes are ALGORITHMIC.

    python scripts/kernel_rooflines.py > gpurun_out/kernel_rooflines.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.models.egnn_utils import get_edges_with_radial_cutoff  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.force_field_augmented_score_network import (  # noqa: E402
    ForceFieldAugmentedScoreNetwork, ForceFieldParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, AXL_COMPOSITION,  # noqa: E402
                                                                       NOISY_AXL_COMPOSITION)
from diffusion_for_multi_scale_molecular_dynamics_b200.oracle.energy_oracle_factory import (  # noqa: E402
    create_energy_oracle, create_energy_oracle_parameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.basis_transformations import \
    map_relative_coordinates_to_unit_cell  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.structure_utils import compute_distances_in_batch  # noqa: E402

dev = torch.device("cuda:0")
lib = _cabi.lib()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
PEAK = peaks["hbm_gbs"]


def timed(fn, repeats=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(repeats):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        fn()
        stop.record()
        torch.cuda.synchronize()
        best = min(best, start.elapsed_time(stop))
    return best


rows = []


def report(name, ms, algorithmic_bytes, note):
    gbs = algorithmic_bytes / (ms * 1e-3) / 1e9
    rows.append({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(algorithmic_bytes / 1e6, 2),
                 "achieved_GBps": round(gbs, 1), "peak_GBps": PEAK, "frac": round(gbs / PEAK, 3), "note": note})


stream = _cabi.stream_ptr(dev)
# ---- predictor / corrector coordinate update: 36 B / atom (noise drawn in registers) ---------------------------------
B, N = 32768, 512
x = torch.rand(B, N, 3, device=dev)
s = torch.randn(B, N, 3, device=dev)
out = torch.empty_like(x)
ms = timed(lambda: _cabi.check(lib.b200md_pc_coordinates_update(_cabi.ptr(x), _cabi.ptr(s), None, 0.01, 0.1, 0.05, B, N, 3,
                                                                1234, 7, 0, _cabi.ptr(out), stream)))
report("pc_coordinates_update_kernel<3> (in-kernel Philox noise)", ms, 36 * B * N, f"{B} x {N} atoms")
# ---- wrap --------------------------------------------------------------------------------------------------------
ms = timed(lambda: map_relative_coordinates_to_unit_cell(s))
report("wrap_kernel", ms, 8 * s.numel(), f"{s.numel()} values (includes the torch.empty_like of the wrapper)")
# ---- wrapped Gaussian score: 12 B / element ------------------------------------------------------------------------
u = torch.rand(64 * 1024 * 1024, device=dev)
sig = torch.rand_like(u) * 0.5 + 0.001
wg_out = torch.empty_like(u)
ms = timed(lambda: _cabi.check(lib.b200md_wrapped_gaussian_score_f32(_cabi.ptr(u), _cabi.ptr(sig), u.numel(), 4,
                                                                     _cabi.ptr(wg_out), stream)))
report("wrapped_gaussian_score_kernel<f32>, kmax 4", ms, 12 * u.numel(), f"{u.numel()} elements, sigma in (0.001, 0.5)")
# the shape the reference calls it with: one sigma per SAMPLE broadcast over its atoms x 3 coordinates (64 atoms: runs of 192)
sig = (torch.rand(u.numel() // 192, 1, device=dev) * 0.5 + 0.001).expand(-1, 192).reshape(-1).contiguous()
ms = timed(lambda: _cabi.check(lib.b200md_wrapped_gaussian_score_f32(_cabi.ptr(u), _cabi.ptr(sig), u.numel(), 4,
                                                                     _cabi.ptr(wg_out), stream)))
report("wrapped_gaussian_score_kernel<f32>, kmax 4, sigma per sample", ms, 12 * u.numel(),
       f"{u.numel()} elements, sigma in (0.001, 0.5) constant over runs of 192 elements (one 64-atom sample): one branch per warp")
del u, sig, wg_out, x, s, out
# ---- atom-type update: 16 + 4C B / atom ------------------------------------------------------------------------------
B, N, C = 32768, 512, 3
logits = torch.randn(B, N, C, device=dev)
a = torch.full((B, N), C - 1, dtype=torch.long, device=dev)
q = torch.eye(C, device=dev) * 0.9 + 0.1 / C
a_out = torch.empty_like(a)
status = torch.zeros(1, dtype=torch.int32, device=dev)
ms = timed(lambda: _cabi.check(lib.b200md_atom_types_update(
    _cabi.ptr(logits), _cabi.ptr(a), _cabi.ptr(q), _cabi.ptr(q), _cabi.ptr(q), None, None, 3, 1e-8, B, N, C, 1234, 7, 0,
    _cabi.ptr(a_out), None, _cabi.ptr(status), stream)))
report("atom_types_update_kernel (greedy + one transition, in-kernel Gumbel)", ms, (16 + 4 * C) * B * N, f"{B} x {N} atoms, C = {C}")
del logits, a, a_out
# ---- radius graph of the EGNN (sorted unique pairs, int64 API form) --------------------------------------------------
for B, N, length, rc in ((8192, 64, 16.5, 7.5), (1024, 512, 21.72, 7.5)):
    rel = torch.rand(B, N, 3, device=dev)
    cell = torch.diag_embed(torch.full((B, 3), length, device=dev))
    edges = get_edges_with_radial_cutoff(rel, cell, rc)
    ms = timed(lambda: get_edges_with_radial_cutoff(rel, cell, rc), repeats=5)
    E = edges.shape[0]
    report(f"radius graph count + scan + fill (get_edges_with_radial_cutoff), N = {N}", ms, B * (12 * N + 36) + 16 * E,
           f"{B} structures, {E / B:.0f} edges/structure; FP32-issue-bound sweep of N^2 pairs x the images that survive "
           "the culling, includes the host-side cumsum and one .item()")
# ---- general periodic adjacency (per-image edges, shifts): the Stillinger-Weber cutoff in the 4x4x4 supercell --------------
from diffusion_for_multi_scale_molecular_dynamics_b200.utils.neighbors import get_periodic_adjacency_information  # noqa: E402
for B, N, length, rc in ((1024, 512, 21.72, 3.77), (8192, 64, 10.86, 3.77)):
    cart = torch.rand(B, N, 3, device=dev) * length
    cell = torch.diag_embed(torch.full((B, 3), length, device=dev))
    info = get_periodic_adjacency_information(cart, cell, rc)
    ms = timed(lambda: get_periodic_adjacency_information(cart, cell, rc), repeats=5)
    E = info.adjacency_matrix.shape[1]
    report(f"periodic adjacency count + fill (get_periodic_adjacency_information), N = {N}, rc = {rc}", ms,
           B * (12 * N + 36) + 36 * E, f"{B} structures, {E / B:.0f} per-image edges/structure (int64 pairs + shifts + batch "
           f"index); L / rc = {length / rc:.1f}: image culling leaves (1 + 2 rc / L)^3 = {(1 + 2 * rc / length) ** 3:.1f} of 27 "
           "images per atom; includes the host-side cumsum and one .item()")
# ---- pair distances ----------------------------------------------------------------------------------------------------
B, N, length = 1024, 512, 21.72
cart = torch.rand(B, N, 3, device=dev) * length
cell = torch.diag_embed(torch.full((B, 3), length, device=dev))
d = compute_distances_in_batch(cart, cell, 5.0)
ms = timed(lambda: compute_distances_in_batch(cart, cell, 5.0), repeats=5)
report("pair_distances count + fill (compute_distances_in_batch), N = 512", ms, B * (12 * N + 36) + 4 * d.numel(),
       f"{B} x {N} atoms, {d.numel() / B:.0f} distances/structure")
B, N, length = 8192, 64, 10.86
cart = torch.rand(B, N, 3, device=dev) * length
cell = torch.diag_embed(torch.full((B, 3), length, device=dev))
d = compute_distances_in_batch(cart, cell, 5.0)
ms = timed(lambda: compute_distances_in_batch(cart, cell, 5.0), repeats=5)
report("pair_distances count + fill (compute_distances_in_batch)", ms, B * (12 * N + 36) + 4 * d.numel(),
       f"{B} x {N} atoms, {d.numel() / B:.0f} distances/structure")
# ---- Stillinger-Weber energy and repulsive pseudo-force ----------------------------------------------------------------
oracle = create_energy_oracle(create_energy_oracle_parameters(dict(name="lammps", sw_coeff_filename="Si.sw"), ["Si"]))
lattice = torch.zeros(B, 6, device=dev)
lattice[:, :3] = length
axl = AXL(A=torch.zeros(B, N, dtype=torch.long, device=dev), X=cart / length, L=lattice)
ms = timed(lambda: oracle.compute_oracle_energies({AXL_COMPOSITION: axl}), repeats=5)
report("stillinger_weber_energy_kernel", ms, B * (20 * N + 24 + 4), f"{B} x {N} atoms (uniform random): compute-bound")


class _Zero(torch.nn.Module):
    def forward(self, batch, conditional=None):
        c = batch[NOISY_AXL_COMPOSITION]
        return AXL(A=c.A, X=torch.zeros_like(c.X), L=c.L)


ff = ForceFieldAugmentedScoreNetwork(_Zero(), ForceFieldParameters(radial_cutoff=1.5, strength=1.0))
ms = timed(lambda: ff.get_relative_coordinates_pseudo_force({NOISY_AXL_COMPOSITION: axl}), repeats=5)
report("repulsive_pseudo_force_kernel", ms, B * (24 * N + 24), f"{B} x {N} atoms, r0 = 1.5: compute-bound image sweep")
print(json.dumps({"peak_source": "MEASURED_PEAKS.json hbm_gbs", "gpu": torch.cuda.get_device_name(0), "kernels": rows}, indent=1))

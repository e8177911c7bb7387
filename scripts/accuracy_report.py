#!/usr/bin/env python
"""Accuracy of the three edge/node-MLP implementations against the fp32 oracle AND an fp64 evaluation of the oracle
(the ground truth both fp32 results approximate).  Production widths, Si 2x2x2.  Run on the GPU box."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (  # noqa: E402
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, CARTESIAN_FORCES, NOISE,  # noqa: E402
                                                                       NOISY_AXL_COMPOSITION, TIME)
from oracle import pc_oracle  # noqa: E402

hp = dict(num_atom_types=2, edges="radial_cutoff", radial_cutoff=7.5, n_layers=4, message_n_hidden_dimensions=4,
          message_hidden_dimensions_size=256, node_n_hidden_dimensions=4, node_hidden_dimensions_size=256,
          coordinate_n_hidden_dimensions=4, coordinate_hidden_dimensions_size=256)
full = dict(residual=True, attention=False, normalize=False, tanh=False, coords_agg="mean", message_agg="mean", **hp)
device = torch.device("cuda:0")
torch.manual_seed(11)
net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp))
sd32 = {k: v.detach().clone() for k, v in net.state_dict().items()}
sd64 = {k: v.double() for k, v in sd32.items()}
net = net.to(device)
B, N = 8, 64
x = torch.rand(B, N, 3)
a = torch.randint(0, 3, (B, N))
lattice = torch.zeros(B, 6)
lattice[:, :3] = 11.084
sigma = torch.rand(B, 1) * 0.4 + 0.005
edges = pc_oracle.edges_with_radial_cutoff(x, torch.diag_embed(lattice[:, :3].clip(min=2.2 * 7.5)), 7.5)
l32, s32 = pc_oracle.egnn_score_forward(sd32, full, a, x, lattice, sigma, edges=edges)
l64, s64 = pc_oracle.egnn_score_forward(sd64, full, a, x.double(), lattice.double(), sigma.double(), edges=edges)
batch = {NOISY_AXL_COMPOSITION: AXL(A=a.to(device), X=x.to(device), L=lattice.to(device)),
         TIME: torch.full((B, 1), 0.5, device=device), NOISE: sigma.to(device),
         CARTESIAN_FORCES: torch.zeros(B, N, 3, device=device)}


def rel(u, v):
    u, v = u.double(), v.double()
    m = torch.isfinite(v)
    return float((u[m] - v[m]).abs().max() / v[m].abs().max())


report = {"oracle_fp32_vs_fp64": {"score": rel(s32, s64), "logits": rel(l32, l64)}}
for name, impl in (("ffma", _cabi.IMPL_FFMA), ("tf32x3", _cabi.IMPL_TCGEN05), ("f16x3", _cabi.IMPL_TCGEN05_F16X3),
                   ("f16x3ts", _cabi.IMPL_TCGEN05_F16X3_TS)):
    net.kernel_impl = impl
    out = net(batch, conditional=False)
    report[name] = {"score_vs_fp32_oracle": rel(out.X.cpu(), s32), "logits_vs_fp32_oracle": rel(out.A.cpu(), l32),
                    "score_vs_fp64": rel(out.X.cpu(), s64), "logits_vs_fp64": rel(out.A.cpu(), l64),
                    "score_abs_vs_fp64": float((out.X.cpu().double() - s64).abs().max())}
report["score_scale"] = float(s64.abs().max())
print(json.dumps(report, indent=1))

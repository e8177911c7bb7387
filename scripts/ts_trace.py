#!/usr/bin/env python
"""Cycle trace of egnn_edge_ts_kernel (CTA 0).  Needs a library built with B200MD_EXTRA_NVCC_FLAGS=-DB200MD_TS_TRACE.
    python scripts/ts_trace.py [samples] > gpurun_out/ts_trace.txt"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (  # noqa: E402
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, CARTESIAN_FORCES, NOISE,  # noqa: E402
                                                                          NOISY_AXL_COMPOSITION, TIME, UNIT_CELL)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = 64
dev = torch.device("cuda:0")
hp = dict(num_atom_types=1, edges="radial_cutoff", radial_cutoff=7.5, n_layers=4, message_n_hidden_dimensions=4,
          message_hidden_dimensions_size=256, node_n_hidden_dimensions=4, node_hidden_dimensions_size=256,
          coordinate_n_hidden_dimensions=4, coordinate_hidden_dimensions_size=256)
torch.manual_seed(0)
net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp)).to(dev)
net.kernel_impl = _cabi.IMPL_TCGEN05_F16X3_TS
x = torch.rand(B, N, 3, device=dev)
a = torch.zeros(B, N, dtype=torch.long, device=dev)
lattice = torch.zeros(B, 6, device=dev)
lattice[:, :3] = 10.86
batch = {NOISY_AXL_COMPOSITION: AXL(A=a, X=x, L=lattice), TIME: torch.full((B, 1), 0.5, device=dev),
         NOISE: torch.full((B, 1), 0.1, device=dev), CARTESIAN_FORCES: torch.zeros(B, N, 3, device=dev)}
if UNIT_CELL:
    batch[UNIT_CELL] = torch.diag_embed(lattice[:, :3])
lib = _cabi.lib()
buf = (ctypes.c_uint32 * (4 * 3 * 2048))()
for _ in range(2):
    net(batch, conditional=False)
torch.cuda.synchronize()
lib.b200md_debug_ts_trace(buf, 3 * 2048)   # reset
net(batch, conditional=False)
n = lib.b200md_debug_ts_trace(buf, 3 * 2048)
rec = np.frombuffer(buf, dtype=np.uint32).reshape(-1, 4)[:n]
# only the first launch of the forward (4 layers record in sequence; keep records until the clock goes backwards a lot)
for w, ev, arg, clk in rec:
    if w:
        print(w, ev, arg, clk)

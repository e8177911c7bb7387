#!/usr/bin/env python
"""Cycle trace of egnn_node_ts_kernel (CTA 0, tiles 2-3 of the launch that updates layer 1).  Needs a library built with
B200MD_EXTRA_NVCC_FLAGS=-DB200MD_TS_TRACE.
    python scripts/node_trace.py [samples] > gpurun_out/node_trace.txt"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (  # noqa: E402
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, CARTESIAN_FORCES, NOISE,  # noqa: E402
                                                                       NOISY_AXL_COMPOSITION, TIME)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = 64
dev = torch.device("cuda:0")
hp = dict(num_atom_types=1, edges="radial_cutoff", radial_cutoff=7.5, n_layers=4, message_n_hidden_dimensions=4,
          message_hidden_dimensions_size=256, node_n_hidden_dimensions=4, node_hidden_dimensions_size=256,
          coordinate_n_hidden_dimensions=4, coordinate_hidden_dimensions_size=256)
torch.manual_seed(0)
net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp)).to(dev)
x = torch.rand(B, N, 3, device=dev)
a = torch.zeros(B, N, dtype=torch.long, device=dev)
lattice = torch.zeros(B, 6, device=dev)
lattice[:, :3] = 10.86
batch = {NOISY_AXL_COMPOSITION: AXL(A=a, X=x, L=lattice), TIME: torch.full((B, 1), 0.5, device=dev),
         NOISE: torch.full((B, 1), 0.1, device=dev), CARTESIAN_FORCES: torch.zeros(B, N, 3, device=dev)}
lib = ctypes.CDLL(_cabi.LIBRARY_PATH)
records = 3 * 4096
buf = (ctypes.c_uint32 * (4 * records))()
for _ in range(2):
    net(batch, conditional=False)
torch.cuda.synchronize()
lib.b200md_debug_node_trace(buf, records)   # reset
net(batch, conditional=False)
n = lib.b200md_debug_node_trace(buf, records)
rec = np.frombuffer(buf, dtype=np.uint32).reshape(-1, 4)[:n]
rec = rec[rec[:, 0] != 0]
names = {1: "acc_wait", 2: "acc_ready", 5: "silu_done", 6: "published", 20: "tile_begin", 21: "first_A_published", 22: "agg_collect", 23: "agg_here", 24: "agg_divided", 25: "agg_zeroed", 26: "side_collect", 27: "side_here", 28: "acc_loaded", 29: "side_here", 31: "stored", 32: "side_requested", 33: "computed", 34: "rows_stored", 35: "staged", 30: "chunk_done",
         100: "A_wait_p0", 102: "A_wait_p1", 110: "A_got_p0", 112: "A_got_p1", 120: "grp0_committed", 121: "grp1_committed",
         130: "dfree_wait0", 131: "dfree_wait1", 140: "dfree_got0", 141: "dfree_got1"}
t0 = int(rec[:, 3].astype(np.int64).min())
events = sorted((int((int(c) - t0) & 0xffffffff), int(w), int(e), int(arg)) for w, e, arg, c in rec)
first_epi = min(w for _, w, _, _ in events if w != 1)
for t, w, e, arg in events:
    col = 0 if w == 1 else (40 if w == first_epi else 80)
    label = f"{names.get(e, e)} g{(arg & 0xfffff) // 16}c{arg % 16}" if e not in (20, 21, 22, 23, 24, 25, 120, 121, 130, 131, 140, 141) else f"{names.get(e, e)} {arg}"
    print(f"{t:9d} " + " " * col + label)

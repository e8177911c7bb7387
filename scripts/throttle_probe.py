#!/usr/bin/env python
"""Does the sustained step run slower than the same launches after an idle period?  Times 60 consecutive EGNN forwards
(4096 x 64 atoms, production widths) one by one after 5 s of idle, with nvidia-smi clocks / power sampled alongside."""
import json
import subprocess
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import measure_edges, network_hyper_parameters  # noqa: E402,F401
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (  # noqa: E402
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (AXL, CARTESIAN_FORCES, NOISE,  # noqa: E402
                                                                       NOISY_AXL_COMPOSITION, TIME)

dev = torch.device("cuda:0")
B, N = 4096, 64
torch.manual_seed(0)
net = EGNNScoreNetwork(EGNNScoreNetworkParameters(**network_hyper_parameters("cfg2"))).to(dev)
lattice = torch.zeros(B, 6, device=dev)
lattice[:, :3] = 10.86
batch = {NOISY_AXL_COMPOSITION: AXL(A=torch.zeros(B, N, dtype=torch.long, device=dev), X=torch.rand(B, N, 3, device=dev), L=lattice),
         TIME: torch.full((B, 1), 0.5, device=dev), NOISE: torch.full((B, 1), 0.1, device=dev),
         CARTESIAN_FORCES: torch.zeros(B, N, 3, device=dev)}
for _ in range(2):
    net(batch, conditional=False)
torch.cuda.synchronize()
time.sleep(5.0)
smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                       stdout=subprocess.PIPE, text=True)
events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
for start, stop in events:
    start.record()
    net(batch, conditional=False)
    stop.record()
torch.cuda.synchronize()
smi.terminate()
samples = [tuple(float(v) for v in line.split(",")) for line in smi.communicate()[0].strip().splitlines() if "," in line]
times = [round(a.elapsed_time(b), 2) for a, b in events]
print(json.dumps({"forward_ms": times, "clocks_power": samples}))

#!/usr/bin/env python
"""GPU timeline of a few predictor-corrector steps (CUPTI activity records through torch.profiler: concurrent, not
serialised like ncu): per-kernel durations and the idle gaps between consecutive kernels of the step.

    python scripts/step_timeline.py [--samples 4096] [--steps 3] > gpurun_out/timeline.txt
"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters  # noqa: E402
from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (  # noqa: E402
    EGNNScoreNetwork, EGNNScoreNetworkParameters)
from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--samples", type=int, default=4096)
parser.add_argument("--steps", type=int, default=3)
args = parser.parse_args()
device = torch.device("cuda:0")
torch.manual_seed(1234)
network = EGNNScoreNetwork(EGNNScoreNetworkParameters(**bench.network_hyper_parameters("cfg2"))).to(device)
sampling = PredictorCorrectorSamplingParameters(number_of_samples=args.samples, number_of_atoms=64, num_atom_types=1,
                                                number_of_corrector_steps=1, use_fixed_lattice_parameters=True,
                                                cell_dimensions=[10.86] * 3)
generator = LangevinGenerator(NoiseParameters(total_time_steps=1000), sampling, network)
generator.seed = 1
generator._begin_trajectory(device)
composition = generator.initialize(args.samples, device)
composition = generator.sample_from_noisy_composition(composition, 1000, 995)   # warm-up
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    composition = generator.sample_from_noisy_composition(composition, 995, 995 - args.steps)
    torch.cuda.synchronize()
events = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name
                 and "Memset" not in e.name), key=lambda e: e.time_range.start)
total = events[-1].time_range.end - events[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in events)
print(f"# {len(events)} kernels over {args.steps} steps of {args.samples} samples: span {total / 1e3:.3f} ms, "
      f"kernel time {busy / 1e3:.3f} ms, idle {(total - busy) / 1e3:.3f} ms ({100 * (total - busy) / total:.2f} %)")
gaps = collections.defaultdict(list)
for before, after in zip(events, events[1:]):
    gaps[(before.name[:40], after.name[:40])].append(after.time_range.start - before.time_range.end)
print("# idle gap between consecutive kernels (us): count, mean, max, total")
for (a, b), values in sorted(gaps.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(values):5d} {sum(values) / len(values):9.1f} {max(values):9.1f} {sum(values):10.1f}   {a} -> {b}")
print("# every launch of the first step (us): start offset, duration, name")
t0 = events[0].time_range.start
per_step = len(events) // args.steps
for e in events[: per_step + 2]:
    print(f"{e.time_range.start - t0:12.1f} {e.time_range.end - e.time_range.start:10.1f}   {e.name[:70]}")
durations = collections.defaultdict(list)
for e in events:
    durations[e.name[:60]].append(e.time_range.end - e.time_range.start)
print("# kernel durations (us): count, mean, total")
for name, values in sorted(durations.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(values):5d} {sum(values) / len(values):10.1f} {sum(values):12.1f}   {name}")

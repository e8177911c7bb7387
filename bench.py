#!/usr/bin/env python
"""Benchmark of the predictor-corrector sampling hot path (BASELINE.json metric: structures/sec, Si 2x2x2 EGNN
PC-Langevin, 1000 steps x 1 corrector, 4096 samples per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--samples B] [--workload cfg2]

A "step" is ONE predictor-corrector iteration (1 predictor + n_corr corrector score-network evaluations plus the
coordinate / atom-type updates) of the whole per-GPU batch; a structure costs T = 1000 such steps, so

    structures/sec = (samples on all ranks) / (T * seconds per step).

The K timed steps are consecutive time indices of a real trajectory starting at i = T (every step does the full
work of a production step: graph build, 4 E_GCL layers at production widths, updates, in-kernel noise).  The
working set of a step (node features, P/Q tiles, edge lists: > 1 GB at 4096 x 64 atoms) is far larger than the
126 MB L2, so successive steps cannot reuse cached inputs ("inputs larger than L2").
`--impl reference` times the CPU restatement of the reference path (oracle/pc_oracle.py, a port: the reference
itself cannot travel to the GPU box) on the host cores with the same metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, atoms, cell length, num_atom_types, sampling flags, default samples per GPU)
    "cfg2": ("Si 2x2x2 (64 atoms) EGNN production widths, PC Langevin 1000 steps x1 corrector", 64, 10.86, 1,
             dict(), 4096),
    "cfg3": ("SiGe 2x2x2 (64 atoms) joint coordinate + atom-type diffusion, 1000 steps x1 corrector", 64, 11.084, 2,
             dict(one_atom_type_transition_per_step=True, atom_type_greedy_sampling=True), 4096),
    "cfg4": ("Si 3x3x3 (216 atoms) repaint / constrained inpainting around a fixed cluster, 1000 steps x1 corrector",
             216, 16.29, 1, dict(), 2048),
    "cfg5": ("synthetic Si 4x4x4 (512 atoms) EGNN, 1000 steps x1 corrector", 512, 21.72, 1, dict(), 1024),
    "cfg1": ("Si 1x1x1 (8 atoms) EGNN, fully connected, PC Langevin", 8, 5.43, 1, dict(), 4096),
}
TOTAL_TIME_STEPS = 1000
# roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum of one edge-kernel launch) cannot be measured from inside
# this process; it is null on the line and the ncu --set full capture of the same command is committed under profiles/
# (ncu_full_*.txt: 1.128 GB per launch at 4096 samples against 0.805 GB of algorithmic P / Q / agg rows).
NCU_TRAFFIC_NOTE = "null: not measured in-run; see profiles/ncu_full_*.txt (ncu --set full of this command)"
PRODUCTION = dict(n_layers=4, message_n_hidden_dimensions=4, message_hidden_dimensions_size=256,
                  node_n_hidden_dimensions=4, node_hidden_dimensions_size=256, coordinate_n_hidden_dimensions=4,
                  coordinate_hidden_dimensions_size=256)
ORACLE_HP = dict(residual=True, attention=False, normalize=False, tanh=False, coords_agg="mean", message_agg="mean")


def network_hyper_parameters(workload: str) -> dict:
    _, atoms, _, num_atom_types, _, _ = WORKLOADS[workload]
    hp = dict(num_atom_types=num_atom_types, **PRODUCTION)
    if atoms <= 8:
        hp.update(edges="fully_connected")
    else:
        hp.update(edges="radial_cutoff", radial_cutoff=7.5)
    return hp


def diamond_cluster_constraint(cells: int, length: float, radius: float):
    """Atoms of the perfect diamond lattice (cells^3 conventional cells) within `radius` of the cell centre: the fixed
    cluster of BASELINE config 4, as a SamplingConstraint occupying the first atom slots."""
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.sampling_constraint import SamplingConstraint
    basis = torch.tensor([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.25, .25, .25], [.25, .75, .75],
                          [.75, .25, .75], [.75, .75, .25]])
    grid = torch.stack(torch.meshgrid(*[torch.arange(cells)] * 3, indexing="ij"), dim=-1).reshape(-1, 1, 3)
    relative = ((grid + basis.unsqueeze(0)) / cells).reshape(-1, 3)
    distance = ((relative - 0.5) * length).norm(dim=1)
    chosen = relative[distance <= radius].float().contiguous()
    return SamplingConstraint(elements=["Si"], constrained_relative_coordinates=chosen,
                              constrained_atom_types=torch.zeros(len(chosen), dtype=torch.long),
                              constrained_indices=torch.arange(len(chosen)))


def algorithmic_flops(hp: dict, n_atoms: int, edges_per_structure: float) -> dict:
    """1 MAC = 2 FLOP; activations, gathers and reductions not counted."""
    H, M, Cx = hp["node_hidden_dimensions_size"], hp["message_hidden_dimensions_size"], hp["coordinate_hidden_dimensions_size"]
    nm, nn, nc = hp["message_n_hidden_dimensions"], hp["node_n_hidden_dimensions"], hp["coordinate_n_hidden_dimensions"]
    f_edge = 2 * ((2 * H + 1) * M + nm * M * M + M * Cx + nc * Cx * Cx + Cx)
    f_node = 2 * ((H + M) * H + nn * H * H + H * H)
    layers = hp["n_layers"]
    C, F0 = hp["num_atom_types"] + 1, hp["num_atom_types"] + 2
    per_forward = layers * (edges_per_structure * f_edge + n_atoms * f_node) + n_atoms * (2 * F0 * H + 2 * H * C)
    return dict(f_edge=f_edge, f_node=f_node, per_forward=per_forward,
                edge_per_forward=layers * edges_per_structure * f_edge)


def workload_config(workload: str, samples_per_gpu: int, world: int) -> dict:
    """The `config` object of the JSON line: identical for the B200 arm and the reference arm on the same workload."""
    description, atoms, _, _, _, _ = WORKLOADS[workload]
    return {"workload": f"{workload}: {description}", "samples_per_gpu": samples_per_gpu, "atoms": atoms,
            "total_time_steps": TOTAL_TIME_STEPS, "number_of_corrector_steps": 1,
            "step": "one predictor-corrector iteration (2 EGNN forwards + updates) of the per-GPU batch; "
                    "value = samples / (1000 x step time)",
            "network": "EGNN 4 x E_GCL, hidden 256 x 4 (4,738,679 parameters), random init, rc 7.5 A",
            "l2": "inputs larger than L2", "parallelism": f"samples sharded x{world}"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.process = None
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.process = None

    def stop(self) -> dict:
        if self.process is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.process.terminate()
        try:
            out, _ = self.process.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.process.kill()
            out, _ = self.process.communicate()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); sm_max.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(sm_max), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path
# ---------------------------------------------------------------------------------------------------------------


def cpu_reference_run(workload: str, samples: int, steps: int, warmup: int, seed: int = 1234):
    """Time `steps` PC iterations of `samples` structures with the oracle (CPU port of the reference). Returns
    (seconds per step, cores)."""
    from oracle import pc_oracle
    from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
        EGNNScoreNetwork, EGNNScoreNetworkParameters)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    _, atoms, length, num_atom_types, flags, _ = WORKLOADS[workload]
    hp = network_hyper_parameters(workload)
    torch.manual_seed(seed)
    network = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp))  # parameter container only (random init)
    sd = {k: v.detach() for k, v in network.state_dict().items()}
    full_hp = dict(ORACLE_HP, **hp)
    tables = pc_oracle.noise_tables(TOTAL_TIME_STEPS, num_atom_types + 1)
    x = torch.rand(samples, atoms, 3)
    a = torch.full((samples, atoms), num_atom_types, dtype=torch.long)
    lattice = torch.zeros(samples, 6)
    lattice[:, :3] = length

    def score_fn(at, xx, lat, sigma):
        return pc_oracle.egnn_score_forward(sd, full_hp, at, xx, lat, torch.full((samples, 1), sigma))

    tape = pc_oracle.NoiseTape(seed)
    index = TOTAL_TIME_STEPS
    with torch.no_grad():
        for _ in range(warmup):
            a, x, _ = pc_oracle.pc_sample(score_fn, tables, a, x, lattice, tape, starting_step_index=index,
                                          ending_step_index=index - 1)
            index -= 1
        t0 = time.perf_counter()
        for _ in range(steps):
            a, x, _ = pc_oracle.pc_sample(score_fn, tables, a, x, lattice, tape, starting_step_index=index,
                                          ending_step_index=index - 1)
            index -= 1
        elapsed = time.perf_counter() - t0
    return elapsed / steps, cores


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    workload = args.workload
    samples = args.cpu_samples
    seconds_per_step, cores = cpu_reference_run(workload, samples, args.steps, args.warmup)
    value = samples / (TOTAL_TIME_STEPS * seconds_per_step)
    line = {
        "impl": "reference", "metric": "structures/sec", "value": value, "unit": "structures/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the workload is the B200 arm's; each timed step processes a bounded sample of it (cpu_baseline.sample)
        "config": workload_config(workload, args.samples or WORKLOADS[workload][5], 1),
        "cpu_baseline": {"value": value, "unit": "structures/s", "cores": cores, "kind": "port",
                         "sample": f"{samples} structures x {args.steps} PC steps of T=1000 (after {args.warmup} "
                                   "warm-up steps), extrapolated linearly in the number of structures and steps; "
                                   "oracle/pc_oracle.py (torch CPU restatement of the reference; dense neighbour "
                                   "search stand-in for pykeops)"},
        "e2e": {"value": value, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------


def run_b200_arm(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist

    from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.langevin_generator import LangevinGenerator
    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
        PredictorCorrectorSamplingParameters
    from diffusion_for_multi_scale_molecular_dynamics_b200.models.score_networks.egnn_score_network import (
        EGNNScoreNetwork, EGNNScoreNetworkParameters)
    from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL
    from diffusion_for_multi_scale_molecular_dynamics_b200.noise_schedulers.noise_parameters import NoiseParameters

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    _cabi.lib()  # fail loudly if the extension is missing
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    workload = args.workload
    description, atoms, length, num_atom_types, flags, default_samples = WORKLOADS[workload]
    samples = args.samples or default_samples
    hp = network_hyper_parameters(workload)
    torch.manual_seed(1234)
    network = EGNNScoreNetwork(EGNNScoreNetworkParameters(**hp)).to(device)
    network.kernel_impl = {"auto": _cabi.IMPL_AUTO, "ffma": _cabi.IMPL_FFMA, "tf32x3": _cabi.IMPL_TCGEN05,
                           "f16x3": _cabi.IMPL_TCGEN05_F16X3, "f16x3ts": _cabi.IMPL_TCGEN05_F16X3_TS}[args.kernels]
    noise = NoiseParameters(total_time_steps=TOTAL_TIME_STEPS)
    sampling = PredictorCorrectorSamplingParameters(
        number_of_samples=samples * world, number_of_atoms=atoms, num_atom_types=num_atom_types,
        number_of_corrector_steps=1, use_fixed_lattice_parameters=True, cell_dimensions=[length] * 3, **flags)
    if workload == "cfg4":
        from diffusion_for_multi_scale_molecular_dynamics_b200.generators.constrained_langevin_generator import \
            ConstrainedLangevinGenerator
        generator = ConstrainedLangevinGenerator(noise, sampling, network, diamond_cluster_constraint(3, length, 5.0))
    else:
        generator = LangevinGenerator(noise, sampling, network)
    generator.seed = 1234
    generator.sample_offset = rank * samples  # weak scaling: every rank owns `samples` structures
    K, W = args.steps, max(args.warmup, 3)
    assert W + 2 * K + 5 <= TOTAL_TIME_STEPS
    if args.full_trajectory:
        run_full_trajectory(args, generator, network, samples, world, rank, device, description)
        return
    if args.strong:
        run_strong(args, generator, network, noise, sampling, world, rank, local_rank, device)
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident measurement: K consecutive PC steps of one trajectory --------------------------------
    generator._begin_trajectory(device)
    composition = generator.initialize(samples, device)
    index = TOTAL_TIME_STEPS
    # warm-up exactly like the timed region: one call of W steps (W >= 3, so the captured step graph is exercised too)
    composition = generator.sample_from_noisy_composition(composition, index, index - W)
    index -= W
    edges_per_structure = measure_edges(network, composition, device)

    def gather_axl(axl):
        """The single collective of the path: all_gather of the sampled (A, X, L)."""
        for part in (axl.A, axl.X, axl.L):
            gathered = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(gathered, part.contiguous())

    if world > 1:
        gather_axl(composition)   # warm-up of the collective too (NCCL builds its channels on the first call per shape)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches_before = _cabi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    # the K timed steps are ONE call, as in production (generator.sample() hands the whole trajectory to
    # b200md_pc_sample): no host synchronisation between steps
    composition = generator.sample_from_noisy_composition(composition, index, index - K)
    index -= K
    if world > 1:
        gather_axl(composition)
    stop.record()
    barrier()
    ms_total = max_over_ranks(start.elapsed_time(stop))
    launches = _cabi.launch_count() - launches_before
    clocks = sampler.stop() if sampler is not None else None
    ms_per_step = ms_total / K
    value = samples * world / (TOTAL_TIME_STEPS * ms_per_step * 1e-3)

    # ---- per-kernel-class timing of the same steps (in-stream CUDA events inside the library) ------------------
    _cabi.profile_enable(True)
    profile_steps = min(K, 3)
    for _ in range(profile_steps):
        composition = generator.sample_from_noisy_composition(composition, index, index - 1)
        index -= 1
    torch.cuda.synchronize(device)
    classes = _cabi.profile_collect()
    _cabi.profile_enable(False)
    edge_ms, edge_launches = classes["edge"]
    total_class_ms = sum(ms for ms, _ in classes.values())
    flops = algorithmic_flops(hp, atoms, edges_per_structure)
    roofline = None
    peaks = load_peaks()
    if edge_launches > 0:
        avg_launch_s = edge_ms * 1e-3 / edge_launches
        flop_per_launch = samples * edges_per_structure * flops["f_edge"]
        achieved = flop_per_launch / avg_launch_s / 1e12
        kernel_name = network.last_edge_kernel_name()
        # Denominator: the measured dense bf16/fp16 tensor throughput (sustained: the kernel is timed inside a long,
        # power-capped step).  The 3xFP16 kernel issues three fp16 MMAs per algorithmic MAC (fp32-exact split), the
        # 3xTF32 kernel three TF32 MMAs at half that rate; neither can exceed 1/3 (resp. 1/6) of the peak times
        # 1.22 (the first message Linear is hoisted to the nodes: 1,180,160 of the 1,442,816 FLOP/edge run per edge).
        peak = peaks["bf16_tflops_sustained"]
        passes, rate = (3, 1.0) if "3xFP16" in kernel_name else ((3, 0.5) if "3xTF32" in kernel_name else (1, None))
        executed_fraction = 1180160.0 / 1442816.0 if hp["message_hidden_dimensions_size"] == 256 else None
        ceiling = None
        if rate is not None and executed_fraction is not None:
            ceiling = peak * rate / passes / executed_fraction
        roofline = {"bound": "tensor", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None, "traffic_note": NCU_TRAFFIC_NOTE,
                    "peak_source": peaks["source"] + ": dense bf16/fp16, sustained",
                    "fp32_exact_passes": passes,
                    "fp32_exact_ceiling_tflops": ceiling,
                    "frac_of_fp32_exact_ceiling": (achieved / ceiling) if ceiling else None,
                    # MMA FLOPs the tensor pipe actually executes (passes x executed share, in fp16-rate units) against
                    # the same sustained dense peak: how close the kernel is to what the power-capped chip can issue
                    "executed_dense_equivalent_tflops": (achieved * passes * executed_fraction / rate)
                    if (rate and executed_fraction) else None,
                    "algorithmic_flop_per_launch": flop_per_launch, "avg_launch_ms": avg_launch_s * 1e3,
                    "share_of_step": edge_ms / total_class_ms if total_class_ms else None,
                    "class_ms_per_step": {k: v[0] / profile_steps for k, v in classes.items()},
                    "class_launches_per_step": {k: v[1] / profile_steps for k, v in classes.items()},
                    "edges_per_structure": edges_per_structure}

    # ---- end-to-end through the public API with host buffers ----------------------------------------------------
    host_x = composition.X.cpu().pin_memory()
    host_a = composition.A.cpu().pin_memory()
    host_l = composition.L.cpu().pin_memory()
    out_x = torch.empty_like(host_x).pin_memory()
    out_a = torch.empty_like(host_a).pin_memory()
    e2e_steps = K
    assert index - 1 - e2e_steps >= 0, "not enough time indices left for the end-to-end leg; lower --steps / --warmup"

    def e2e_step(idx):
        comp = AXL(A=host_a.to(device, non_blocking=True), X=host_x.to(device, non_blocking=True),
                   L=host_l.to(device, non_blocking=True))
        result = generator.sample_from_noisy_composition(comp, idx, idx - 1)
        out_x.copy_(result.X, non_blocking=True)
        out_a.copy_(result.A, non_blocking=True)

    e2e_step(index)
    barrier()
    start.record()
    for k in range(e2e_steps):   # every step: pinned host -> device, one PC step through the public API, device -> host
        e2e_step(index - 1 - k)
    stop.record()
    barrier()
    e2e_ms = max_over_ranks(start.elapsed_time(stop)) / e2e_steps
    h2d = host_x.numel() * 4 + host_a.numel() * 8 + host_l.numel() * 4
    d2h = out_x.numel() * 4 + out_a.numel() * 8
    e2e = {"value": samples * world / (TOTAL_TIME_STEPS * e2e_ms * 1e-3), "unit": "structures/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms}

    if rank != 0:
        return
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        seconds_per_step, cores = cpu_reference_run(workload, args.cpu_samples, 2, 1)
        cpu_baseline = {"value": args.cpu_samples / (TOTAL_TIME_STEPS * seconds_per_step), "unit": "structures/s",
                        "cores": cores, "kind": "port",
                        "sample": f"{args.cpu_samples} structures x 2 PC steps of T=1000 (after 1 warm-up step), "
                                  "extrapolated linearly; oracle/pc_oracle.py"}
    line = {
        "metric": "structures/sec", "value": value, "unit": "structures/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(workload, samples, world),
        "kernels": {"edge_mlp": network.last_edge_kernel_name(), "flop_per_structure_forward": flops["per_forward"],
                    "loop": "one CUDA graph launch per PC step (b200md_pc_sample)"},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)


def run_strong(args, generator, network, noise, sampling, world, rank, local_rank, device):
    """Strong scaling through the product shard point: ``create_batch_of_samples`` (sampling/diffusion_sampling.py)
    draws a FIXED total number of samples (the configuration as BASELINE states it: 4096 for cfg2, 8192 for cfg5),
    sharded over the ranks by ``shard_range`` with the counter-based noise keyed by the global sample index, and
    gathers (A, X, L) on every rank with NCCL at the end.  To keep the run bounded the trajectory is the last K steps of
    the T = 1000 schedule (time indices K .. 1, started from a fresh random configuration): per-step work is the
    production step's; value = total samples / (1000 x time per step)."""
    import torch.distributed as dist

    from diffusion_for_multi_scale_molecular_dynamics_b200.generators.trajectory_initializer import \
        FullRandomTrajectoryInitializer
    from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL_COMPOSITION
    from diffusion_for_multi_scale_molecular_dynamics_b200.sampling.diffusion_sampling import create_batch_of_samples

    total = args.samples or {"cfg5": 8192}.get(args.workload, WORKLOADS[args.workload][5])
    K, W = args.steps, max(args.warmup, 3)
    steps_to_run = {"n": W}

    class LastSteps(type(generator.trajectory_initializer)):
        def create_start_time_step_index(self, number_of_discretization_steps: int) -> int:
            return steps_to_run["n"]

    assert isinstance(generator.trajectory_initializer, FullRandomTrajectoryInitializer)
    generator.trajectory_initializer.__class__ = LastSteps
    generator.sample_offset = 0
    sampling.number_of_samples = total
    sampling.sample_batchsize = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    create_batch_of_samples(generator, sampling, device)   # warm-up: W steps, packing, workspace, NCCL channels
    steps_to_run["n"] = K
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    from diffusion_for_multi_scale_molecular_dynamics_b200 import _cabi
    launches_before = _cabi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    batch = create_batch_of_samples(generator, sampling, device)
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = _cabi.launch_count() - launches_before
    clocks = sampler.stop() if sampler is not None else None
    axl = batch[AXL_COMPOSITION]
    assert axl.X.shape[0] == total and axl.A.shape[0] == total and axl.L.shape[0] == total
    if rank != 0:
        return
    ms_per_step = ms / K
    config = workload_config(args.workload, None, world)
    config.update(samples_total=total, samples_per_gpu=f"{total} / {world}",
                  trajectory=f"last {K} steps of the T=1000 schedule through create_batch_of_samples, incl. "
                             "initialisation and the final NCCL all_gather of (A, X, L)")
    print(json.dumps({
        "metric": "structures/sec", "value": total / (TOTAL_TIME_STEPS * ms_per_step * 1e-3), "unit": "structures/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "kernels": {"edge_mlp": network.last_edge_kernel_name()}, "gpu_launches": int(launches), "clocks": clocks,
        "all_unmasked": bool((axl.A != generator.masked_atom_type_index).all()),
        "gathered_bytes_per_rank": int(axl.X[: total // world].numel() * 4 + axl.A[: total // world].numel() * 8
                                       + axl.L[: total // world].numel() * 4)}), flush=True)


def run_full_trajectory(args, generator, network, samples, world, rank, device, description):
    """Validation of the per-step extrapolation: time generator.sample() -- initialisation, all 1000 steps, final
    device->host copy -- for `samples` structures per GPU and report structures/sec directly."""
    import torch.distributed as dist
    generator.sample(min(samples, 64), device)  # warm-up (weight packing, workspace)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    result = generator.sample(samples, device)
    host = (result.X.cpu(), result.A.cpu())
    stop.record()
    torch.cuda.synchronize(device)
    seconds = start.elapsed_time(stop) * 1e-3
    if world > 1:
        t = torch.tensor([seconds], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "structures/sec", "mode": "full trajectory (1000 PC steps, generator.sample)",
                          "value": samples * world / seconds, "unit": "structures/s", "n_gpus": world,
                          "seconds": seconds, "samples_per_gpu": samples, "workload": description,
                          "edge_mlp": network.last_edge_kernel_name(),
                          "all_unmasked": bool((host[1] != generator.masked_atom_type_index).all()),
                          "x_in_unit_cell": bool(((host[0] >= 0) & (host[0] < 1)).all())}), flush=True)


def measure_edges(network, composition, device) -> float:
    """Directed edges per structure of the current batch (one extra forward, outside the timed region)."""
    from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import (CARTESIAN_FORCES, NOISE,
                                                                           NOISY_AXL_COMPOSITION, TIME)
    samples = composition.X.shape[0]
    batch = {NOISY_AXL_COMPOSITION: composition, TIME: torch.full((samples, 1), 0.5, device=device),
             NOISE: torch.full((samples, 1), 0.1, device=device), CARTESIAN_FORCES: torch.zeros_like(composition.X)}
    network(batch, conditional=False)
    return network.last_number_of_edges() / samples


def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fd:
            data = json.load(fd)
        return {"hbm_gbs": data["hbm_gbs"], "bf16_tflops": data["bf16_tflops"],
                "bf16_tflops_sustained": data.get("bf16_tflops_sustained", data["bf16_tflops"]),
                "source": "of measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "of fallback (B200_PROFILING.md)"}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", choices=["b200", "reference"], default="b200")
    parser.add_argument("--workload", choices=list(WORKLOADS), default="cfg2")
    parser.add_argument("--samples", type=int, default=None, help="samples per GPU (default: the workload's)")
    parser.add_argument("--cpu-samples", type=int, default=32, help="structures per CPU-baseline step")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--full-trajectory", action="store_true",
                        help="time one complete generator.sample() of --samples structures instead of K steps")
    parser.add_argument("--strong", action="store_true",
                        help="strong scaling: a fixed TOTAL number of samples (--samples, default the configuration's: "
                             "4096 for cfg2, 8192 for cfg5) sharded over the ranks by create_batch_of_samples")
    parser.add_argument("--kernels", choices=["auto", "ffma", "tf32x3", "f16x3", "f16x3ts"], default="auto",
                        help="edge/node MLP implementation (auto = what the library picks for the configuration)")
    args = parser.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_b200_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python scripts/summarize_ncu.py <tag>      # reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(tag: str) -> str:
    with open(f"gpurun_out/launches_{tag}.csv") as fd:
        lines = [line for line in fd if not line.startswith("==")]
    totals = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        value = float(row["Metric Value"].replace(",", ""))
        value *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        totals[row["Kernel Name"]][0] += 1
        totals[row["Kernel Name"]][1] += value
    whole = sum(v[1] for v in totals.values())
    out = [f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, cold-cache serialised launches -- compare SHARES",
           f"# total {whole:.3f} ms over {sum(v[0] for v in totals.values())} launches"]
    for name, (count, ms) in sorted(totals.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{ms / whole * 100:7.2f}%  {ms:11.3f} ms  {count:5d} launches  {name[:110]}")
    return "\n".join(out) + "\n"


def full(tag: str) -> str:
    import glob
    reports = sorted(glob.glob(f"gpurun_out/prof_{tag}*.ncu-rep"))
    return "".join(full_one(tag, report) for report in reports)


def full_one(tag: str, report: str) -> str:
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    header, units = rows[0], rows[1]
    index = {h: i for i, h in enumerate(header)}
    out = [f"# ncu --set full ({tag}, {report.split('/')[-1]}), one block per captured launch"]
    for row in rows[2:]:
        out.append(f"kernel: {row[index['Kernel Name']]}")
        for metric in METRICS:
            if metric in index:
                out.append(f"  {metric} = {row[index[metric]]} {units[index[metric]]}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    tag = sys.argv[1]
    open(f"profiles/launches_{tag}.txt", "w").write(launches(tag))
    open(f"profiles/ncu_full_{tag}.txt", "w").write(full(tag))
    print(open(f"profiles/launches_{tag}.txt").read()[:1500])
    print(open(f"profiles/ncu_full_{tag}.txt").read()[:2500])

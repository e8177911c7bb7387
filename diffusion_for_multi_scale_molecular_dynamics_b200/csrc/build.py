"""Build ``libb200md.so`` in-tree with nvcc for sm_100a (no torch headers: the library is plain C ABI).

    python -m diffusion_for_multi_scale_molecular_dynamics_b200.csrc.build [--force] [--verbose]

Objects are compiled in parallel and cached under ``csrc/build/`` keyed by source mtime.
"""
import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

CSRC = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "elementwise.cu", "graph.cu", "structure.cu", "egnn.cu", "egnn_tc.cu", "egnn_node_tc.cu", "sampler.cu"]
HEADERS = ["common.cuh", "graph.cuh", "egnn.cuh", "tc_common.cuh", "egnn_edge_ts.cuh", "ts_common.cuh", "egnn_node_ts.cuh", "sampler.cuh", os.path.join("..", "..", "include", "b200md.h")]
LIBRARY = os.path.join(CSRC, "libb200md.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found; the sm_100a extension cannot be built")
    return nvcc


def _newest_header_mtime() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = find_nvcc()
    build_dir = os.path.join(CSRC, "build")
    os.makedirs(build_dir, exist_ok=True)
    header_time = _newest_header_mtime()
    jobs = []
    objects = []
    for src in SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        objects.append(obj)
        stale = force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src_path), header_time)
        if stale:
            cmd = [nvcc] + NVCC_FLAGS + os.environ.get("B200MD_EXTRA_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", src_path, "-o", obj]
            jobs.append(cmd)
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs))
    for cmd, res in zip(jobs, results):
        if verbose or res.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {cmd[-3]}")
    if jobs or force or not os.path.isfile(LIBRARY):
        link = [nvcc, "-shared", "-o", LIBRARY] + objects + ["-lcudart"]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libb200md.so failed")
    return LIBRARY


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--force", action="store_true")
    parser.add_argument("--verbose", action="store_true")
    args = parser.parse_args()
    print(build(force=args.force, verbose=args.verbose))

"""CPU, world_size 2 over gloo: the multi-GPU shard point of the sampling path (sampling/diffusion_sampling.py).

The generator is a stand-in that returns samples determined by their GLOBAL index (what the counter-based noise of
the CUDA path guarantees), so the test checks the host logic: contiguous partition, per-chunk sample_offset, the
final all_gather, and that the assembled batch is identical to the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_for_multi_scale_molecular_dynamics_b200.generators.predictor_corrector_axl_generator import \
    PredictorCorrectorSamplingParameters
from diffusion_for_multi_scale_molecular_dynamics_b200.namespace import AXL, AXL_COMPOSITION, CARTESIAN_POSITIONS
from diffusion_for_multi_scale_molecular_dynamics_b200.sampling.diffusion_sampling import (chunk_ranges,
                                                                                          create_batch_of_samples,
                                                                                          shard_range)

N_ATOMS = 5


class IndexKeyedGenerator:
    """sample i is a pure function of its global index, like the Philox-keyed CUDA generator."""

    def __init__(self):
        self.sample_offset = 0
        self.calls = []

    def sample(self, number_of_samples, device):
        self.calls.append((self.sample_offset, number_of_samples))
        index = torch.arange(self.sample_offset, self.sample_offset + number_of_samples, dtype=torch.float32)
        x = (index.view(-1, 1, 1) * 0.01 + torch.arange(N_ATOMS * 3).view(1, N_ATOMS, 3) * 0.001) % 1.0
        a = (index.long().view(-1, 1) + torch.arange(N_ATOMS).view(1, -1)) % 2
        lattice = torch.zeros(number_of_samples, 6)
        lattice[:, :3] = 5.43
        return AXL(A=a, X=x, L=lattice)


def sampling_parameters(n, chunk):
    return PredictorCorrectorSamplingParameters(number_of_samples=n, sample_batchsize=chunk, number_of_atoms=N_ATOMS,
                                                num_atom_types=1, use_fixed_lattice_parameters=True,
                                                cell_dimensions=[5.43] * 3)


def test_shard_and_chunk_ranges():
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert chunk_ranges(3, 10, 3) == [(3, 6), (6, 9), (9, 10)]
    assert chunk_ranges(5, 5, 4) == []


def test_single_process_matches_reference_chunking():
    gen = IndexKeyedGenerator()
    batch = create_batch_of_samples(gen, sampling_parameters(10, 4), torch.device("cpu"))
    assert gen.calls == [(0, 4), (4, 4), (8, 2)]
    whole = IndexKeyedGenerator().sample(10, "cpu")
    assert torch.equal(batch[AXL_COMPOSITION].X, whole.X) and torch.equal(batch[AXL_COMPOSITION].A, whole.A)
    torch.testing.assert_close(batch[CARTESIAN_POSITIONS], whole.X * 5.43)
    assert gen.sample_offset == 0  # restored


def _worker(rank, world, port, n, chunk, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = IndexKeyedGenerator()
        batch = create_batch_of_samples(gen, sampling_parameters(n, chunk), torch.device("cpu"))
        axl = batch[AXL_COMPOSITION]
        queue.put((rank, gen.calls, axl.A.clone(), axl.X.clone(), axl.L.clone(), batch[CARTESIAN_POSITIONS].clone()))
    finally:
        dist.destroy_process_group()


def _run_two_ranks(n, chunk):
    """Spawn the two gloo ranks on a free port; None if the rendezvous failed (the port can be taken between the probe
    and the bind), so the caller can retry on another port."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, chunk, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = []
    try:
        results = [queue.get(timeout=120) for _ in procs]
    except Exception:  # queue.Empty: a rank died before reporting
        pass
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
            p.join()
    if len(results) != 2 or any(p.exitcode != 0 for p in procs):
        return None
    return sorted(results, key=lambda t: t[0])


@pytest.mark.parametrize("n,chunk", [(11, 3), (8, None), (1, 4)])
def test_two_ranks_assemble_the_single_process_batch(n, chunk):
    results = None
    for _ in range(3):
        results = _run_two_ranks(n, chunk)
        if results is not None:
            break
    assert results is not None, "two-rank gloo run failed three times"
    expected = IndexKeyedGenerator().sample(n, "cpu")
    for rank, calls, a, x, lattice, cart in results:
        lo, hi = shard_range(n, rank, 2)
        assert [c[0] for c in calls] == [r[0] for r in chunk_ranges(lo, hi, chunk or n)]
        assert sum(c[1] for c in calls) == hi - lo
        assert torch.equal(a, expected.A) and torch.equal(x, expected.X) and torch.equal(lattice, expected.L)
        torch.testing.assert_close(cart, expected.X * 5.43)

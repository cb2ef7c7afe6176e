"""The N>1 path on CPU: two gloo ranks shard a read set in contiguous blocks, seed their block
(with the CPU oracle standing in for the GPU, which this container does not have), and rank 0 gathers
in input order.  The gathered result must equal the single-process result (determinism contract,
SURVEY 8b/8e): output independent of the shard count."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from compseed_b200 import sharding, synth
    from oracle import oracle_py as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = synth.random_reference(60_000, seed=401)
    bases, off, _ = synth.simulate_reads(ref, 1700, [100, 150], 0.01, seed=402, n_rate=0.001)
    idx = O.OracleIndex.build(ref)          # replica of the index on every rank
    b, o = sharding.take_shard(bases, off, sharding.shard_bounds(1700, world)[rank])
    r = idx.seed(b, o)
    parts = [None] * world if rank == 0 else None
    dist.gather_object((r.mem_off, r.mems, r.seed_off, r.rbeg), parts, dst=0)
    dist.barrier()
    if rank == 0:
        got = sharding.gather_in_input_order(parts)
        want = idx.seed(bases, off)
        q.put(all(np.array_equal(a, b) for a, b in zip(got, (want.mem_off, want.mems, want.seed_off, want.rbeg))))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(oracle_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_bounds_are_contiguous_and_block_aligned():
    from compseed_b200.sharding import shard_bounds
    for n, w in [(1700, 2), (10_000_000, 8), (5, 4), (512, 3), (0, 2)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        for (s0, e0), (s1, e1) in zip(b, b[1:]):
            assert e0 == s1 and s0 <= e0
        assert all(s % 512 == 0 for s, _ in b if s < n)

"""The multi-device pipeline behind the C-ABI (cs_multi_*) and the compact wire format, on the GPU box.
Every comparison is bit for bit against the oracle (itself pinned to the reference); the two-GPU cases run when the
box has more than one device (`gpurun --gpus 2`)."""
import numpy as np
import pytest

from compseed_b200 import synth

pytestmark = pytest.mark.gpu


def _same(r, w):
    assert np.array_equal(r.mem_off.astype(np.uint64), w.mem_off.astype(np.uint64))
    assert np.array_equal(r.mems, w.mems)
    assert np.array_equal(r.seed_off.astype(np.uint64), w.seed_off.astype(np.uint64))
    assert np.array_equal(r.rbeg, w.rbeg)


@pytest.fixture(scope="module")
def workload(oracle_lib):
    ref = synth.repeat_rich_reference(400_000, seed=801, n_segdup=60, segdup_len=2000, n_tandem=40)
    sets = []
    for i, n in enumerate((21_000, 9_001)):      # neither a multiple of the batch size nor of 512
        bases, off, _ = synth.simulate_reads(ref, n, [100, 150, 250], 0.015, seed=802 + i, n_rate=0.002)
        sets.append((bases, off))
    oi = oracle_lib.OracleIndex.build(ref)
    want = [oi.seed(b, o, n_threads=8) for b, o in sets]
    return ref, oi, sets, want


def test_compact_wire_format_equals_plain_results(cuda_lib, workload):
    """cs_seed_batch_wait_compact (20 B per mem, 5 B per seed position) expands to exactly what cs_seed_batch_wait returns."""
    _, oi, sets, want = workload
    idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=1)
    bases, off = sets[1]
    n = off.shape[0] - 1
    ctx = cuda_lib.SeedContext(idx, n, int(off[-1]), 256, n * 64, n * 600, 2, cuda_lib.CtxConfig(compact_results=1))
    ctx.submit(0, bases, off, cuda_lib.SeedOpt())
    ctx.submit(1, bases, off, cuda_lib.SeedOpt())
    a = ctx.wait_compact(0, expand_threads=3)
    b = ctx.wait(1)
    _same(a, want[1]); _same(b, want[1])
    ctx.submit(0, bases, off, cuda_lib.SeedOpt())
    mem_off, cm, seed_off, lo, hi = ctx.wait_compact(0)          # the raw wire arrays, decoded here by the header's rules
    assert cm.shape == (want[1].mems.shape[0], 5) and lo.shape[0] == want[1].rbeg.shape[0]
    cm = cm.astype(np.uint64)
    x0 = cm[:, 0] | ((cm[:, 4] & np.uint64(31)) << np.uint64(32))
    x2 = cm[:, 2] | (((cm[:, 4] >> np.uint64(10)) & np.uint64(31)) << np.uint64(32))
    info = ((cm[:, 3] >> np.uint64(16)) << np.uint64(32)) | (cm[:, 3] & np.uint64(0xffff))
    assert np.array_equal(x0, want[1].mems[:, 0]) and np.array_equal(x2, want[1].mems[:, 2]) and np.array_equal(info, want[1].mems[:, 3])
    assert np.array_equal(lo.astype(np.int64) | (hi.astype(np.int64) << 32), want[1].rbeg)
    ctx.close()
    c2 = cuda_lib.SeedContext(idx, n, int(off[-1]), 256, n * 64, n * 600, 1)
    c2.submit(0, bases, off, cuda_lib.SeedOpt())
    with pytest.raises(cuda_lib.CompSeedError):                  # a ctx without compact_results has no compact copy to hand out
        c2.wait_compact(0)
    _same(c2.wait(0), want[1])
    c2.close()
    idx.close()


@pytest.mark.parametrize("form", ["bytes", "packed"])
def test_multi_pipeline_one_device(cuda_lib, workload, form):
    """cs_multi on one GPU: two read sets in flight, batches smaller than the sets, tiny sizing estimates (the slot buffers
    overflow and the ctx is re-created with what cs_ctx_need reports; the block arrays grow) -- same answer as the oracle."""
    _, oi, sets, want = workload
    idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=1)
    for mems_per_read, seeds_per_read, batch in ((0, 0, 4096), (1, 1, 3000)):
        ms = cuda_lib.MultiSeeder([idx], batch_reads=batch, max_read_len=256, n_slots=3, mems_per_read=mems_per_read, seeds_per_read=seeds_per_read)
        for rep in range(2):
            for s, (bases, off) in enumerate(sets):
                o64 = off.astype(np.uint64)
                if form == "bytes":
                    ms.submit(s, bases, o64, cuda_lib.SeedOpt())
                else:
                    pk, nm = cuda_lib.pack_reads_host64(bases, o64, 4)
                    ms.submit_packed(s, pk, nm, o64, cuda_lib.SeedOpt())
            for s in (1, 0):                                     # waited out of order
                r = ms.wait(s)
                _same(r, want[s])
                assert r.info["blocks"] == [(0, 0, want[s].mem_off.shape[0] - 1)]
        assert ms.launches > 0
        ms.close()
    idx.close()


def test_multi_pipeline_errors(cuda_lib, workload):
    _, oi, sets, _ = workload
    idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=1)
    ms = cuda_lib.MultiSeeder([idx], batch_reads=2048, max_read_len=120, n_slots=2)
    bases, off = sets[1]
    ms.submit(0, bases, off.astype(np.uint64), cuda_lib.SeedOpt())
    with pytest.raises(cuda_lib.CompSeedError):      # the set is still in flight
        ms.submit(0, bases, off.astype(np.uint64), cuda_lib.SeedOpt())
    with pytest.raises(cuda_lib.CompSeedError):      # reads longer than max_read_len: reported by the device's worker, not a crash
        ms.wait(0)
    with pytest.raises(cuda_lib.CompSeedError):      # nothing submitted on set 1
        ms.wait(1)
    ms.close()
    idx.close()


def test_multi_pipeline_two_devices(cuda_lib, workload):
    """Index replicated device-to-device, reads split into contiguous 512-aligned blocks in input order, one per GPU, results
    read back through the accessor in input order: equal to the single-GPU answer (north_star: no collective, host gather)."""
    if cuda_lib.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _, oi, sets, want = workload
    idx0 = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=1)
    others = [cuda_lib.replicate_index(idx0, d) for d in range(1, min(cuda_lib.device_count(), 4))]
    ms = cuda_lib.MultiSeeder([idx0] + others, batch_reads=2048, max_read_len=256, n_slots=3)
    for s, (bases, off) in enumerate(sets):
        ms.submit(s, bases, off.astype(np.uint64), cuda_lib.SeedOpt())
    for s in (0, 1):
        r = ms.wait(s)
        _same(r, want[s])
        blocks = r.info["blocks"]
        assert len(blocks) == 1 + len(others) and blocks[0][1] == 0 and blocks[-1][2] == want[s].mem_off.shape[0] - 1
        assert all(b[2] == nb[1] for b, nb in zip(blocks, blocks[1:])) and all(b[1] % 512 == 0 for b in blocks)
        assert [b[0] for b in blocks] == list(range(len(blocks)))
    ms.close()
    # a replica answers the unit-level probes like the original
    k = np.arange(1, 5000, 7, dtype=np.uint64)
    assert np.array_equal(others[0].occ4(k), idx0.occ4(k)) and np.array_equal(others[0].sa(k), idx0.sa(k))
    for i in others + [idx0]:
        i.close()

"""GPU parity tests (run on the B200 box): every call goes through the C-ABI
(include/compseed_b200.h) and is compared bit-exactly with the golden vectors of the unmodified
reference and with the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from compseed_b200 import synth

pytestmark = pytest.mark.gpu


def _opts(g, i):
    o = g[f"opt{i}"]
    return dict(min_seed_len=int(o[0]), split_len=int(o[1]), split_width=int(o[2]), max_mem_intv=int(o[3]), max_occ=int(o[4]))


class _Opt:
    """SeedOpt with an explicit split_len (the goldens store the integer)."""
    def __init__(self, cs, d):
        self.o = cs.SeedOpt(min_seed_len=d["min_seed_len"], split_width=d["split_width"], max_mem_intv=d["max_mem_intv"], max_occ=d["max_occ"])
        self.split_len = d["split_len"]

    def _c(self):
        c = self.o._c()
        c.split_len = self.split_len
        return c


def _upload(cs, g, dense=0):
    return cs.FMIndex.upload(int(g["primary"]), g["L2"], int(g["seq_len"]), g["bwt"], g["sa"], int(g["sa_intv"]), dense_sa_intv=dense)


def _assert_same(r, mem_off, mems, seed_off, rbeg):
    assert np.array_equal(r.mem_off, mem_off)
    assert np.array_equal(r.mems, mems)
    assert np.array_equal(r.seed_off, seed_off)
    assert np.array_equal(r.rbeg, rbeg)


def test_golden_primitives(cuda_lib, golden):
    idx = _upload(cuda_lib, golden)
    assert np.array_equal(idx.occ4(golden["occ_k"]), golden["occ_cnt"])
    assert np.array_equal(idx.extend(golden["ext_ik"], golden["ext_back"]), golden["ext_ok"])
    assert np.array_equal(idx.sa(golden["sa_k"]), golden["sa_v"])


@pytest.mark.parametrize("dense", [0, 4, 1])
def test_golden_seeding(cuda_lib, golden, dense):
    idx = _upload(cuda_lib, golden, dense)
    assert idx.sa_intv == (dense or int(golden["sa_intv"]))
    for i in range(golden["n_opts"]):
        r = cuda_lib.seed_reads(idx, golden["bases"], golden["off"], _Opt(cuda_lib, _opts(golden, i)), batch_reads=128, n_slots=3)
        _assert_same(r, golden[f"mem_off{i}"], golden[f"mems{i}"], golden[f"seed_off{i}"], golden[f"rbeg{i}"])
        assert r.counters["sal_queries"] == golden[f"rbeg{i}"].shape[0]
        if dense == 1:
            assert r.counters["sal_calls"] == 0


def test_index_download_roundtrip(cuda_lib, golden):
    idx = _upload(cuda_lib, golden, dense=1)
    d = idx.download(sa_intv=int(golden["sa_intv"]))
    assert np.array_equal(d["bwt"], golden["bwt"])
    assert np.array_equal(d["sa"], golden["sa"])


@pytest.mark.parametrize("dense", [0, 1])
@pytest.mark.parametrize("kind,n_reads", [("random", 20000), ("repeat", 4000), ("long", 300)])
def test_oracle_parity_seeded(cuda_lib, oracle_lib, kind, n_reads, dense):
    if kind == "random":      # config 1 in miniature
        ref = synth.random_reference(400_000, seed=101)
        bases, off, _ = synth.simulate_reads(ref, n_reads, 150, 0.01, seed=102)
    elif kind == "repeat":    # config 4 in miniature: x[2] > max_occ, 3rd-round reseeding, long SA walks
        ref = synth.repeat_rich_reference(300_000, seed=103, n_segdup=100, segdup_len=2000, n_tandem=60)
        bases, off, _ = synth.simulate_reads(ref, n_reads, [100, 150, 250], 0.02, seed=104, n_rate=0.003)
    else:                     # reads longer than 255 and longer than the shared-memory interval list
        ref = synth.repeat_rich_reference(200_000, seed=105, n_segdup=40, segdup_len=3000, n_tandem=40)
        bases, off, _ = synth.simulate_reads(ref, n_reads, [400, 1000, 3000], 0.01, seed=106, n_rate=0.001)
    oi = oracle_lib.OracleIndex.build(ref)
    # dense = 1 also switches on the unique-match text paths (forward and backward) of k_seed
    idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=dense)
    for r_factor, y, c in [(1.5, 20, 500), (1.0, 20, 50), (2.5, 40, 500)]:
        opt = cuda_lib.SeedOpt(split_factor=r_factor, max_mem_intv=y, max_occ=c)
        want = oi.seed(bases, off, split_len=opt.split_len, max_mem_intv=y, max_occ=c, n_threads=8)
        got = cuda_lib.seed_reads(idx, bases, off, opt, batch_reads=8192)
        _assert_same(got, want.mem_off, want.mems, want.seed_off, want.rbeg)
        # the occurrence filter only ever removes work: queries <= bwamem's bwt_extend call count (SURVEY 8d "E")
        assert got.counters["ext_queries"] <= want.counters["ext"]
        assert got.counters["sal_calls"] == (0 if dense else want.counters["lf"])


def test_batching_and_slots_do_not_change_results(cuda_lib, oracle_lib):
    """Determinism contract: output independent of batch size, slot count, read order."""
    ref = synth.random_reference(100_000, seed=111)
    bases, off, _ = synth.simulate_reads(ref, 3000, [100, 150], 0.01, seed=112, n_rate=0.001)
    oi = oracle_lib.OracleIndex.build(ref)
    idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=1)
    a = cuda_lib.seed_reads(idx, bases, off, batch_reads=3000, n_slots=1)
    b = cuda_lib.seed_reads(idx, bases, off, batch_reads=257, n_slots=4)
    _assert_same(b, a.mem_off, a.mems, a.seed_off, a.rbeg)
    sb, so, perm = synth.shuffle_reads(bases, off)
    c = cuda_lib.seed_reads(idx, sb, so, batch_reads=1000)
    for j in (0, 17, 2999):
        r = int(perm[j])
        assert np.array_equal(c.mems[c.mem_off[j]:c.mem_off[j + 1]], a.mems[a.mem_off[r]:a.mem_off[r + 1]])
        assert np.array_equal(c.rbeg[c.seed_off[j]:c.seed_off[j + 1]], a.rbeg[a.seed_off[r]:a.seed_off[r + 1]])


@pytest.mark.parametrize("dense", [0, 1])
def test_boundary_reads(cuda_lib, oracle_lib, dense):
    """synth.boundary_reads: text start / end of both strands, strand-bridging matches, word-edge substitutions and Ns,
    lengths around multiples of 32 and 255/256 (the oracle is checked against the reference on the same set in
    tests/test_oracle.py)."""
    ref = synth.random_reference(60_000, seed=51)
    bases, off = synth.boundary_reads(ref)
    oi = oracle_lib.OracleIndex.build(ref)
    idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=dense)
    for opt in (cuda_lib.SeedOpt(), cuda_lib.SeedOpt(split_factor=1.0, max_mem_intv=40, max_occ=50)):
        want = oi.seed(bases, off, split_len=opt.split_len, max_mem_intv=opt.max_mem_intv, max_occ=opt.max_occ, n_threads=4)
        got = cuda_lib.seed_reads(idx, bases, off, opt, batch_reads=500)
        _assert_same(got, want.mem_off, want.mems, want.seed_off, want.rbeg)
    idx.close()


@pytest.mark.parametrize("dense", [0, 1])
def test_packed_submission_equals_byte_submission(cuda_lib, golden, dense):
    """cs_seed_batch_submit_packed (reads 2-bit packed by the host) == cs_seed_batch_submit == golden."""
    idx = _upload(cuda_lib, golden, dense)
    bases, off = golden["bases"], golden["off"]
    n = off.shape[0] - 1
    packed, nmask = cuda_lib.pack_reads(bases, off)
    ctx = cuda_lib.SeedContext(idx, n, int(off[-1]), 256, n * 64, n * 600, 2)
    opt = _Opt(cuda_lib, _opts(golden, 0))
    ctx.submit_packed(0, packed, nmask, off, opt)
    ctx.submit(1, bases, off, opt)
    a, b = ctx.wait(0), ctx.wait(1)
    _assert_same(a, b.mem_off, b.mems, b.seed_off, b.rbeg)
    _assert_same(a, golden["mem_off0"], golden["mems0"], golden["seed_off0"], golden["rbeg0"])
    ctx.submit_packed(1, packed, nmask, off, opt)       # the other slot, after it held a byte batch
    _assert_same(ctx.wait(1), golden["mem_off0"], golden["mems0"], golden["seed_off0"], golden["rbeg0"])
    ctx.close()


def test_slot_reuse_across_different_batches(cuda_lib, oracle_lib):
    """A slot that has held another batch (stale deferred-call queue, scratch, chains) gives the same answer as a
    fresh context.  (Regression: k_seed_walk once scanned queue entries that a concurrent lane had reserved but not
    yet written, and could pick up a stale task of the previous batch.)"""
    ref = synth.random_reference(300_000, seed=601)
    sets = [synth.simulate_reads(ref, 6000, [100, 150, 250], 0.02, seed=602 + i, n_rate=0.002)[:2] for i in range(3)]
    oi = oracle_lib.OracleIndex.build(ref)
    idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=1)
    want = [oi.seed(b, o, n_threads=8) for b, o in sets]
    n_max = max(o.shape[0] - 1 for _, o in sets)
    ctx = cuda_lib.SeedContext(idx, n_max, max(int(o[-1]) for _, o in sets), 256, n_max * 64, n_max * 600, 1)
    for rep in range(4):
        for (b, o), w in zip(sets, want):
            ctx.submit(0, b, o, cuda_lib.SeedOpt())
            r = ctx.wait(0)
            _assert_same(r, w.mem_off, w.mems, w.seed_off, w.rbeg)
    ctx.close()


def test_fast_route_equals_literal_route_midsize(cuda_lib):
    """100 Mbp reference (too large for the CPU oracle in test time): the fast route (k_seed_fast, k_seed_walk,
    k_seed in call mode, k_seed_r3_fast) against the literal route (k_seed in read mode, k_seed_r3), which the
    tests above pin to the oracle; every mem and seed position.  The index itself is checked by cs_index_verify."""
    ref = synth.random_reference(100_000_000, seed=701)
    bases, off, _ = synth.simulate_reads(ref, 400_000, [100, 150, 250], 0.01, seed=702, n_rate=0.0005)
    idx = cuda_lib.FMIndex.build(ref, sa_intv=1)
    v = idx.verify(ref, stride=1)
    assert v["ok"] and v["order_rows"] == idx.seq_len and v["text_bases"] == idx.seq_len, v
    out = {}
    for fast in (1, 0):
        out[fast] = cuda_lib.seed_reads(idx, bases, off, batch_reads=150_001, n_slots=2, config=cuda_lib.CtxConfig(use_fast=fast))
    a, b = out[1], out[0]
    assert a.counters["deferred_calls"] > 0 and b.counters["deferred_calls"] == 0
    _assert_same(a, b.mem_off, b.mems, b.seed_off, b.rbeg)
    # the third pass on its own stream next to the walk / literal kernels, or after them on the same stream: same answer
    c = cuda_lib.seed_reads(idx, bases, off, batch_reads=150_001, n_slots=2, config=cuda_lib.CtxConfig(overlap_streams=1))
    _assert_same(c, a.mem_off, a.mems, a.seed_off, a.rbeg)
    idx.close()


def test_staged_device_resident_run_and_fetch(cuda_lib, golden):
    idx = _upload(cuda_lib, golden, dense=1)
    n = golden["off"].shape[0] - 1
    ctx = cuda_lib.SeedContext(idx, n, int(golden["off"][-1]), 256, n * 64, n * 600, 1)
    ctx.stage(0, golden["bases"], golden["off"])
    opt = _Opt(cuda_lib, _opts(golden, 0))
    for _ in range(2):   # stage once, run twice: same answer
        ctx.run_staged(0, opt)
        d = ctx.wait_device(0)
        assert d.n_mems_device == golden["mems0"].shape[0] and d.n_seeds_device == golden["rbeg0"].shape[0]
        r = ctx.fetch(0)
        _assert_same(r, golden["mem_off0"], golden["mems0"], golden["seed_off0"], golden["rbeg0"])
        assert r.kernel_ms[0] > 0
    ctx.close()


def test_overflow_is_reported_not_truncated(cuda_lib, golden):
    if golden["name"] != "repeat30k":
        pytest.skip("needs the repeat-rich fixture")
    idx = _upload(cuda_lib, golden)
    n = golden["off"].shape[0] - 1
    ctx = cuda_lib.SeedContext(idx, n, int(golden["off"][-1]), 256, n * 16, n * 4, 1)   # seeds cannot fit
    ctx.submit(0, golden["bases"], golden["off"], _Opt(cuda_lib, _opts(golden, 0)))
    with pytest.raises(cuda_lib.CompSeedError) as e:
        ctx.wait(0)
    assert e.value.code == -3
    ctx.close()


def test_argument_errors(cuda_lib, golden):
    idx = _upload(cuda_lib, golden)
    ctx = cuda_lib.SeedContext(idx, 16, 1024, 100, 0, 0, 1)
    with pytest.raises(cuda_lib.CompSeedError):   # read longer than max_read_len
        ctx.submit(0, np.zeros(200, np.uint8), np.array([0, 200], np.uint32), cuda_lib.SeedOpt())
    with pytest.raises(cuda_lib.CompSeedError):   # wait without submit
        ctx.wait(0)
    with pytest.raises(cuda_lib.CompSeedError):   # bad slot
        ctx.submit(3, np.zeros(10, np.uint8), np.array([0, 10], np.uint32), cuda_lib.SeedOpt())
    ctx.close()


def test_random_gather_probe_runs(cuda_lib):
    gb, gl = cuda_lib.probe_random_gather(0, 256 << 20, 32, 1 << 22, 1)
    assert gb > 0 and gl > 0


@pytest.mark.parametrize("kind", ["random", "repeat", "tiny", "homopolymer"])
def test_gpu_index_build_matches_bwaidx_semantics(cuda_lib, oracle_lib, kind):
    """cs_index_build (device suffix sort) == the oracle builder, which is pinned to bwaidx by the goldens."""
    if kind == "random":
        ref = synth.random_reference(300_000, seed=201)
    elif kind == "repeat":   # long ties: tandem repeats and exact duplications force many refinement rounds
        ref = synth.repeat_rich_reference(200_000, seed=202, n_segdup=60, segdup_len=2500, n_tandem=40, divergence=0.0)
    elif kind == "tiny":
        ref = synth.random_reference(37, seed=203)
    else:                    # worst case for ties and for the '$'-is-smallest rule
        ref = np.zeros(5000, dtype=np.uint8)
        ref[2500:] = 3
    oi = oracle_lib.OracleIndex.build(ref)
    for intv in (32, 1):
        idx = cuda_lib.FMIndex.build(ref, sa_intv=intv)
        assert idx.primary == oi.primary and np.array_equal(idx.L2, oi.L2) and idx.seq_len == oi.seq_len
        d = idx.download(sa_intv=32)
        assert np.array_equal(d["bwt"], oi.bwt)
        assert np.array_equal(d["sa"], oi.sa)
        idx.close()


def test_golden_reference_rebuilt_on_gpu(cuda_lib, golden):
    """Index built on the GPU from the golden reference sequence == the index bwaidx wrote."""
    idx = cuda_lib.FMIndex.build(golden["ref"], sa_intv=int(golden["sa_intv"]))
    d = idx.download(sa_intv=int(golden["sa_intv"]))
    assert idx.primary == int(golden["primary"])
    assert np.array_equal(d["bwt"], golden["bwt"]) and np.array_equal(d["sa"], golden["sa"])
    r = cuda_lib.seed_reads(idx, golden["bases"], golden["off"], _Opt(cuda_lib, _opts(golden, 0)))
    _assert_same(r, golden["mem_off0"], golden["mems0"], golden["seed_off0"], golden["rbeg0"])


def test_result_neutral_caches_can_be_switched_off(cuda_lib, oracle_lib):
    """Top-of-search table and occurrence filter are pure caches/filters: with both disabled the kernel
    issues exactly bwamem's bwt_extend calls (E of SURVEY 8d) and still returns the same mems.  Every switch of
    cs_index_config_t / cs_ctx_config_t is exercised."""
    ref = synth.random_reference(250_000, seed=501)
    bases, off, _ = synth.simulate_reads(ref, 4000, [100, 150], 0.015, seed=502, n_rate=0.002)
    oi = oracle_lib.OracleIndex.build(ref)
    want = oi.seed(bases, off, n_threads=8)
    IC, CC = cuda_lib.IndexConfig, cuda_lib.CtxConfig
    results = {}
    for name, (dense, icfg, ccfg) in {
            "plain": (0, IC(kmer_table_depth=0, prune_k=0), CC()), "table": (0, IC(prune_k=0), CC()),
            "filter": (0, IC(kmer_table_depth=0), CC()), "both": (0, IC(), CC()), "deep": (0, IC(kmer_table_depth=11, prune_k=16), CC()),
            # dense SA => unique-match text paths on; they must account for exactly the extends they replace
            "text": (1, IC(prune_k=0), CC()), "text_isa1": (1, IC(prune_k=0, isa_intv=1), CC()),
            "text_isa32": (1, IC(kmer_table_depth=0, prune_k=0, isa_intv=32), CC()),
            "all_dense": (1, IC(), CC()), "all_dense_tiny_queue": (1, IC(), CC(defer_cap=7)), "all_dense_nofast": (1, IC(), CC(use_fast=0)),
            "all_dense_r3slow": (1, IC(), CC(use_r3_fast=0)), "all_dense_k12": (1, IC(prune_k=12, kmer_table_depth=9), CC()),
            "all_dense_serial": (1, IC(), CC(overlap_streams=1)), "all_dense_r3slow_serial": (1, IC(), CC(use_r3_fast=0, overlap_streams=1)),
            "all_dense_isa4": (1, IC(isa_intv=4), CC()), "all_dense_norep": (1, IC(repeat_lengths=0), CC()),
            "all_dense_l2window": (1, IC(), CC(l2_persist_mb=16)), "all_dense_litcap": (1, IC(), CC(lit_ctas_per_sm=1)),
            "all_dense_prefetch": (1, IC(), CC(prefetch_results=1))}.items():
        idx = cuda_lib.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=dense, config=icfg)
        got = cuda_lib.seed_reads(idx, bases, off, batch_reads=2048, config=ccfg)
        _assert_same(got, want.mem_off, want.mems, want.seed_off, want.rbeg)
        results[name] = got.counters
        idx.close()
    assert results["plain"]["ext_queries"] == want.counters["ext"] == results["plain"]["ext_calls"] + (results["plain"]["ext_queries"] - results["plain"]["ext_calls"])
    assert results["table"]["ext_queries"] == want.counters["ext"] and results["table"]["ext_calls"] < results["plain"]["ext_calls"]
    assert results["filter"]["ext_queries"] < want.counters["ext"]
    assert results["both"]["ext_calls"] < results["table"]["ext_calls"]
    for name in ("text", "text_isa1", "text_isa32"):
        assert results[name]["ext_queries"] == want.counters["ext"], name
        assert results[name]["ext_calls"] < results["plain"]["ext_calls"], name
    assert results["all_dense"]["ext_queries"] <= results["all_dense_nofast"]["ext_queries"] <= results["both"]["ext_queries"]
    assert results["all_dense"]["deferred_calls"] > 7                 # the fast kernel ran and handed calls over
    assert results["all_dense_nofast"]["deferred_calls"] == 0
    # a queue too small for the batch: rerun through the literal kernel alone, same answer (checked above)
    assert results["all_dense_tiny_queue"]["ext_queries"] == results["all_dense_nofast"]["ext_queries"]
    assert results["all_dense_serial"] == results["all_dense"]
    # repeat lengths: second-pass calls inside one-occurrence SMEMs read neither the FM-index nor the filter
    assert results["all_dense"]["ext_calls"] < results["all_dense_norep"]["ext_calls"]

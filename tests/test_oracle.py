"""CPU tests: the oracle (oracle/cs_oracle.c) against the golden vectors produced by the unmodified
reference, and -- when oracle/_ref is present -- against the reference itself, live."""
import os

import numpy as np
import pytest

from compseed_b200 import synth


def _opts(g, i):
    o = g[f"opt{i}"]
    return dict(min_seed_len=int(o[0]), split_len=int(o[1]), split_width=int(o[2]), max_mem_intv=int(o[3]), max_occ=int(o[4]))


def test_index_build_matches_bwaidx(oracle_lib, golden):
    idx = oracle_lib.OracleIndex.build(golden["ref"], sa_intv=int(golden["sa_intv"]))
    assert idx.primary == int(golden["primary"])
    assert np.array_equal(idx.L2, golden["L2"])
    assert idx.seq_len == int(golden["seq_len"])
    assert np.array_equal(idx.bwt, golden["bwt"])
    assert np.array_equal(idx.sa, golden["sa"])


def test_primitives_match_reference(oracle_lib, golden):
    idx = oracle_lib.OracleIndex.build(golden["ref"])
    assert np.array_equal(idx.occ4(golden["occ_k"]), golden["occ_cnt"])
    assert np.array_equal(idx.extend(golden["ext_ik"], golden["ext_back"]), golden["ext_ok"])
    assert np.array_equal(idx.sa_lookup(golden["sa_k"]), golden["sa_v"])


def test_seeding_matches_reference(oracle_lib, golden):
    idx = oracle_lib.OracleIndex.build(golden["ref"])
    for i in range(golden["n_opts"]):
        r = idx.seed(golden["bases"], golden["off"], n_threads=2, **_opts(golden, i))
        assert np.array_equal(r.mem_off, golden[f"mem_off{i}"])
        assert np.array_equal(r.mems, golden[f"mems{i}"])
        assert np.array_equal(r.seed_off, golden[f"seed_off{i}"])
        assert np.array_equal(r.rbeg, golden[f"rbeg{i}"])
        assert r.counters["mem"] == golden[f"mems{i}"].shape[0]
        assert r.counters["sa"] == golden[f"rbeg{i}"].shape[0]


def test_index_file_roundtrip(oracle_lib, golden, tmp_path):
    idx = oracle_lib.OracleIndex.build(golden["ref"])
    p = str(tmp_path / "idx")
    idx.dump(p)
    assert os.path.getsize(p + ".bwt") == 40 + 4 * idx.bwt_size
    assert os.path.getsize(p + ".sa") == 56 + 8 * (idx.n_sa - 1)
    back = oracle_lib.OracleIndex.load(p)
    assert back.primary == idx.primary and np.array_equal(back.bwt, idx.bwt) and np.array_equal(back.sa, idx.sa)
    if oracle_lib.have_ref():  # the reference's own loader accepts our files
        ri = oracle_lib.RefIndex.load(p)
        assert ri.primary == idx.primary and np.array_equal(ri.bwt, idx.bwt) and np.array_equal(ri.sa, idx.sa)


def test_sa_sampling_does_not_change_bwt_sa(oracle_lib, golden):
    a = oracle_lib.OracleIndex.build(golden["ref"], sa_intv=32)
    b = oracle_lib.OracleIndex.build(golden["ref"], sa_intv=1)
    k = golden["sa_k"]
    assert np.array_equal(a.sa_lookup(k), b.sa_lookup(k))
    r32 = a.seed(golden["bases"], golden["off"], **_opts(golden, 0))
    r1 = b.seed(golden["bases"], golden["off"], **_opts(golden, 0))
    assert r32.same_as(r1) and r1.counters["lf"] == 0 and r32.counters["lf"] > 0


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libcsref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind", ["random", "repeat"])
def test_live_against_reference(oracle_lib, kind):
    if kind == "random":
        ref = synth.random_reference(150_000, seed=31)
        bases, off, _ = synth.simulate_reads(ref, 3000, [100, 150, 250], 0.015, seed=32, n_rate=0.001)
    else:
        ref = synth.repeat_rich_reference(120_000, seed=41, n_segdup=60, segdup_len=1500, n_tandem=30)
        bases, off, _ = synth.simulate_reads(ref, 2000, [100, 150, 250], 0.02, seed=42, n_rate=0.002)
    oi = oracle_lib.OracleIndex.build(ref)
    ri = oracle_lib.RefIndex.from_arrays(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv)
    for r in (1.0, 1.5, 2.5):
        a = oi.seed(bases, off, split_len=synth.split_len_bwamem(19, r), n_threads=4)
        b = ri.seed(bases, off, "bwamem", split_factor=r, n_threads=4)
        c = ri.seed(bases, off, "compseed", split_factor=r, n_threads=2)
        assert a.same_as(b), f"oracle != bwt_smem1 path at -r {r}"
        assert a.same_as(c), f"oracle != CompSeed SST path at -r {r}"
        # E is bwamem's bwt_extend call count; CompSeed's own query counter (comp_seed.cpp:81,123,151) differs
        # from it only by a handful of calls on reads containing N
        assert abs(a.counters["ext"] - c.counters["ext_queries"]) <= 1e-4 * a.counters["ext"]
        assert a.counters["sa"] == c.counters["sal_queries"]


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libcsref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_boundary_reads_against_reference(oracle_lib):
    """Text start / end of both strands, strand-bridging matches, lengths around 32-multiples and 255/256."""
    ref = synth.random_reference(60_000, seed=51)
    bases, off = synth.boundary_reads(ref)
    oi = oracle_lib.OracleIndex.build(ref)
    ri = oracle_lib.RefIndex.from_arrays(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv)
    for kw in (dict(), dict(min_seed_len=19, split_factor=1.0, max_mem_intv=40, max_occ=50)):
        a = oi.seed(bases, off, min_seed_len=kw.get("min_seed_len", 19), split_len=synth.split_len_bwamem(19, kw.get("split_factor", 1.5)),
                    max_mem_intv=kw.get("max_mem_intv", 20), max_occ=kw.get("max_occ", 500), n_threads=4)
        b = ri.seed(bases, off, "bwamem", n_threads=4, **kw)
        c = ri.seed(bases, off, "compseed", n_threads=2, **kw)
        assert a.same_as(b) and a.same_as(c)
    assert a.mems.shape[0] > off.shape[0] - 1        # the set does produce seeds (most reads match end to end)


def test_split_len_rounding():
    assert synth.split_len_bwamem(19, 1.5) == 28 and synth.split_len_compseed(19, 1.5) == 28
    assert synth.split_len_bwamem(19, 1.0) == 19 and synth.split_len_bwamem(19, 2.5) == 47


def test_synth_reads_are_reproducible():
    ref = synth.random_reference(5000, seed=1)
    a = synth.simulate_reads(ref, 50, 100, 0.01, seed=9)
    b = synth.simulate_reads(ref, 50, 100, 0.01, seed=9)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.all(np.diff(a[2]) >= 0)              # position-sorted (SPRING-like order)
    sb, so, perm = synth.shuffle_reads(a[0], a[1])
    r = int(perm[3])
    assert np.array_equal(sb[so[3]:so[4]], a[0][a[1][r]:a[1][r + 1]])


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libcsref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_cfg1_full_size_oracle_against_reference(oracle_lib, tmp_path):
    """BASELINE.json configs[0] at its stated size (5 Mbp, 200 k position-sorted 150-bp reads, defaults): the oracle's
    index == bwaidx's files, and its mems / seeds == both seeding paths of the unmodified reference."""
    ref = synth.random_reference(5_000_000, seed=20261018)
    bases, off, _ = synth.simulate_reads(ref, 200_000, 150, 0.01, seed=1)
    d = str(tmp_path)
    synth.write_fasta(os.path.join(d, "ref.fa"), ref)
    oracle_lib.bwaidx(os.path.join(d, "ref.fa"), os.path.join(d, "ref"))
    ri = oracle_lib.RefIndex.load(os.path.join(d, "ref"))
    oi = oracle_lib.OracleIndex.build(ref)
    assert oi.primary == ri.primary and np.array_equal(oi.bwt, ri.bwt) and np.array_equal(oi.sa, ri.sa)
    a = oi.seed(bases, off, n_threads=8)
    b = ri.seed(bases, off, "bwamem", n_threads=8)
    c = ri.seed(bases, off, "compseed", n_threads=8)
    assert a.same_as(b) and a.same_as(c)
    assert a.mems.shape[0] > 1_000_000

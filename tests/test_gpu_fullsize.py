"""GPU parity at the sizes BASELINE.json states (run on the B200 box), against the UNMODIFIED reference
(oracle/_ref: bwaidx, bwt_restore_*, collect_mem_with_sst / bwt_smem1 / bwt_sa through libcsref.so):

  cfg1  5 Mbp reference, 200 k position-sorted 150-bp reads, default options          -> both reference seeding paths
  cfg4  20 Mbp repeat-rich reference, 50 k reads                                       -> the oracle (pinned to the reference)
  rows >= 2^32 (a 2.2 Gbp reference, 4.4 G rows)                                      -> the reference's own bwt_occ4 /
        bwt_extend / bwt_sa and both seeding paths, on rows the 32-bit halves of the device layout cannot hold
  the on-disk format: what cs_index_write writes is what bwaidx writes, and bwt_restore_bwt / bwt_restore_sa load it

bench.py does the same comparison on the 3.1 Gbp index of the headline line (`parity` in its JSON)."""
import filecmp
import os

import numpy as np
import pytest

from compseed_b200 import synth

pytestmark = pytest.mark.gpu


def _assert_same(r, w):
    assert np.array_equal(r.mem_off, w.mem_off)
    assert np.array_equal(r.mems, w.mems)
    assert np.array_equal(r.seed_off, w.seed_off)
    assert np.array_equal(r.rbeg, w.rbeg)


def test_cfg1_full_size_against_the_reference(cuda_lib, oracle_lib, tmp_path):
    """BASELINE.json configs[0] at its stated size, end to end through the reference's own files: bwaidx writes the
    index, cs_index_load reads it, and the mems / seed positions equal those of CompSeed's SST path and of bwamem's."""
    if not oracle_lib.have_ref():
        pytest.skip("needs oracle/_ref (the unmodified reference, built where /root/reference exists)")
    d = str(tmp_path)
    ref = synth.random_reference(5_000_000, seed=20261018)
    bases, off, _ = synth.simulate_reads(ref, 200_000, 150, 0.01, seed=1)      # position-sorted (SPRING-like order)
    synth.write_fasta(os.path.join(d, "ref.fa"), ref)
    oracle_lib.bwaidx(os.path.join(d, "ref.fa"), os.path.join(d, "ref"))
    ri = oracle_lib.RefIndex.load(os.path.join(d, "ref"))
    want_cs = ri.seed(bases, off, "compseed", n_threads=os.cpu_count() or 4)
    want_bw = ri.seed(bases, off, "bwamem", n_threads=os.cpu_count() or 4)
    assert want_cs.same_as(want_bw)
    for dense in (1, 0):
        idx = cuda_lib.FMIndex.load(os.path.join(d, "ref"), dense_sa_intv=dense)
        got = cuda_lib.seed_reads(idx, bases, off, cuda_lib.SeedOpt(caller="compseed"), batch_reads=65_536, n_slots=3)
        _assert_same(got, want_cs)
        if dense:
            assert idx.verify(ref)["ok"]
        idx.close()
    # shuffled reads: same per-read answer (the SST of the reference and every structure here are result-neutral)
    sb, so, perm = synth.shuffle_reads(bases, off)
    idx = cuda_lib.FMIndex.build(ref, sa_intv=1)
    got = cuda_lib.seed_reads(idx, sb, so, batch_reads=100_000)
    lens_w = np.diff(want_cs.mem_off.astype(np.int64))
    assert np.array_equal(np.diff(got.mem_off.astype(np.int64)), lens_w[perm])
    for j in range(0, 200_000, 997):
        r = int(perm[j])
        assert np.array_equal(got.mems[got.mem_off[j]:got.mem_off[j + 1]], want_cs.mems[want_cs.mem_off[r]:want_cs.mem_off[r + 1]])
        assert np.array_equal(got.rbeg[got.seed_off[j]:got.seed_off[j + 1]], want_cs.rbeg[want_cs.seed_off[r]:want_cs.seed_off[r + 1]])
    # the index built on the GPU, written in the reference's format: the same bytes bwaidx wrote, loadable by bwt_restore_*
    idx.write(os.path.join(d, "gpu"), sa_intv=32)
    assert filecmp.cmp(os.path.join(d, "gpu.bwt"), os.path.join(d, "ref.bwt"), shallow=False)
    assert filecmp.cmp(os.path.join(d, "gpu.sa"), os.path.join(d, "ref.sa"), shallow=False)
    rg = oracle_lib.RefIndex.load(os.path.join(d, "gpu"))
    assert rg.primary == ri.primary and np.array_equal(rg.bwt, ri.bwt) and np.array_equal(rg.sa, ri.sa)
    idx.close()


def test_index_files_are_validated(cuda_lib, tmp_path):
    """Truncated / corrupt index files give CS_E_IO, not a crash (bwt_restore_* would abort or read garbage)."""
    d = str(tmp_path)
    ref = synth.random_reference(3000, seed=9)
    idx = cuda_lib.FMIndex.build(ref, sa_intv=32)
    idx.write(os.path.join(d, "x"), sa_intv=32)
    idx.close()
    good_bwt = open(os.path.join(d, "x.bwt"), "rb").read()
    good_sa = open(os.path.join(d, "x.sa"), "rb").read()
    cuda_lib.FMIndex.load(os.path.join(d, "x")).close()

    def expect_io(bwt, sa):
        open(os.path.join(d, "y.bwt"), "wb").write(bwt)
        open(os.path.join(d, "y.sa"), "wb").write(sa)
        with pytest.raises(cuda_lib.CompSeedError) as e:
            cuda_lib.FMIndex.load(os.path.join(d, "y"))
        assert e.value.code == -4, e.value

    expect_io(good_bwt[:20], good_sa)                                   # shorter than the header
    expect_io(good_bwt[:-8], good_sa)                                   # truncated body
    expect_io(good_bwt, good_sa[:40])                                   # SA header cut
    expect_io(good_bwt, good_sa[:-16])                                  # SA body cut
    expect_io(good_bwt, good_sa[:40] + (0).to_bytes(8, "little") + good_sa[48:])      # sa_intv 0
    expect_io(good_bwt, good_sa[:40] + (24).to_bytes(8, "little") + good_sa[48:])     # not a power of two
    expect_io(good_bwt, (7).to_bytes(8, "little") + good_sa[8:])        # primary disagrees with the .bwt


def test_cfg4_repeat_rich_20mbp_against_the_oracle(cuda_lib, oracle_lib):
    """BASELINE.json configs[3] at 20 Mbp: segmental duplications + tandem repeats (x[2] > -c, third-pass reseeding, the
    literal kernel carries most calls), 50 k reads against the oracle on the index the GPU built."""
    ref = synth.repeat_rich_reference(20_000_000, seed=401, n_segdup=200, segdup_len=5000, n_tandem=400)
    bases, off, _ = synth.simulate_reads(ref, 50_000, [100, 150, 250], 0.02, seed=402, n_rate=0.001)
    idx = cuda_lib.FMIndex.build(ref, sa_intv=1)
    v = idx.verify(ref)
    assert v["ok"], v
    h = idx.download(sa_intv=32)
    oi = oracle_lib.OracleIndex.from_arrays(h["primary"], h["L2"], h["seq_len"], h["bwt"], h["sa"], h["sa_intv"])
    for opt in (cuda_lib.SeedOpt(), cuda_lib.SeedOpt(split_factor=1.0, max_mem_intv=40, max_occ=50)):
        want = oi.seed(bases, off, split_len=opt.split_len, max_mem_intv=opt.max_mem_intv, max_occ=opt.max_occ, n_threads=os.cpu_count() or 4)
        got = cuda_lib.seed_reads(idx, bases, off, opt, batch_reads=20_000, n_slots=2)
        _assert_same(got, want)
        assert got.counters["deferred_calls"] > 0
    if oracle_lib.have_ref():      # and the reference itself, on a slice (its SST path is slow on repeats)
        ri = oracle_lib.RefIndex.from_arrays(h["primary"], h["L2"], h["seq_len"], h["bwt"], h["sa"], h["sa_intv"])
        n = 10_000
        w = ri.seed(bases[:int(off[n])], off[:n + 1], "compseed", n_threads=os.cpu_count() or 4)
        g = cuda_lib.seed_reads(idx, bases[:int(off[n])], off[:n + 1], cuda_lib.SeedOpt(caller="compseed"), batch_reads=4096)
        _assert_same(g, w)
    idx.close()


def test_rows_beyond_2_to_the_32(cuda_lib, oracle_lib):
    """A 2.2 Gbp reference: 4.4 G BWT rows, so rows, SA values and text positions need more than 32 bits -- the hi byte
    of the 40-bit checkpoints, the 37-bit table / list packing, 64-bit SA entries.  Primitives and seeding against the
    reference's own code on the index downloaded from the GPU; the index itself against its definitions."""
    import torch
    l_pac = 2_200_000_000
    ref_t = synth.random_reference_torch(l_pac, 4242, "cuda:0")
    bases, off, _ = synth.simulate_reads_torch(ref_t, 40_000, 150, 0.01, seed=77)
    ref = ref_t.cpu().numpy()
    del ref_t
    torch.cuda.empty_cache()
    idx = cuda_lib.FMIndex.build(ref, sa_intv=1)
    assert idx.seq_len == 2 * l_pac > (1 << 32)
    v = idx.verify(ref, stride=1)
    assert v["ok"] and v["order_rows"] == idx.seq_len and v["bwt_rows"] == idx.seq_len + 1, v
    del ref
    h = idx.download(sa_intv=32)
    if oracle_lib.have_ref():
        ri = oracle_lib.RefIndex.from_arrays(h["primary"], h["L2"], h["seq_len"], h["bwt"], h["sa"], h["sa_intv"])
        want = ri.seed(bases, off, "compseed", n_threads=os.cpu_count() or 4)
        want_bw = ri.seed(bases, off, "bwamem", n_threads=os.cpu_count() or 4)
        assert want.same_as(want_bw)
    else:
        ri = oracle_lib.OracleIndex.from_arrays(h["primary"], h["L2"], h["seq_len"], h["bwt"], h["sa"], h["sa_intv"])
        want = ri.seed(bases, off, n_threads=os.cpu_count() or 4)
    got = cuda_lib.seed_reads(idx, bases, off, cuda_lib.SeedOpt(caller="compseed"), batch_reads=16_384, n_slots=2)
    _assert_same(got, want)
    hi = want.mems[:, 0] >= np.uint64(1 << 32)
    assert hi.sum() > 1000 and (want.rbeg >= (1 << 32)).sum() > 1000          # the regime is really exercised
    lit = cuda_lib.seed_reads(idx, bases, off, cuda_lib.SeedOpt(caller="compseed"), batch_reads=16_384, config=cuda_lib.CtxConfig(use_fast=0))
    _assert_same(lit, want)
    # primitives on rows above 2^32
    rng = np.random.default_rng(3)
    k = np.concatenate([rng.integers(1 << 32, idx.seq_len + 1, 4000, dtype=np.uint64),
                        np.array([(1 << 32) - 1, 1 << 32, (1 << 32) + 1, idx.primary - 1, idx.primary, idx.primary + 1, idx.seq_len - 1, idx.seq_len], dtype=np.uint64)])
    assert np.array_equal(idx.occ4(k), ri.occ4(k))
    assert np.array_equal(idx.sa(k), ri.sa_lookup(k))
    ik = np.ascontiguousarray(want.mems[hi][:3000, :3])
    for back in (0, 1):
        flags = np.full(ik.shape[0], back, np.int32)
        assert np.array_equal(idx.extend(ik, flags), ri.extend(ik, flags))
    idx.close()


def test_overflow_tells_the_capacities_it_needs(cuda_lib, golden):
    """CS_E_OVERFLOW names what the batch needs (cs_ctx_need): one re-creation of the ctx is enough."""
    if golden["name"] != "repeat30k":
        pytest.skip("needs the repeat-rich fixture")
    idx = cuda_lib.FMIndex.upload(int(golden["primary"]), golden["L2"], int(golden["seq_len"]), golden["bwt"], golden["sa"], int(golden["sa_intv"]), dense_sa_intv=1)
    n = golden["off"].shape[0] - 1
    opt = cuda_lib.SeedOpt()
    o = golden["opt0"]
    assert (opt.min_seed_len, opt.split_len, opt.split_width, opt.max_mem_intv, opt.max_occ) == tuple(int(x) for x in o)
    for cap_m, cap_s in ((n * 64, n * 2), (n * 2, n * 600), (n, n)):     # seeds too small / mems too small / both
        ctx = cuda_lib.SeedContext(idx, n, int(golden["off"][-1]), 256, cap_m, cap_s, 1)
        ctx.submit(0, golden["bases"], golden["off"], opt)
        with pytest.raises(cuda_lib.CompSeedError) as e:
            ctx.wait(0)
        assert e.value.code == -3
        need_m, need_s = ctx.need(0)
        ctx.close()
        assert need_m >= min(cap_m, golden["mems0"].shape[0]) and need_s >= min(cap_s, golden["rbeg0"].shape[0])
        ctx = cuda_lib.SeedContext(idx, n, int(golden["off"][-1]), 256, need_m, need_s, 1)
        ctx.submit(0, golden["bases"], golden["off"], opt)
        try:
            r = ctx.wait(0)
        except cuda_lib.CompSeedError as e2:   # the pool overflowed: the seed count was an estimate; the second answer is exact
            assert e2.code == -3
            need_m, need_s = ctx.need(0)
            ctx.close()
            ctx = cuda_lib.SeedContext(idx, n, int(golden["off"][-1]), 256, need_m, need_s, 1)
            ctx.submit(0, golden["bases"], golden["off"], opt)
            r = ctx.wait(0)
        assert np.array_equal(r.mems, golden["mems0"]) and np.array_equal(r.rbeg, golden["rbeg0"])
        assert ctx.launches >= 9
        ctx.close()
    idx.close()

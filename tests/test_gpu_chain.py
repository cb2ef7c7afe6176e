"""Chaining + chain filtering on the GPU (cs_ctx_set_chaining / cs_seed_batch_wait_chains) against the
reference's mem_chain + mem_chain_flt (comp_seed.cpp:241-354 == bwamem.c:359-497): every chain, its weight, kept flag,
seeds and order, bit for bit.  tests/test_chain_emul.py runs the same kernels' source on the CPU against the same oracle."""
from types import SimpleNamespace

import numpy as np
import pytest

from compseed_b200 import synth
from test_chain_emul import assert_chains_equal

pytestmark = pytest.mark.gpu


def _as_dict(c):
    return dict(chain_off=c.chain_off, cseed_off=c.cseed_off, rid=c.rid, w=c.w, kept=c.kept, n=c.n, l_rep=c.l_rep, s_rbeg=c.s_rbeg, s_qbeg=c.s_qbeg, s_len=c.s_len)


def test_golden_chains(cuda_lib, golden):
    idx = cuda_lib.FMIndex.upload(int(golden["primary"]), golden["L2"], int(golden["seq_len"]), golden["bwt"], golden["sa"], int(golden["sa_intv"]), dense_sa_intv=1)
    n = golden["off"].shape[0] - 1
    ctx = cuda_lib.SeedContext(idx, n, int(golden["off"][-1]), 256, n * 64, n * 600, 2)
    ctx.set_chaining([golden["ref"].shape[0]])
    want = SimpleNamespace(**{k: golden["chain_" + k + "0"] for k in ("pos", "rid", "w", "kept", "n", "s_rbeg", "s_qbeg", "s_len", "frac_rep")}, chain_off=golden["chain_off0"])
    opt = cuda_lib.SeedOpt()
    for slot in (0, 1, 0):
        ctx.submit(slot, golden["bases"], golden["off"], opt)
        got = ctx.wait_chains(slot)
        assert_chains_equal(_as_dict(got), want, golden["off"])
    # the mems and seed positions of a chained batch are still there for whoever wants them
    r = ctx.fetch(0)
    assert np.array_equal(r.mems, golden["mems0"]) and np.array_equal(r.rbeg, golden["rbeg0"])
    if golden["name"] == "random20k":   # ordinary reads: far fewer bytes than mems + seed positions (repeat-rich reads have a chain for almost every seed)
        assert got.wire_bytes < (32 * r.mems.shape[0] + 8 * r.rbeg.shape[0] + 8 * n) // 2
    ctx.set_chaining(None)
    ctx.submit(0, golden["bases"], golden["off"], opt)
    with pytest.raises(cuda_lib.CompSeedError):
        ctx.wait_chains(0)
    ctx.wait(0)
    ctx.close(); idx.close()


@pytest.mark.parametrize("kind", ["random", "repeat", "contigs"])
def test_chains_against_the_reference(cuda_lib, oracle_lib, kind):
    if not oracle_lib.have_ref():
        pytest.skip("needs oracle/_ref")
    if kind == "random":       # cfg1-like: one chain for most reads
        ref = synth.random_reference(2_000_000, seed=911)
        bases, off, _ = synth.simulate_reads(ref, 60_000, 150, 0.01, seed=912)
        lens, alt = [ref.shape[0]], None
    elif kind == "repeat":     # hundreds of chains per read: B-tree splits, equal keys, equal weights
        ref = synth.repeat_rich_reference(400_000, seed=913, n_segdup=100, segdup_len=2000, n_tandem=200)
        bases, off, _ = synth.simulate_reads(ref, 6000, [100, 150, 250], 0.02, seed=914, n_rate=0.002)
        lens, alt = [ref.shape[0]], None
    else:
        ref = synth.repeat_rich_reference(300_000, seed=915, n_segdup=60, segdup_len=1500, n_tandem=40)
        bases, off, _ = synth.simulate_reads(ref, 8000, [100, 150], 0.01, seed=916)
        lens, alt = [100_000, 50_000, 150_000], [0, 1, 0]
    idx = cuda_lib.FMIndex.build(ref, sa_intv=1)
    n = off.shape[0] - 1
    ctx = cuda_lib.SeedContext(idx, n, int(off[-1]), 256, n * 64, n * 2000, 1)
    for co in (cuda_lib.ChainOpt(), cuda_lib.ChainOpt(w=50, max_chain_gap=300, min_chain_weight=30, max_chain_extend=3, mask_level=0.3, drop_ratio=0.8)):
        ctx.set_chaining(lens, co, alt)
        ctx.submit(0, bases, off, cuda_lib.SeedOpt())
        try:
            got = ctx.wait_chains(0)
        except cuda_lib.CompSeedError as e:
            # repeat-rich reads: more chains than the mems the ctx was sized for (chain records share max_mems).  The library says
            # what the batch needs; a caller re-creates the ctx once with that (what cs_multi_* and the shim do)
            assert e.code == -3 and "chain buffers too small" in str(e), e
            need_m, need_s = ctx.need(0)
            assert need_m > n * 64 or need_s > n * 2000
            ctx.close()
            ctx = cuda_lib.SeedContext(idx, n, int(off[-1]), 256, need_m, need_s, 1)
            ctx.set_chaining(lens, co, alt)
            ctx.submit(0, bases, off, cuda_lib.SeedOpt())
            got = ctx.wait_chains(0)
        seeds = ctx.fetch(0)
        want = oracle_lib.ref_chain(off, seeds, lens, w=co.w, max_chain_gap=co.max_chain_gap, min_chain_weight=co.min_chain_weight,
                                    max_chain_extend=co.max_chain_extend, mask_level=co.mask_level, drop_ratio=co.drop_ratio, is_alt=alt)
        assert_chains_equal(_as_dict(got), want, off)
        assert got.chain_off[-1] > 0
    ctx.close(); idx.close()


def test_chains_through_the_multi_pipeline(cuda_lib, oracle_lib, golden):
    """cs_multi_set_chaining: the block arrays carry chains (DMA, no host copy); same chains as the single-ctx path."""
    if golden["name"] != "repeat30k":
        pytest.skip("one fixture is enough")
    idx = cuda_lib.FMIndex.upload(int(golden["primary"]), golden["L2"], int(golden["seq_len"]), golden["bwt"], golden["sa"], int(golden["sa_intv"]), dense_sa_intv=1)
    want = SimpleNamespace(**{k: golden["chain_" + k + "0"] for k in ("pos", "rid", "w", "kept", "n", "s_rbeg", "s_qbeg", "s_len", "frac_rep")}, chain_off=golden["chain_off0"])
    idxs = [idx] + [cuda_lib.replicate_index(idx, d) for d in range(1, min(2, cuda_lib.device_count()))]
    ms = cuda_lib.MultiSeeder(idxs, batch_reads=64, max_read_len=256, n_slots=3, mems_per_read=1, seeds_per_read=1)   # tiny estimates: everything grows
    ms.set_chaining([golden["ref"].shape[0]])
    o64 = golden["off"].astype(np.uint64)
    for s in (0, 1):
        ms.submit(s, golden["bases"], o64, cuda_lib.SeedOpt())
    for s in (0, 1):
        got = ms.wait(s)
        assert_chains_equal(_as_dict(got), want, golden["off"])
    ms.set_chaining(None)                      # and back to mems + seed positions
    ms.submit(0, golden["bases"], o64, cuda_lib.SeedOpt())
    r = ms.wait(0)
    assert np.array_equal(r.mems, golden["mems0"]) and np.array_equal(r.rbeg, golden["rbeg0"])
    ms.close()
    for i in idxs:
        i.close()

"""The extension kernel through the C-ABI (cs_bsw_*) on the GPU against the golden vectors of the unmodified reference
(BandedPairWiseSW::scalarBandedSWAWrapper == ksw_extend2 per pair; tests/golden/bsw3k.npz) and against the oracle's restatement on
seeded batches: score, qle, tle, gtle, gscore, max_off of every pair, bit for bit."""
import os

import numpy as np
import pytest

from compseed_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def full_pairs(p8):
    p = np.zeros((p8.shape[0], 14), np.int32)
    p[:, :8] = p8
    return p


def test_golden_bsw(cuda_lib):
    z = np.load(os.path.join(ROOT, "tests", "golden", "bsw3k.npz"))
    pairs = full_pairs(z["pairs"])
    ex = cuda_lib.BswExtender(0, 1024, 1 << 16, 1 << 16, 64)      # sized too small on purpose: every buffer grows
    for k, (w, o_del, e_del, o_ins, e_ins, zdrop, eb, a, b) in enumerate(z["opts"].tolist()):
        opt = cuda_lib.BswOpt(o_del, e_del, o_ins, e_ins, zdrop, eb, cuda_lib.bwa_fill_scmat(a, b))
        got = ex.extend(pairs.copy(), z["seq_buf_ref"], z["seq_buf_qer"], w, opt)
        assert np.array_equal(got[:, 8:], z["res%d" % k]), k
        assert np.array_equal(got[:, :8], pairs[:, :8])            # the caller's fields are left alone
    assert ex.launches > 0
    ex.close()


@pytest.mark.parametrize("case", ["default", "long", "clean", "noisy", "tight"])
def test_bsw_equals_oracle(cuda_lib, oracle_lib, case):
    kw = dict(w=100, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5)
    a, b = 1, 4
    if case == "default":
        pairs, ref, qer = synth.extension_pairs(6000, seed=401)
    elif case == "long":
        pairs, ref, qer = synth.extension_pairs(800, seed=402, max_qlen=1500, max_seed_len=400)
    elif case == "clean":
        pairs, ref, qer = synth.extension_pairs(4000, seed=403, sub_rate=0.0, indel_rate=0.0, n_rate=0.0, unrelated_frac=0.0)
    elif case == "noisy":
        pairs, ref, qer = synth.extension_pairs(4000, seed=404, sub_rate=0.08, indel_rate=0.03, n_rate=0.02, unrelated_frac=0.2)
    else:
        pairs, ref, qer = synth.extension_pairs(4000, seed=405)
        kw = dict(w=7, o_del=3, e_del=2, o_ins=5, e_ins=1, zdrop=15, end_bonus=0)
        a, b = 2, 5
    want, cells = oracle_lib.oracle_bsw(pairs, ref, qer, mat=oracle_lib.bsw_mat(a, b), n_threads=8, **kw)
    ex = cuda_lib.BswExtender(0, pairs.shape[0], ref.nbytes, qer.nbytes, 256)
    opt = cuda_lib.BswOpt(kw["o_del"], kw["e_del"], kw["o_ins"], kw["e_ins"], kw["zdrop"], kw["end_bonus"], cuda_lib.bwa_fill_scmat(a, b))
    got = ex.extend(pairs.copy(), ref, qer, kw["w"], opt)
    assert np.array_equal(got, want)
    # the staged path (inputs resident on the device), twice: same answers, same cells as the oracle
    ex.stage(pairs, ref, qer)
    for _ in range(2):
        ms, gcells = ex.run_staged(kw["w"], opt)
        assert gcells == cells and ms > 0
    assert np.array_equal(ex.fetch(pairs.copy()), want)
    # the kernel with its DP rows in the HBM scratch instead of shared memory (what long queries fall back to): same answers
    ex.set_rows_in_smem(False)
    ms, gcells = ex.run_staged(kw["w"], opt)
    assert gcells == cells and np.array_equal(ex.fetch(pairs.copy()), want)
    ex.close()


def test_bsw_large_batch_and_order_independence(cuda_lib, oracle_lib):
    """1 M pairs (the size class of a bench step) against the oracle; and a shuffled copy of the batch gives the same per-pair answers."""
    pairs, ref, qer = synth.extension_pairs_fast(1_000_000, seed=411)
    want, cells = oracle_lib.oracle_bsw(pairs, ref, qer, n_threads=os.cpu_count() or 8)
    ex = cuda_lib.BswExtender(0, pairs.shape[0], ref.nbytes, qer.nbytes, 256)
    got = ex.extend(pairs.copy(), ref, qer)
    assert np.array_equal(got, want)
    perm = np.random.default_rng(3).permutation(pairs.shape[0])
    got2 = ex.extend(np.ascontiguousarray(pairs[perm]), ref, qer)
    assert np.array_equal(got2, want[perm])
    ex.close()


def test_bsw_argument_errors(cuda_lib):
    ex = cuda_lib.BswExtender(0, 16)
    pairs = np.zeros((2, 14), np.int32)
    pairs[:, 3], pairs[:, 4], pairs[:, 5] = 4, 4, 19
    buf = np.zeros(8, np.uint8)
    bad = pairs.copy(); bad[1, 4] = 0                      # an empty query
    with pytest.raises(cuda_lib.CompSeedError):
        ex.extend(bad, buf, buf)
    bad = pairs.copy(); bad[1, 0] = 6                      # target runs past the buffer
    with pytest.raises(cuda_lib.CompSeedError):
        ex.extend(bad, buf, buf)
    with pytest.raises(cuda_lib.CompSeedError):            # nothing staged
        cuda_lib.BswExtender(0, 16).run_staged()
    ok = ex.extend(pairs.copy(), buf, buf)                 # AAAA vs AAAA: 4 matches on top of h0
    assert ok[0, 8] == 23 and ok[0, 11] == 4 and ok[0, 9] == 4 and ok[0, 12] == 23
    ex.close()

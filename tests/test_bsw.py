"""The extension stage on the CPU: (1) the oracle's restatement of ksw_extend2 (oracle/cs_oracle.c: cso_bsw_extend)
against the golden vectors the unmodified reference wrote (tests/golden/bsw3k.npz, make_golden_bsw.py) and, when oracle/_ref is
present, live against BandedPairWiseSW::scalarBandedSWAWrapper / getScores8 / getScores16 and ksw_extend2; (2) the SOURCE of the
CUDA kernel (compseed_b200/csrc/cs_bsw.cuh: bsw_one_pair) compiled as plain C++ and run serially (tests/emul/bsw_emul.cpp)
against the same vectors.  Test infrastructure only -- the shipped library has no CPU path; tests/test_gpu_bsw.py runs the
real kernel through the C-ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from compseed_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcsref.so"))


@pytest.fixture(scope="module")
def bsw_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "bsw3k.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libbsw_emul.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-w", "-o", so, os.path.join(ROOT, "tests", "emul", "bsw_emul.cpp")])
    L = C.CDLL(so)
    L.bsw_emul.restype = C.c_uint64
    L.bsw_emul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_int] * 7 + [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
    return L


def full_pairs(p8):
    p = np.zeros((p8.shape[0], 14), np.int32)
    p[:, :8] = p8
    return p


def run_emul(L, pairs, ref, qer, w, o_del, e_del, o_ins, e_ins, zdrop, eb, mat, stride=1, wide=0, qpack=0):
    out = np.ascontiguousarray(pairs, np.int32).copy()
    m = np.ascontiguousarray(mat, np.int8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cells = L.bsw_emul(p(out), p(ref), p(qer), out.shape[0], w, o_del, e_del, o_ins, e_ins, zdrop, eb, p(m), stride, wide, qpack)
    return out, int(cells)


def test_oracle_bsw_equals_golden(oracle_lib, bsw_golden):
    g = bsw_golden
    pairs = full_pairs(g["pairs"])
    for k, (w, o_del, e_del, o_ins, e_ins, zdrop, eb, a, b) in enumerate(g["opts"].tolist()):
        got, cells = oracle_lib.oracle_bsw(pairs, g["seq_buf_ref"], g["seq_buf_qer"], w, o_del, e_del, o_ins, e_ins, zdrop, eb, oracle_lib.bsw_mat(a, b), n_threads=4)
        assert np.array_equal(got[:, 8:], g["res%d" % k]), k
        assert np.array_equal(got[:, :8], pairs[:, :8]) and cells > 0


@pytest.mark.parametrize("stride,wide,qpack", [(1, 0, 0), (5, 1, 0), (3, 0, 0), (4, 0, 1), (2, 1, 1)])
def test_kernel_source_equals_golden(emul, oracle_lib, bsw_golden, stride, wide, qpack):
    g = bsw_golden
    pairs = full_pairs(g["pairs"])
    for k, (w, o_del, e_del, o_ins, e_ins, zdrop, eb, a, b) in enumerate(g["opts"].tolist()):
        got, cells = run_emul(emul, pairs, g["seq_buf_ref"], g["seq_buf_qer"], w, o_del, e_del, o_ins, e_ins, zdrop, eb, oracle_lib.bsw_mat(a, b), stride, wide, qpack)
        assert np.array_equal(got[:, 8:], g["res%d" % k]), k
        _, ocells = oracle_lib.oracle_bsw(pairs, g["seq_buf_ref"], g["seq_buf_qer"], w, o_del, e_del, o_ins, e_ins, zdrop, eb, oracle_lib.bsw_mat(a, b))
        assert cells == ocells   # the same cells, not just the same answers


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", ["default", "long", "clean", "noisy", "tight"])
def test_oracle_and_kernel_source_equal_reference_bsw(emul, oracle_lib, case):
    kw = dict(w=100, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5)
    a, b = 1, 4
    if case == "default":
        pairs, ref, qer = synth.extension_pairs(4000, seed=301)
    elif case == "long":       # 250-bp reads and beyond: the 16-bit class of the reference
        pairs, ref, qer = synth.extension_pairs(600, seed=302, max_qlen=600, max_seed_len=300)
    elif case == "clean":      # error-free: every extension runs to the end of the query (gscore / gtle paths)
        pairs, ref, qer = synth.extension_pairs(3000, seed=303, sub_rate=0.0, indel_rate=0.0, n_rate=0.0, unrelated_frac=0.0)
    elif case == "noisy":      # many indels and Ns: band growth, z-drop
        pairs, ref, qer = synth.extension_pairs(3000, seed=304, sub_rate=0.08, indel_rate=0.03, n_rate=0.02, unrelated_frac=0.2)
    else:                      # small band, small z-drop, asymmetric penalties, other match / mismatch scores
        pairs, ref, qer = synth.extension_pairs(3000, seed=305)
        kw = dict(w=7, o_del=3, e_del=2, o_ins=5, e_ins=1, zdrop=15, end_bonus=0)
        a, b = 2, 5
    want, _ = oracle_lib.ref_bsw(pairs, ref, qer, mode=0, a=a, b=b, **kw)
    got, cells = oracle_lib.oracle_bsw(pairs, ref, qer, mat=oracle_lib.bsw_mat(a, b), n_threads=4, **kw)
    assert np.array_equal(got, want)
    got2, cells2 = run_emul(emul, pairs, ref, qer, kw["w"], kw["o_del"], kw["e_del"], kw["o_ins"], kw["e_ins"], kw["zdrop"], kw["end_bonus"], oracle_lib.bsw_mat(a, b), 3, 0)
    assert np.array_equal(got2, want) and cells2 == cells
    got3, cells3 = run_emul(emul, pairs, ref, qer, kw["w"], kw["o_del"], kw["e_del"], kw["o_ins"], kw["e_ins"], kw["zdrop"], kw["end_bonus"], oracle_lib.bsw_mat(a, b), 2, 1)
    assert np.array_equal(got3, want) and cells3 == cells
    got4, cells4 = run_emul(emul, pairs, ref, qer, kw["w"], kw["o_del"], kw["e_del"], kw["o_ins"], kw["e_ins"], kw["zdrop"], kw["end_bonus"], oracle_lib.bsw_mat(a, b), 4, 0, 1)
    assert np.array_equal(got4, want) and cells4 == cells   # query read from its packed copy (the shared-memory kernel)
    if case in ("default", "clean"):   # what CompSeed really calls: the SIMD twins, each on its class of pairs (comp_seed.cpp:1556-1564)
        m8 = (pairs[:, 3] < 128) & (pairs[:, 4] < 128) & (pairs[:, 5] + np.minimum(pairs[:, 3], pairs[:, 4]) < 128)
        v8, _ = oracle_lib.ref_bsw(np.ascontiguousarray(pairs[m8]), ref, qer, mode=1, **kw)
        v16, _ = oracle_lib.ref_bsw(np.ascontiguousarray(pairs[~m8]), ref, qer, mode=2, **kw)
        assert np.array_equal(v8[:, 8:], got[m8][:, 8:]) and np.array_equal(v16[:, 8:], got[~m8][:, 8:])
    # and ksw_extend2 itself, the call of bwamem's mem_chain2aln (bwamem.c:720,746)
    for i in range(0, pairs.shape[0], 211):
        p = pairs[i]
        sc = oracle_lib.ref_ksw_extend2(qer[p[1]:p[1] + p[4]], ref[p[0]:p[0] + p[3]], int(p[5]), a=a, b=b, **kw)
        assert sc == (want[i, 8], want[i, 11], want[i, 9], want[i, 10], want[i, 12], want[i, 13])


def test_bsw_abi_rejects_bad_batches_without_a_device():
    """Argument checks of cs_bsw_* need no device; and without one, creation fails loudly (no CPU path)."""
    import compseed_b200 as cs
    if cs.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(cs.CompSeedError):
        cs.BswExtender(0, 1024)

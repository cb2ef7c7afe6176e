"""The chaining stage (compseed_b200/csrc/cs_chain.cu: mem_chain + mem_chain_flt on the device) checked on the CPU: the
same source compiled as plain C++ with the CUDA qualifiers defined away (tests/emul/chain_emul.cpp) and run serially,
against the UNMODIFIED reference's mem_chain / mem_chain_flt (oracle/_ref) and against the committed golden chains.
Test infrastructure only -- the shipped library has no CPU path; the GPU tests run the real kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from compseed_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcsref.so"))


class _SeedOpt(C.Structure):
    _fields_ = [("min_seed_len", C.c_int32), ("split_len", C.c_int32), ("split_width", C.c_int32), ("max_mem_intv", C.c_int32), ("max_occ", C.c_int32)]


class _ChainOpt(C.Structure):
    _fields_ = [("w", C.c_int32), ("max_chain_gap", C.c_int32), ("min_chain_weight", C.c_int32), ("max_chain_extend", C.c_int32),
                ("mask_level", C.c_float), ("drop_ratio", C.c_float)]


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libchain_emul.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-w", "-o", so, os.path.join(ROOT, "tests", "emul", "chain_emul.cpp")])
    return C.CDLL(so)


def run_emul(L, read_off, res, contig_lens, so=(19, 28, 10, 20, 500), co=(100, 10000, 0, 1 << 30, 0.5, 0.5), is_alt=None):
    n = read_off.shape[0] - 1
    ns = int(res.seed_off[-1])
    lens = np.asarray(contig_lens, dtype=np.int64)
    offs = np.ascontiguousarray(np.concatenate([[0], np.cumsum(lens)[:-1]]), dtype=np.int64)
    alt = np.ascontiguousarray(is_alt, dtype=np.uint8) if is_alt is not None else None
    chain_off = np.zeros(n + 1, np.uint32); cseed_off = np.zeros(n + 1, np.uint32)
    chains = np.zeros((ns + 1, 4), np.uint32); lo = np.zeros(ns + 1, np.uint32); hi = np.zeros(ns + 1, np.uint8)
    qb = np.zeros(ns + 1, np.uint16); ln = np.zeros(ns + 1, np.uint16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    arrs = [np.ascontiguousarray(read_off, np.uint32), np.ascontiguousarray(res.mem_off, np.uint32), np.ascontiguousarray(res.mems, np.uint64),
            np.ascontiguousarray(res.seed_off, np.uint32), np.ascontiguousarray(res.rbeg, np.int64)]
    o1, o2 = _SeedOpt(*so), _ChainOpt(*co)
    L.chain_emul.argtypes = [C.c_uint32] + [C.c_void_p] * 5 + [C.POINTER(_SeedOpt), C.POINTER(_ChainOpt), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p] + [C.c_void_p] * 7
    rc = L.chain_emul(n, *[p(a) for a in arrs], C.byref(o1), C.byref(o2), int(lens.sum()), lens.shape[0], p(offs), p(alt) if alt is not None else None,
                      p(chain_off), p(cseed_off), p(chains), p(lo), p(hi), p(qb), p(ln))
    assert rc == 0
    nc, nk = int(chain_off[-1]), int(cseed_off[-1])
    wk = chains[:nc, 1]
    rb = ((lo[:nk].astype(np.int64) | (hi[:nk].astype(np.int64) << 32)) << 24) >> 24
    return dict(chain_off=chain_off, cseed_off=cseed_off, rid=chains[:nc, 0].astype(np.int32), w=(wk & 0x1fffffff).astype(np.int32),
                kept=(((wk >> 29) & 3) | ((wk >> 31) << 8)).astype(np.int32), n=chains[:nc, 2].astype(np.int32), l_rep=chains[:nc, 3].copy(),
                s_rbeg=rb, s_qbeg=qb[:nk].astype(np.int32), s_len=ln[:nk].astype(np.int32))


def assert_chains_equal(got, want, read_off):
    """got: dict from run_emul / SeedContext.wait_chains fields; want: oracle_py.ChainResult (the reference's output)."""
    assert np.array_equal(got["chain_off"], want.chain_off)
    assert np.array_equal(got["rid"], want.rid) and np.array_equal(got["w"], want.w) and np.array_equal(got["kept"], want.kept) and np.array_equal(got["n"], want.n)
    assert np.array_equal(got["s_rbeg"], want.s_rbeg) and np.array_equal(got["s_qbeg"], want.s_qbeg) and np.array_equal(got["s_len"], want.s_len)
    # pos of a chain == rbeg of its first seed; frac_rep == (float) l_rep / l_seq for every chain of the read
    lens = np.diff(read_off.astype(np.int64))
    per_chain_read = np.repeat(np.arange(lens.shape[0]), np.diff(want.chain_off.astype(np.int64)))
    first_seed = np.concatenate([[0], np.cumsum(want.n)[:-1]]).astype(np.int64) if want.n.shape[0] else np.empty(0, np.int64)
    assert np.array_equal(want.pos, want.s_rbeg[first_seed])
    frac = got["l_rep"].astype(np.float32) / lens[per_chain_read].astype(np.float32)
    assert np.array_equal(frac, want.frac_rep[per_chain_read])


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind", ["random", "repeat", "tandem", "contigs"])
def test_chain_kernels_equal_reference_chaining(emul, oracle_lib, kind):
    if kind == "random":
        ref = synth.random_reference(300_000, seed=901)
        bases, off, _ = synth.simulate_reads(ref, 5000, [100, 150, 250], 0.01, seed=902, n_rate=0.001)
        lens = [ref.shape[0]]
    elif kind == "repeat":     # hundreds of chains per read: the B-tree splits, equal keys, long sorts
        ref = synth.repeat_rich_reference(300_000, seed=903, n_segdup=100, segdup_len=2000, n_tandem=60)
        bases, off, _ = synth.simulate_reads(ref, 3000, [100, 150, 250], 0.02, seed=904, n_rate=0.003)
        lens = [ref.shape[0]]
    elif kind == "tandem":     # tandem repeats inside the reads: several chains at one reference position, equal weights
        ref = synth.repeat_rich_reference(120_000, seed=905, n_segdup=10, segdup_len=1000, n_tandem=300)
        bases, off, _ = synth.simulate_reads(ref, 2500, [150, 250], 0.01, seed=906)
        lens = [ref.shape[0]]
    else:                      # several reference sequences, one of them ALT: seeds bridging two of them are dropped, rid separates chains
        ref = synth.repeat_rich_reference(200_000, seed=907, n_segdup=50, segdup_len=1500, n_tandem=30)
        bases, off, _ = synth.simulate_reads(ref, 3000, [100, 150], 0.01, seed=908)
        lens = [50_000, 70_000, 30_000, 50_000]
    oi = oracle_lib.OracleIndex.build(ref)
    res = oi.seed(bases, off, n_threads=8)
    alt = [0, 0, 1, 0] if kind == "contigs" else None
    for co in ((100, 10000, 0, 1 << 30, 0.5, 0.5), (50, 300, 30, 3, 0.3, 0.8)):
        want = oracle_lib.ref_chain(off, res, lens, w=co[0], max_chain_gap=co[1], min_chain_weight=co[2], max_chain_extend=co[3],
                                    mask_level=co[4], drop_ratio=co[5], is_alt=alt)
        got = run_emul(emul, off, res, lens, co=co, is_alt=alt)
        assert_chains_equal(got, want, off)
    assert want.chain_off[-1] > 0


def test_chain_kernels_equal_golden_chains(emul, golden):
    """Chains of the committed fixtures (written by the reference through tests/golden/make_golden.py)."""
    if "chain_off0" not in golden:
        pytest.skip("fixture without chains")
    from types import SimpleNamespace
    res = SimpleNamespace(mem_off=golden["mem_off0"], mems=golden["mems0"], seed_off=golden["seed_off0"], rbeg=golden["rbeg0"])
    want = SimpleNamespace(**{k: golden["chain_" + k + "0"] for k in ("off", "pos", "rid", "w", "kept", "n", "s_rbeg", "s_qbeg", "s_len", "frac_rep")})
    want.chain_off = want.off
    o = golden["opt0"]
    got = run_emul(emul, golden["off"], res, [golden["ref"].shape[0]], so=tuple(int(x) for x in o))
    assert_chains_equal(got, want, golden["off"])

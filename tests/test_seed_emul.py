"""The seeding and index-construction kernels of compseed_b200/csrc/cs_kernels.cu checked on the CPU: the same source compiled as
plain C++ (tests/emul/seed_emul.cpp) and run (a) with one-lane warps, serially -- k_seed_fast, k_seed_walk, k_seed_r3_fast, the calls
handed on to the literal kernel resolved with the oracle's bwt_smem1a --, and (b) with REAL warps, one host thread per lane and every
warp intrinsic a rendezvous of the 32 -- all seeding kernels, the literal k_seed (call mode and read mode) and the general third pass
included, nothing resolved by the oracle.  The result must equal the oracle's mem_collect_intv, with and without the repeat-length
array (DevIndex::rep).  Test infrastructure only -- the shipped library has no CPU path; the GPU tests run the real kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from compseed_b200 import synth
from oracle import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _SeedOpt(C.Structure):
    _fields_ = [("min_seed_len", C.c_int32), ("split_len", C.c_int32), ("split_width", C.c_int32), ("max_mem_intv", C.c_int32), ("max_occ", C.c_int32)]


def _build(tmp_path_factory, name, defs):
    d = tmp_path_factory.mktemp(name)
    obj, so = str(d / "oracle.o"), str(d / "libseed_emul.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-w", "-c", os.path.join(ROOT, "oracle", "cs_oracle.c"), "-o", obj])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-DCS_STATS", *defs, "-o", so, os.path.join(ROOT, "tests", "emul", "seed_emul.cpp"), obj, "-lpthread"])
    L = C.CDLL(so)
    L.seed_emul_index.restype = C.c_void_p
    L.seed_emul_index.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64] + [C.c_int] * 5
    L.seed_emul_run.restype = C.c_int64
    L.seed_emul_run.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(_SeedOpt), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.seed_emul_free.argtypes = [C.c_void_p]
    L.seed_emul_rep.restype = C.c_void_p
    L.seed_emul_rep.argtypes = [C.c_void_p]
    L.seed_emul_run32.restype = C.c_int64
    L.seed_emul_run32.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(_SeedOpt), C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    return L


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    return _build(tmp_path_factory, "seed_emul", [])


@pytest.fixture(scope="module")
def emul_spec(tmp_path_factory):
    """the build with the diagonal speculation of k_seed_fast switched on (CS_SPEC_DIAG, off in the product: measured slower)"""
    return _build(tmp_path_factory, "seed_emul_spec", ["-DCS_SPEC_DIAG=1"])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_k_depth(seq_len):
    """K and table depth as cs_api.cu picks them for a text of seq_len rows."""
    K = 2
    while K < 19 and (1 << (2 * (K - 2))) < seq_len:
        K += 1
    K = max(K, 8)
    depth = 0
    while depth < 13 and (seq_len >> (2 * (depth + 1))) >= 64:
        depth += 1
    return K, depth


def run_emul(L, oi, bases, off, so, use_rep, isa_intv=2):
    K, depth = default_k_depth(oi.seq_len)
    K = min(K, so[0])
    L2 = np.ascontiguousarray(oi.L2, np.uint64)
    E = L.seed_emul_index(oi.primary, _p(L2), oi.seq_len, _p(oi.bwt), oi.bwt_size, _p(oi.sa), oi.n_sa, oi.sa_intv, K, depth, isa_intv, use_rep)
    try:
        n = off.shape[0] - 1
        cap = n * 64
        mems = np.zeros((cap, 4), np.uint64)
        mem_off = np.zeros(n + 1, np.uint32)
        stats = np.zeros(64, np.uint64)
        rc = L.seed_emul_run(E, n, _p(bases), _p(off), C.byref(_SeedOpt(*so)), _p(mems), cap, _p(mem_off), _p(stats))
        assert rc >= 0, rc
        rep = None
        if use_rep:
            rep = np.ctypeslib.as_array(C.cast(L.seed_emul_rep(E), C.POINTER(C.c_uint8)), shape=(oi.seq_len,)).copy()
        return mem_off, mems[:rc].copy(), stats, rep
    finally:
        L.seed_emul_free(E)


CASES = [
    ("random", lambda: synth.random_reference(150_000, seed=3), dict(n=1500, lens=[100, 150, 250], err=0.01, n_rate=0.002)),
    ("random_3pct", lambda: synth.random_reference(150_000, seed=4), dict(n=1000, lens=[150], err=0.03, n_rate=0.0)),
    ("repeat_rich", lambda: synth.repeat_rich_reference(150_000, seed=41, n_segdup=40, segdup_len=2500, n_tandem=20), dict(n=1500, lens=[100, 150, 250], err=0.01, n_rate=0.001)),
]


@pytest.mark.parametrize("name,mk,rd", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("so", [(19, 28, 10, 0, 500), (19, 19, 10, 0, 500), (25, 30, 3, 0, 50)], ids=["default", "r1.0", "k25s3"])
def test_fast_and_walk_kernels_equal_the_oracle_with_and_without_repeat_lengths(emul, name, mk, rd, so):
    ref = mk()
    bases, off, _ = synth.simulate_reads(ref, rd["n"], rd["lens"], rd["err"], seed=11, n_rate=rd["n_rate"])
    oi = O.OracleIndex.build(ref)
    want = oi.seed(bases, off, min_seed_len=so[0], split_len=so[1], split_width=so[2], max_mem_intv=0, max_occ=so[4])
    n = off.shape[0] - 1
    req = {}
    for use_rep in (0, 1):
        mem_off, mems, stats, _ = run_emul(emul, oi, bases, off, so, use_rep)
        assert np.array_equal(mem_off, want.mem_off) and np.array_equal(mems, want.mems), (name, so, use_rep)
        req[use_rep] = int(stats[4])
        if use_rep:   # the second-pass calls answered where their SMEM was found: some, and none of them changed a result
            answered = int(stats[16 + 12] + stats[16 + 13] + stats[16 + 14])
            assert answered > 0
    assert req[1] < req[0]   # fewer executed gathers with the repeat lengths
    if name.startswith("random"):
        assert req[1] < 0.85 * req[0]


@pytest.mark.parametrize("name,mk,rd", CASES, ids=[c[0] for c in CASES])
def test_diagonal_speculation_build_equals_the_oracle(emul_spec, name, mk, rd):
    so = (19, 28, 10, 0, 500)
    ref = mk()
    bases, off, _ = synth.simulate_reads(ref, rd["n"], rd["lens"], rd["err"], seed=12, n_rate=rd["n_rate"])
    oi = O.OracleIndex.build(ref)
    want = oi.seed(bases, off, min_seed_len=so[0], split_len=so[1], split_width=so[2], max_mem_intv=0, max_occ=so[4])
    mem_off, mems, stats, _ = run_emul(emul_spec, oi, bases, off, so, 1)
    assert np.array_equal(mem_off, want.mem_off) and np.array_equal(mems, want.mems)
    assert int(stats[16 + 7]) > 0    # speculation taken


@pytest.mark.parametrize("name,mk,rd", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("so", [(19, 28, 10, 20, 500), (19, 28, 10, 40, 50)], ids=["y20", "y40c50"])
def test_third_pass_kernel_equals_the_oracle(emul, name, mk, rd, so):
    """all three passes: k_seed_fast + k_seed_walk + k_seed_r3_fast (the text-assisted third pass) against the oracle's mem_collect_intv"""
    ref = mk()
    bases, off, _ = synth.simulate_reads(ref, rd["n"], rd["lens"], rd["err"], seed=13, n_rate=rd["n_rate"])
    oi = O.OracleIndex.build(ref)
    want = oi.seed(bases, off, min_seed_len=so[0], split_len=so[1], split_width=so[2], max_mem_intv=so[3], max_occ=so[4])
    req = {}
    for use_rep in (0, 1):
        mem_off, mems, stats, _ = run_emul(emul, oi, bases, off, so, use_rep)
        assert np.array_equal(mem_off, want.mem_off) and np.array_equal(mems, want.mems), (name, so, use_rep)
        req[use_rep] = int(stats[12])
    assert req[0] > 0 and req[1] > 0


@pytest.mark.parametrize("name,mk,rd", CASES, ids=[c[0] for c in CASES])
def test_real_warps_every_seeding_kernel_equals_the_oracle(emul, name, mk, rd):
    """32 host threads per warp: k_seed_fast -> k_seed_walk -> k_seed (call mode) -> k_seed_r3_fast, and k_seed alone (read mode) ->
    k_seed_r3; vote-controlled loops, warp-aggregated atomics and the literal kernel's cooperative filter with their real semantics."""
    so = (19, 28, 10, 20, 500)
    ref = mk()
    n_reads = 500
    bases, off, _ = synth.simulate_reads(ref, n_reads, rd["lens"], rd["err"], seed=14, n_rate=rd["n_rate"])
    oi = O.OracleIndex.build(ref)
    want = oi.seed(bases, off, min_seed_len=so[0], split_len=so[1], split_width=so[2], max_mem_intv=so[3], max_occ=so[4])
    K, depth = default_k_depth(oi.seq_len)
    L2 = np.ascontiguousarray(oi.L2, np.uint64)
    E = emul.seed_emul_index(oi.primary, _p(L2), oi.seq_len, _p(oi.bwt), oi.bwt_size, _p(oi.sa), oi.n_sa, oi.sa_intv, min(K, so[0]), depth, 2, 1)
    try:
        for mode in (0, 1):
            cap = n_reads * 64
            mems = np.zeros((cap, 4), np.uint64)
            mem_off = np.zeros(n_reads + 1, np.uint32)
            stats = np.zeros(64, np.uint64)
            rc = emul.seed_emul_run32(E, n_reads, _p(bases), _p(off), C.byref(_SeedOpt(*so)), mode, _p(mems), cap, _p(mem_off), _p(stats))
            assert rc == want.mem_off[-1], (name, mode, rc)
            assert np.array_equal(mem_off, want.mem_off) and np.array_equal(mems[:rc], want.mems), (name, mode)
            if mode == 0:
                assert int(stats[6]) > 0          # calls were handed on
    finally:
        emul.seed_emul_free(E)


@pytest.mark.parametrize("so", [(19, 28, 10, 20, 500), (19, 19, 10, 40, 50)], ids=["default", "r1.0y40c50"])
def test_real_warps_boundary_reads(emul, so):
    """synth.boundary_reads (text ends of both strands, strand-bridging matches, substitutions and Ns at 32-base word edges, lengths
    around multiples of 32 and 255/256) through every seeding kernel with real warps."""
    ref = synth.random_reference(60_000, seed=51)
    bases, off = synth.boundary_reads(ref)
    n_reads = off.shape[0] - 1
    oi = O.OracleIndex.build(ref)
    want = oi.seed(bases, off, min_seed_len=so[0], split_len=so[1], split_width=so[2], max_mem_intv=so[3], max_occ=so[4])
    K, depth = default_k_depth(oi.seq_len)
    L2 = np.ascontiguousarray(oi.L2, np.uint64)
    E = emul.seed_emul_index(oi.primary, _p(L2), oi.seq_len, _p(oi.bwt), oi.bwt_size, _p(oi.sa), oi.n_sa, oi.sa_intv, min(K, so[0]), depth, 2, 1)
    try:
        for mode in (0, 1):
            cap = n_reads * 64
            mems = np.zeros((cap, 4), np.uint64)
            mem_off = np.zeros(n_reads + 1, np.uint32)
            rc = emul.seed_emul_run32(E, n_reads, _p(bases), _p(off), C.byref(_SeedOpt(*so)), mode, _p(mems), cap, _p(mem_off), None)
            assert rc == want.mem_off[-1], (mode, rc)
            assert np.array_equal(mem_off, want.mem_off) and np.array_equal(mems[:rc], want.mems), mode
    finally:
        emul.seed_emul_free(E)


def test_repeat_lengths_match_their_definition(emul):
    """rep[p] against a brute-force count of occurrences on a small repeat-rich text."""
    ref = synth.repeat_rich_reference(6_000, seed=5, n_segdup=6, segdup_len=300, n_tandem=4)
    oi = O.OracleIndex.build(ref)
    bases, off, _ = synth.simulate_reads(ref, 4, [100], 0.0, seed=1)
    _, _, _, rep = run_emul(emul, oi, bases, off, (19, 28, 10, 0, 500), 1)
    text = np.concatenate([ref, (3 - ref)[::-1]]).astype(np.uint8)   # T = fwd + revcomp(fwd)
    n = text.shape[0]
    assert rep.shape[0] == n
    s = bytes(text)
    rng = np.random.default_rng(0)
    for p in rng.integers(0, n, 400).tolist() + [0, n - 1, n // 2 - 1, n // 2]:
        R = int(rep[p])
        if R:
            assert p + R <= n and s.find(s[p:p + R]) != s.rfind(s[p:p + R]), p          # T[p, p+R) occurs at least twice
        if R < 255 and p + R < n:
            sub = s[p:p + R + 1]
            assert s.find(sub) == s.rfind(sub) == p, p                                    # one base more: only here

"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/liboracle.so (our CPU restatement) and
oracle/_ref/libcsref.so (the unmodified reference + harness).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcsref.so")
REF_BIN = os.path.join(HERE, "_ref")

CNT_NAMES = ["ext", "ext2", "lf", "sa", "mem", "ext_r1", "ext_r2", "ext_r3"]


def build(ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "-j8"])


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class _Index(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
                ("bwt", C.POINTER(C.c_uint32)), ("sa_intv", C.c_int32), ("n_sa", C.c_uint64), ("sa", C.POINTER(C.c_uint64))]


class _Opt(C.Structure):
    _fields_ = [("min_seed_len", C.c_int32), ("split_len", C.c_int32), ("split_width", C.c_int32),
                ("max_mem_intv", C.c_int32), ("max_occ", C.c_int32)]


class _RefOpt(C.Structure):
    _fields_ = [("min_seed_len", C.c_int32), ("split_factor", C.c_float), ("split_width", C.c_int32),
                ("max_mem_intv", C.c_int32), ("max_occ", C.c_int32)]


@dataclass
class SeedResult:
    mem_off: np.ndarray   # u32 [n+1]
    mems: np.ndarray      # u64 [n_mems, 4]  (x0, x1, x2, info)
    seed_off: np.ndarray  # u32 [n+1]
    rbeg: np.ndarray      # i64 [n_seeds]
    seconds: float = 0.0
    counters: dict | None = None

    def same_as(self, o: "SeedResult") -> bool:
        return (np.array_equal(self.mem_off, o.mem_off) and np.array_equal(self.mems, o.mems)
                and np.array_equal(self.seed_off, o.seed_off) and np.array_equal(self.rbeg, o.rbeg))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.cso_index_build.restype = C.POINTER(_Index)
        L.cso_index_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.cso_index_load.restype = C.POINTER(_Index)
        L.cso_index_load.argtypes = [C.c_char_p]
        L.cso_index_dump.argtypes = [C.POINTER(_Index), C.c_char_p]
        L.cso_index_free.argtypes = [C.POINTER(_Index)]
        L.cso_occ4_many.argtypes = [C.POINTER(_Index), C.c_int, C.c_void_p, C.c_void_p]
        L.cso_extend_many.argtypes = [C.POINTER(_Index), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cso_sa_many.argtypes = [C.POINTER(_Index), C.c_int, C.c_void_p, C.c_void_p]
        L.cso_seed.restype = C.c_void_p
        L.cso_seed.argtypes = [C.POINTER(_Index), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(_Opt)]
        L.cso_result_n_mems.restype = C.c_uint64
        L.cso_result_n_mems.argtypes = [C.c_void_p]
        L.cso_result_n_seeds.restype = C.c_uint64
        L.cso_result_n_seeds.argtypes = [C.c_void_p]
        L.cso_result_seconds.restype = C.c_double
        L.cso_result_seconds.argtypes = [C.c_void_p]
        L.cso_result_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.cso_result_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.cso_result_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleIndex:
    """FM-index in the reference's in-memory layout, owned by liboracle.so."""

    def __init__(self, handle, keep=None):
        self.h = handle
        self._keep = keep   # arrays we only borrow (from_arrays): never freed by the C side
        i = handle.contents
        self.primary = int(i.primary)
        self.L2 = np.array(list(i.L2), dtype=np.uint64)
        self.seq_len = int(i.seq_len)
        self.bwt_size = int(i.bwt_size)
        self.sa_intv = int(i.sa_intv)
        self.n_sa = int(i.n_sa)
        # zero-copy views (valid while this object lives)
        self.bwt = np.ctypeslib.as_array(i.bwt, shape=(self.bwt_size,))
        self.sa = np.ctypeslib.as_array(i.sa, shape=(self.n_sa,))

    @classmethod
    def build(cls, fwd: np.ndarray, sa_intv: int = 32) -> "OracleIndex":
        fwd = np.ascontiguousarray(fwd, dtype=np.uint8)
        h = lib().cso_index_build(_ptr(fwd), fwd.shape[0], sa_intv)
        if not h:
            raise RuntimeError("cso_index_build failed")
        return cls(h)

    @classmethod
    def load(cls, prefix: str) -> "OracleIndex":
        h = lib().cso_index_load(prefix.encode())
        if not h:
            raise RuntimeError(f"cannot load index {prefix}")
        return cls(h)

    @classmethod
    def from_arrays(cls, primary, L2, seq_len, bwt, sa, sa_intv) -> "OracleIndex":
        """Wrap arrays in the reference layout (e.g. an index downloaded from the GPU builder)."""
        lib()
        bwt = np.ascontiguousarray(bwt, dtype=np.uint32)
        sa = np.ascontiguousarray(sa, dtype=np.uint64)
        st = _Index()
        st.primary, st.seq_len, st.bwt_size = int(primary), int(seq_len), int(bwt.shape[0])
        for i in range(5):
            st.L2[i] = int(L2[i])
        st.bwt = bwt.ctypes.data_as(C.POINTER(C.c_uint32))
        st.sa_intv, st.n_sa = int(sa_intv), int(sa.shape[0])
        st.sa = sa.ctypes.data_as(C.POINTER(C.c_uint64))
        return cls(C.pointer(st), keep=(st, bwt, sa))

    def dump(self, prefix: str) -> None:
        if lib().cso_index_dump(self.h, prefix.encode()) != 0:
            raise RuntimeError("cso_index_dump failed")

    def __del__(self):
        try:
            if self._keep is None:
                lib().cso_index_free(self.h)
        except Exception:
            pass

    def occ4(self, k: np.ndarray) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty((k.shape[0], 4), dtype=np.uint64)
        lib().cso_occ4_many(self.h, k.shape[0], _ptr(k), _ptr(out))
        return out

    def extend(self, ik: np.ndarray, is_back: np.ndarray) -> np.ndarray:
        ik = np.ascontiguousarray(ik, dtype=np.uint64)
        is_back = np.ascontiguousarray(is_back, dtype=np.int32)
        out = np.empty((ik.shape[0], 4, 3), dtype=np.uint64)
        lib().cso_extend_many(self.h, ik.shape[0], _ptr(ik), _ptr(is_back), _ptr(out))
        return out

    def sa_lookup(self, k: np.ndarray) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty(k.shape[0], dtype=np.uint64)
        lib().cso_sa_many(self.h, k.shape[0], _ptr(k), _ptr(out))
        return out

    def seed(self, bases: np.ndarray, off: np.ndarray, min_seed_len=19, split_len=28, split_width=10,
             max_mem_intv=20, max_occ=500, n_threads: int = 1) -> SeedResult:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n = off.shape[0] - 1
        opt = _Opt(min_seed_len, split_len, split_width, max_mem_intv, max_occ)
        L = lib()
        r = L.cso_seed(self.h, n_threads, n, _ptr(bases), _ptr(off), C.byref(opt))
        try:
            res = _copy_result(L, "cso", r, n)
            cnt = np.zeros(len(CNT_NAMES), dtype=np.int64)
            L.cso_result_counters(r, _ptr(cnt))
            res.counters = dict(zip(CNT_NAMES, cnt.tolist()))
        finally:
            L.cso_result_free(r)
        return res


def _copy_result(L, prefix: str, r, n: int) -> SeedResult:
    nm = getattr(L, prefix + "_result_n_mems")(r)
    ns = getattr(L, prefix + "_result_n_seeds")(r)
    mem_off = np.empty(n + 1, dtype=np.uint32)
    seed_off = np.empty(n + 1, dtype=np.uint32)
    mems = np.empty((nm, 4), dtype=np.uint64)
    rbeg = np.empty(ns, dtype=np.int64)
    getattr(L, prefix + "_result_copy")(r, _ptr(mem_off), _ptr(mems), _ptr(seed_off), _ptr(rbeg))
    return SeedResult(mem_off, mems, seed_off, rbeg, getattr(L, prefix + "_result_seconds")(r))


# ---------------------------------------------------------------------------------------------
# The unmodified reference (oracle/_ref/libcsref.so)
# ---------------------------------------------------------------------------------------------
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libcsref.so not built (needs /root/reference; see oracle/Makefile)")
        L = C.CDLL(REF_SO)
        L.csref_index_load.restype = C.c_void_p
        L.csref_index_load.argtypes = [C.c_char_p]
        L.csref_index_from_arrays.restype = C.c_void_p
        L.csref_index_from_arrays.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                              C.c_void_p, C.c_uint64, C.c_int]
        L.csref_index_info.argtypes = [C.c_void_p, C.c_void_p]
        L.csref_index_bwt.restype = C.POINTER(C.c_uint32)
        L.csref_index_bwt.argtypes = [C.c_void_p]
        L.csref_index_sa.restype = C.POINTER(C.c_uint64)
        L.csref_index_sa.argtypes = [C.c_void_p]
        L.csref_index_free.argtypes = [C.c_void_p]
        L.csref_occ4.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.csref_extend.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.csref_sa.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.csref_seed.restype = C.c_void_p
        L.csref_seed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(_RefOpt)]
        L.csref_result_n_mems.restype = C.c_uint64
        L.csref_result_n_mems.argtypes = [C.c_void_p]
        L.csref_result_n_seeds.restype = C.c_uint64
        L.csref_result_n_seeds.argtypes = [C.c_void_p]
        L.csref_result_seconds.restype = C.c_double
        L.csref_result_seconds.argtypes = [C.c_void_p]
        L.csref_result_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.csref_result_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.csref_result_free.argtypes = [C.c_void_p]
        _ref = L
    return _ref


class RefIndex:
    """bwt_t of the unmodified reference, either loaded by its own bwt_restore_bwt/sa or wrapped
    around arrays we own (kept alive here)."""

    def __init__(self, handle, keep=None):
        self.h = handle
        self._keep = keep
        info = np.zeros(10, dtype=np.uint64)
        ref_lib().csref_index_info(self.h, _ptr(info))
        self.primary = int(info[0])
        self.L2 = info[1:6].copy()
        self.seq_len, self.bwt_size, self.n_sa, self.sa_intv = int(info[6]), int(info[7]), int(info[8]), int(info[9])
        self.bwt = np.ctypeslib.as_array(ref_lib().csref_index_bwt(self.h), shape=(self.bwt_size,))
        self.sa = np.ctypeslib.as_array(ref_lib().csref_index_sa(self.h), shape=(self.n_sa,))

    @classmethod
    def load(cls, prefix: str) -> "RefIndex":
        return cls(ref_lib().csref_index_load(prefix.encode()))

    @classmethod
    def from_arrays(cls, primary, L2, seq_len, bwt, sa, sa_intv) -> "RefIndex":
        L2 = np.ascontiguousarray(L2, dtype=np.uint64)
        bwt = np.ascontiguousarray(bwt, dtype=np.uint32)
        sa = np.ascontiguousarray(sa, dtype=np.uint64)
        h = ref_lib().csref_index_from_arrays(primary, _ptr(L2), seq_len, _ptr(bwt), bwt.shape[0], _ptr(sa), sa.shape[0], sa_intv)
        return cls(h, keep=(L2, bwt, sa))

    def __del__(self):
        try:
            ref_lib().csref_index_free(self.h)
        except Exception:
            pass

    def occ4(self, k):
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty((k.shape[0], 4), dtype=np.uint64)
        ref_lib().csref_occ4(self.h, k.shape[0], _ptr(k), _ptr(out))
        return out

    def extend(self, ik, is_back):
        ik = np.ascontiguousarray(ik, dtype=np.uint64)
        is_back = np.ascontiguousarray(is_back, dtype=np.int32)
        out = np.empty((ik.shape[0], 4, 3), dtype=np.uint64)
        ref_lib().csref_extend(self.h, ik.shape[0], _ptr(ik), _ptr(is_back), _ptr(out))
        return out

    def sa_lookup(self, k):
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty(k.shape[0], dtype=np.uint64)
        ref_lib().csref_sa(self.h, k.shape[0], _ptr(k), _ptr(out))
        return out

    def seed(self, bases, off, mode: str = "bwamem", min_seed_len=19, split_factor=1.5, split_width=10,
             max_mem_intv=20, max_occ=500, n_threads: int = 1) -> SeedResult:
        """mode 'bwamem': bwt_smem1/bwt_seed_strategy1/bwt_sa; mode 'compseed': the SST path."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n = off.shape[0] - 1
        opt = _RefOpt(min_seed_len, split_factor, split_width, max_mem_intv, max_occ)
        L = ref_lib()
        r = L.csref_seed(self.h, 0 if mode == "bwamem" else 1, n_threads, n, _ptr(bases), _ptr(off), C.byref(opt))
        try:
            res = _copy_result(L, "csref", r, n)
            cnt = np.zeros(4, dtype=np.int64)
            L.csref_result_counters(r, _ptr(cnt))
            res.counters = dict(zip(["ext_queries", "ext_calls", "sal_queries", "sal_calls"], cnt.tolist()))
        finally:
            L.csref_result_free(r)
        return res


def bwaidx(fasta: str, prefix: str) -> None:
    """Run the reference's own index builder (oracle/_ref/bwaidx)."""
    subprocess.check_call([os.path.join(REF_BIN, "bwaidx"), "-p", prefix, fasta],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


# ---------------------------------------------------------------------------------------------
# Chains: the reference's own mem_chain + mem_chain_flt on given mems / seeds
# ---------------------------------------------------------------------------------------------
class _ChainOpt(C.Structure):
    _fields_ = [("w", C.c_int32), ("max_chain_gap", C.c_int32), ("min_chain_weight", C.c_int32), ("max_chain_extend", C.c_int32),
                ("mask_level", C.c_float), ("drop_ratio", C.c_float)]


@dataclass
class ChainResult:
    chain_off: np.ndarray   # u32 [n_reads+1]
    pos: np.ndarray         # i64 [n_chains]
    rid: np.ndarray         # i32
    w: np.ndarray           # i32
    kept: np.ndarray        # i32  kept | is_alt << 8
    n: np.ndarray           # i32  seeds in the chain
    s_rbeg: np.ndarray      # i64 [n_chain_seeds], chain after chain
    s_qbeg: np.ndarray      # i32
    s_len: np.ndarray       # i32
    frac_rep: np.ndarray    # f32 [n_reads]


def ref_chain(read_off, res: SeedResult, contig_lens, min_seed_len=19, split_factor=1.5, split_width=10, max_mem_intv=20, max_occ=500,
              w=100, max_chain_gap=10000, min_chain_weight=0, max_chain_extend=1 << 30, mask_level=0.5, drop_ratio=0.5, is_alt=None) -> ChainResult:
    """mem_chain + mem_chain_flt of the unmodified reference (comp_seed.cpp:241-354) on the mems / seeds of `res`.
    contig_lens: lengths of the reference sequences (their sum is l_pac).  Defaults: mem_opt_init, comp_seed.cpp:26-61."""
    L = ref_lib()
    L.csref_chain.restype = C.c_void_p
    L.csref_chain.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.POINTER(_RefOpt), C.POINTER(_ChainOpt), C.c_int, C.c_void_p, C.c_void_p]
    L.csref_chains_n.restype = C.c_uint64; L.csref_chains_n.argtypes = [C.c_void_p]
    L.csref_chains_n_seeds.restype = C.c_uint64; L.csref_chains_n_seeds.argtypes = [C.c_void_p]
    L.csref_chains_copy.argtypes = [C.c_void_p] * 11
    L.csref_chains_free.argtypes = [C.c_void_p]
    read_off = np.ascontiguousarray(read_off, dtype=np.uint32)
    mem_off = np.ascontiguousarray(res.mem_off, dtype=np.uint32); mems = np.ascontiguousarray(res.mems, dtype=np.uint64)
    seed_off = np.ascontiguousarray(res.seed_off, dtype=np.uint32); rbeg = np.ascontiguousarray(res.rbeg, dtype=np.int64)
    lens = np.ascontiguousarray(contig_lens, dtype=np.int32)
    alt = np.ascontiguousarray(is_alt if is_alt is not None else np.zeros(lens.shape[0]), dtype=np.uint8)
    n = read_off.shape[0] - 1
    so = _RefOpt(min_seed_len, split_factor, split_width, max_mem_intv, max_occ)
    co = _ChainOpt(w, max_chain_gap, min_chain_weight, max_chain_extend, mask_level, drop_ratio)
    h = L.csref_chain(n, _ptr(read_off), _ptr(mem_off), _ptr(mems), _ptr(seed_off), _ptr(rbeg), C.byref(so), C.byref(co), lens.shape[0], _ptr(lens), _ptr(alt))
    try:
        nc, ns = int(L.csref_chains_n(h)), int(L.csref_chains_n_seeds(h))
        out = ChainResult(np.empty(n + 1, np.uint32), np.empty(nc, np.int64), np.empty(nc, np.int32), np.empty(nc, np.int32), np.empty(nc, np.int32),
                          np.empty(nc, np.int32), np.empty(ns, np.int64), np.empty(ns, np.int32), np.empty(ns, np.int32), np.empty(n, np.float32))
        L.csref_chains_copy(h, _ptr(out.chain_off), _ptr(out.pos), _ptr(out.rid), _ptr(out.w), _ptr(out.kept), _ptr(out.n),
                            _ptr(out.s_rbeg), _ptr(out.s_qbeg), _ptr(out.s_len), _ptr(out.frac_rep))
    finally:
        L.csref_chains_free(h)
    return out


# ---- banded Smith-Waterman extension ----
def bsw_mat(a: int = 1, b: int = 4, ambig: int = -1) -> np.ndarray:
    """bwa_fill_scmat (bwalib/bwa.c:419-431)."""
    m = np.full((5, 5), -b, dtype=np.int8)
    np.fill_diagonal(m, a)
    m[4, :] = ambig
    m[:, 4] = ambig
    return np.ascontiguousarray(m.reshape(25))


def oracle_bsw(pairs: np.ndarray, seq_buf_ref: np.ndarray, seq_buf_qer: np.ndarray, w: int = 100, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100,
               end_bonus=5, mat: np.ndarray | None = None, n_threads: int = 1):
    """Our C restatement of ksw_extend2 over a batch (oracle/cs_oracle.c: cso_bsw_extend).  Returns (pairs copy with results, cells)."""
    L = lib()
    L.cso_bsw_extend.restype = C.c_uint64
    L.cso_bsw_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
    out = np.ascontiguousarray(pairs, dtype=np.int32).copy()
    m = np.ascontiguousarray(mat if mat is not None else bsw_mat(), dtype=np.int8)
    ref = np.ascontiguousarray(seq_buf_ref, dtype=np.uint8); qer = np.ascontiguousarray(seq_buf_qer, dtype=np.uint8)
    cells = L.cso_bsw_extend(_ptr(out), _ptr(ref), _ptr(qer), out.shape[0], w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, _ptr(m), n_threads)
    return out, int(cells)


def ref_bsw(pairs: np.ndarray, seq_buf_ref: np.ndarray, seq_buf_qer: np.ndarray, w: int = 100, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100,
            end_bonus=5, a: int = 1, b: int = 4, mode: int = 0):
    """The unmodified reference: BandedPairWiseSW::scalarBandedSWAWrapper (mode 0), getScores8 (1), getScores16 (2)
    (mapping/bandedSWA.cpp).  Returns (pairs copy with results, seconds)."""
    L = ref_lib()
    L.csref_bsw.restype = C.c_double
    L.csref_bsw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_int] * 6 + [C.c_void_p, C.c_int, C.c_int, C.c_int]
    out = np.ascontiguousarray(pairs, dtype=np.int32).copy()
    m = bsw_mat(a, b)
    ref = np.ascontiguousarray(seq_buf_ref, dtype=np.uint8); qer = np.ascontiguousarray(seq_buf_qer, dtype=np.uint8)
    # the SIMD twins read whole vectors past the last base of a sequence: pad the buffers as the caller's are (comp_seed.h: seqBuf sizes)
    ref = np.concatenate([ref, np.zeros(1024, np.uint8)]); qer = np.concatenate([qer, np.zeros(1024, np.uint8)])
    sec = L.csref_bsw(_ptr(out), _ptr(ref), _ptr(qer), out.shape[0], w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, _ptr(m), a, b, mode)
    return out, float(sec)


def ref_ksw_extend2(query: np.ndarray, target: np.ndarray, h0: int, w: int = 100, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, a=1, b=4):
    """ksw_extend2 of the reference (bwalib/ksw.c:380) for one pair: (score, qle, tle, gtle, gscore, max_off)."""
    L = ref_lib()
    L.csref_ksw_extend2.restype = C.c_int
    L.csref_ksw_extend2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
    out5 = np.zeros(5, np.int32)
    m = bsw_mat(a, b)
    q = np.ascontiguousarray(query, np.uint8); t = np.ascontiguousarray(np.concatenate([target, np.zeros(1, np.uint8)]), np.uint8)
    sc = L.csref_ksw_extend2(q.shape[0], _ptr(q), target.shape[0], _ptr(t), _ptr(m), o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, _ptr(out5))
    return (int(sc),) + tuple(int(x) for x in out5)

"""Host-side mirror of the reference's seeding interface, bound to the C-ABI in
include/compseed_b200.h through ctypes.

Names follow the reference: an FM-index (`bwt_t`, FM_index/bwt.h:48-60) with `occ4` / `extend` /
`sa` queries (bwt_occ4 / bwt_extend / bwt_sa), seeding options that are the -k/-r/-s/-y/-c fields
of `mem_opt_t` (mapping/comp_seed.h:41-73), and a batch call that returns, per read, the sorted mems
(`aux.match[r]`, comp_seed.cpp:2261-2301 == `aux->mem`, bwamem.c:218-272) and the resolved seed
positions in emission order (comp_seed.cpp:2306-2346 == bwamem.c:386-399).

There is NO CPU fallback: if libcompseed_b200.so is missing or no CUDA device is visible, every
call raises.  PyTorch is not used here at all; numpy arrays are the host buffers.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import synth

_HERE = os.path.dirname(os.path.abspath(__file__))
# COMPSEED_LIB_TAG selects an experiment build of the same sources (see build.py); default: the product library
_TAG = os.environ.get("COMPSEED_LIB_TAG", "")
LIB_PATH = os.path.join(_HERE, "_lib", f"libcompseed_b200{('_' + _TAG) if _TAG else ''}.so")

CS_OK, CS_E_ARG, CS_E_CUDA, CS_E_OVERFLOW, CS_E_IO, CS_E_NODEVICE, CS_E_STATE, CS_E_READ_OVERFLOW = 0, -1, -2, -3, -4, -5, -6, -7


class CompSeedError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"compseed_b200 error {code}: {msg}")
        self.code = code


class _BwtView(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
                ("bwt", C.c_void_p), ("sa_intv", C.c_int32), ("n_sa", C.c_uint64), ("sa", C.c_void_p)]


class _SeedOpt(C.Structure):
    _fields_ = [("min_seed_len", C.c_int32), ("split_len", C.c_int32), ("split_width", C.c_int32),
                ("max_mem_intv", C.c_int32), ("max_occ", C.c_int32)]


class _Counters(C.Structure):
    _fields_ = [("ext_queries", C.c_uint64), ("ext_calls", C.c_uint64), ("sal_queries", C.c_uint64), ("sal_calls", C.c_uint64)]


class _Result(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_mems", C.c_uint64), ("n_seeds", C.c_uint64),
                ("mem_off", C.POINTER(C.c_uint32)), ("mems", C.POINTER(C.c_uint64)),
                ("seed_off", C.POINTER(C.c_uint32)), ("rbeg", C.POINTER(C.c_int64)),
                ("counters", _Counters), ("kernel_ms", C.c_float * 8), ("n_deferred", C.c_uint64),
                ("gather_requests", C.c_uint64 * 6), ("kernel_ms2", C.c_float * 4)]


class _Block(C.Structure):
    _fields_ = [("r0", C.c_uint64), ("r1", C.c_uint64), ("batch_reads", C.c_uint32), ("n_batches", C.c_uint32),
                ("mem_base", C.POINTER(C.c_uint64)), ("seed_base", C.POINTER(C.c_uint64)),
                ("mem_off", C.POINTER(C.c_uint32)), ("seed_off", C.POINTER(C.c_uint32)),
                ("cmems", C.POINTER(C.c_uint32)), ("rbeg_lo", C.POINTER(C.c_uint32)), ("rbeg_hi", C.POINTER(C.c_uint8)),
                ("device", C.c_int), ("chains", C.POINTER(C.c_uint32)), ("qbeg", C.POINTER(C.c_uint16)), ("len", C.POINTER(C.c_uint16))]


class _MultiResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_mems", C.c_uint64), ("n_seeds", C.c_uint64), ("n_blocks", C.c_int),
                ("blocks", C.POINTER(_Block)), ("counters", _Counters), ("seconds", C.c_double), ("host_s", C.c_double * 3), ("gpu_ms", C.c_double * 4)]


class _BnsView(C.Structure):
    _fields_ = [("l_pac", C.c_int64), ("n_seqs", C.c_int32), ("offset", C.c_void_p), ("is_alt", C.c_void_p)]


class _ChainOpt(C.Structure):
    _fields_ = [("w", C.c_int32), ("max_chain_gap", C.c_int32), ("min_chain_weight", C.c_int32), ("max_chain_extend", C.c_int32),
                ("mask_level", C.c_float), ("drop_ratio", C.c_float)]


class _ChainResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_chains", C.c_uint64), ("n_cseeds", C.c_uint64),
                ("chain_off", C.POINTER(C.c_uint32)), ("cseed_off", C.POINTER(C.c_uint32)), ("chains", C.POINTER(C.c_uint32)),
                ("rbeg_lo", C.POINTER(C.c_uint32)), ("rbeg_hi", C.POINTER(C.c_uint8)), ("qbeg", C.POINTER(C.c_uint16)), ("len", C.POINTER(C.c_uint16))]


@dataclass
class ChainOpt:
    """The chaining scalars of mem_opt_t with the defaults of mem_opt_init (comp_seed.cpp:26-61)."""
    w: int = 100
    max_chain_gap: int = 10000
    min_chain_weight: int = 0
    max_chain_extend: int = 1 << 30
    mask_level: float = 0.5
    drop_ratio: float = 0.5

    def _c(self) -> _ChainOpt:
        return _ChainOpt(self.w, self.max_chain_gap, self.min_chain_weight, self.max_chain_extend, self.mask_level, self.drop_ratio)


@dataclass
class ChainResult:
    """What mem_chain + mem_chain_flt return for a batch (mem_chain_v per read, flattened)."""
    chain_off: np.ndarray   # u32 [n+1]
    cseed_off: np.ndarray   # u32 [n+1]
    rid: np.ndarray         # i32 [n_chains]
    w: np.ndarray           # i32
    kept: np.ndarray        # i32  kept | is_alt << 8
    n: np.ndarray           # i32  seeds per chain
    l_rep: np.ndarray       # u32 per chain (frac_rep = l_rep / l_seq)
    s_rbeg: np.ndarray      # i64 [n_cseeds]
    s_qbeg: np.ndarray      # i32
    s_len: np.ndarray       # i32
    wire_bytes: int = 0


class _IndexConfig(C.Structure):
    _fields_ = [("kmer_table_depth", C.c_int32), ("prune_k", C.c_int32), ("isa_intv", C.c_int32), ("repeat_lengths", C.c_int32)]


class _CtxConfig(C.Structure):
    _fields_ = [("use_fast", C.c_int32), ("use_r3_fast", C.c_int32), ("defer_cap", C.c_int32), ("lit_ctas_per_sm", C.c_int32),
                ("prefetch_results", C.c_int32), ("l2_persist_mb", C.c_int32), ("overlap_streams", C.c_int32), ("compact_results", C.c_int32), ("batch_order", C.c_int32)]


class _BswOpt(C.Structure):
    _fields_ = [("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("zdrop", C.c_int32),
                ("end_bonus", C.c_int32), ("mat", C.c_int8 * 25)]


class _CompactResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_mems", C.c_uint64), ("n_seeds", C.c_uint64),
                ("mem_off", C.POINTER(C.c_uint32)), ("cmems", C.POINTER(C.c_uint32)),
                ("seed_off", C.POINTER(C.c_uint32)), ("rbeg_lo", C.POINTER(C.c_uint32)), ("rbeg_hi", C.POINTER(C.c_uint8)),
                ("counters", _Counters)]


@dataclass
class IndexConfig:
    """cs_index_config_t: the result-neutral structures built next to the index (-1 = default for the index size)."""
    kmer_table_depth: int = -1
    prune_k: int = -1
    isa_intv: int = -1
    repeat_lengths: int = -1

    def _c(self) -> _IndexConfig:
        return _IndexConfig(self.kmer_table_depth, self.prune_k, self.isa_intv, self.repeat_lengths)


@dataclass
class CtxConfig:
    """cs_ctx_config_t: execution switches of a context (none of them changes a result)."""
    use_fast: int = -1
    use_r3_fast: int = -1
    defer_cap: int = -1
    lit_ctas_per_sm: int = -1
    prefetch_results: int = 0
    l2_persist_mb: int = 0
    overlap_streams: int = 0
    compact_results: int = 0
    batch_order: int = -1

    def _c(self) -> _CtxConfig:
        return _CtxConfig(self.use_fast, self.use_r3_fast, self.defer_cap, self.lit_ctas_per_sm, self.prefetch_results,
                          self.l2_persist_mb, self.overlap_streams, self.compact_results, self.batch_order)


_lib = None


def load_library():
    """dlopen the in-tree CUDA library.  Fails loudly: there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m compseed_b200.build` "
                           "(nvcc, sm_100a).  compseed_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.cs_last_error.restype = C.c_char_p
    L.cs_last_error_code.restype = C.c_int
    L.cs_device_count.restype = C.c_int
    L.cs_index_upload.restype = C.c_void_p
    L.cs_index_upload.argtypes = [C.POINTER(_BwtView), C.c_int, C.c_int]
    L.cs_index_load.restype = C.c_void_p
    L.cs_index_load.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.cs_index_build.restype = C.c_void_p
    L.cs_index_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    L.cs_index_upload_ex.restype = C.c_void_p
    L.cs_index_upload_ex.argtypes = [C.POINTER(_BwtView), C.c_int, C.c_int, C.POINTER(_IndexConfig)]
    L.cs_index_load_ex.restype = C.c_void_p
    L.cs_index_load_ex.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(_IndexConfig)]
    L.cs_index_build_ex.restype = C.c_void_p
    L.cs_index_build_ex.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(_IndexConfig)]
    L.cs_index_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.cs_index_verify.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
    L.cs_probe_index_gather.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.cs_ctx_create_ex.restype = C.c_void_p
    L.cs_ctx_create_ex.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(_CtxConfig)]
    L.cs_ctx_need.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.cs_seed_batch_wait_compact.argtypes = [C.c_void_p, C.c_int, C.POINTER(_CompactResult)]
    L.cs_compact_expand.argtypes = [C.POINTER(_CompactResult), C.c_void_p, C.c_void_p, C.c_int]
    L.cs_multi_create.restype = C.c_void_p
    L.cs_multi_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(_CtxConfig)]
    L.cs_multi_free.argtypes = [C.c_void_p]
    L.cs_index_replicate.restype = C.c_void_p
    L.cs_index_replicate.argtypes = [C.c_void_p, C.c_int]
    L.cs_multi_submit.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(_SeedOpt)]
    L.cs_multi_submit_packed.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_SeedOpt)]
    L.cs_multi_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(_MultiResult)]
    L.cs_multi_gather.argtypes = [C.POINTER(_MultiResult), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.cs_pack_reads_host64.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.cs_multi_set_chaining.argtypes = [C.c_void_p, C.POINTER(_BnsView), C.POINTER(_ChainOpt)]
    L.cs_multi_launches.restype = C.c_uint64
    L.cs_multi_launches.argtypes = [C.c_void_p]
    L.cs_multi_trace.restype = C.c_uint32
    L.cs_multi_trace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
    L.cs_ctx_set_chaining.argtypes = [C.c_void_p, C.POINTER(_BnsView), C.POINTER(_ChainOpt)]
    L.cs_seed_batch_wait_chains.argtypes = [C.c_void_p, C.c_int, C.POINTER(_ChainResult)]
    L.cs_ctx_launches.restype = C.c_uint64
    L.cs_ctx_launches.argtypes = [C.c_void_p]
    L.cs_index_download.argtypes = [C.c_void_p, C.POINTER(_BwtView), C.c_void_p, C.c_void_p, C.c_int]
    L.cs_index_info.argtypes = [C.c_void_p, C.POINTER(_BwtView), C.POINTER(C.c_uint64)]
    L.cs_index_free.argtypes = [C.c_void_p]
    L.cs_occ4.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.cs_extend.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cs_sa.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.cs_ctx_create.restype = C.c_void_p
    L.cs_ctx_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int]
    L.cs_ctx_free.argtypes = [C.c_void_p]
    L.cs_seed_batch_submit.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(_SeedOpt)]
    L.cs_seed_batch_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Result)]
    L.cs_seed_batch_submit_packed.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_SeedOpt)]
    L.cs_pack_reads_host.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.cs_packed_words.argtypes = [C.c_uint32, C.c_void_p]
    L.cs_packed_words.restype = C.c_uint64
    L.cs_seed_batch_stage.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
    L.cs_seed_batch_run_staged.argtypes = [C.c_void_p, C.c_int, C.POINTER(_SeedOpt)]
    L.cs_seed_batch_wait_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Result)]
    L.cs_seed_batch_fetch.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Result)]
    L.cs_probe_random_gather.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.cs_probe_random_gather_ex.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.cs_flush_l2.argtypes = [C.c_int]
    L.cs_debug_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.cs_bsw_create.restype = C.c_void_p
    L.cs_bsw_create.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32]
    L.cs_bsw_free.argtypes = [C.c_void_p]
    L.cs_bsw_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.POINTER(_BswOpt)]
    L.cs_bsw_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32]
    L.cs_bsw_run_staged.argtypes = [C.c_void_p, C.c_int32, C.POINTER(_BswOpt), C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    L.cs_bsw_fetch.argtypes = [C.c_void_p, C.c_void_p]
    L.cs_bsw_launches.restype = C.c_uint64
    L.cs_bsw_launches.argtypes = [C.c_void_p]
    L.cs_bsw_set_ctas_per_sm.argtypes = [C.c_void_p, C.c_int]
    L.cs_bsw_set_rows_in_smem.argtypes = [C.c_void_p, C.c_int]
    L.cs_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.cs_host_unregister.argtypes = [C.c_void_p]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != CS_OK:
        raise CompSeedError(rc, load_library().cs_last_error().decode())


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return load_library().cs_device_count()


@dataclass
class SeedOpt:
    """-k / -r / -s / -y / -c of mem_opt_t with the defaults of mem_opt_init (comp_seed.cpp:26-61)."""
    min_seed_len: int = 19
    split_factor: float = 1.5
    split_width: int = 10
    max_mem_intv: int = 20
    max_occ: int = 500
    caller: str = "bwamem"   # whose rounding of split_len to reproduce: "bwamem" (bwamem.c:223) or "compseed" (comp_seed.cpp:2279)

    @property
    def split_len(self) -> int:
        f = synth.split_len_bwamem if self.caller == "bwamem" else synth.split_len_compseed
        return f(self.min_seed_len, self.split_factor)

    def _c(self) -> _SeedOpt:
        return _SeedOpt(self.min_seed_len, self.split_len, self.split_width, self.max_mem_intv, self.max_occ)


@dataclass
class SeedResult:
    mem_off: np.ndarray      # u32 [n+1]
    mems: np.ndarray         # u64 [n_mems, 4]: x0 (k), x1 (l), x2 (s), info = start << 32 | end
    seed_off: np.ndarray     # u32 [n+1]
    rbeg: np.ndarray         # i64 [n_seeds]
    counters: dict = field(default_factory=dict)
    kernel_ms: tuple = (0.0,) * 8
    gather_requests: tuple = (0,) * 6   # executed memory requests per kernel (cs_result_t.gather_requests)
    kernel_ms2: tuple = (0.0,) * 4

    @property
    def n_reads(self) -> int:
        return self.mem_off.shape[0] - 1


class FMIndex:
    """GPU-resident FM-index (device twin of bwt_t)."""

    def __init__(self, handle, device: int):
        if not handle:
            raise CompSeedError(load_library().cs_last_error_code() or CS_E_CUDA, load_library().cs_last_error().decode())
        self.h = C.c_void_p(handle)
        self.device = device
        v = _BwtView()
        nbytes = C.c_uint64()
        _check(load_library().cs_index_info(self.h, C.byref(v), C.byref(nbytes)))
        self.primary, self.seq_len, self.bwt_size = int(v.primary), int(v.seq_len), int(v.bwt_size)
        self.L2 = np.array(list(v.L2), dtype=np.uint64)
        self.sa_intv, self.n_sa = int(v.sa_intv), int(v.n_sa)
        self.device_bytes = int(nbytes.value)

    @classmethod
    def upload(cls, primary: int, L2, seq_len: int, bwt: np.ndarray, sa: np.ndarray, sa_intv: int,
               device: int = 0, dense_sa_intv: int = 0, config: "IndexConfig | None" = None) -> "FMIndex":
        """From arrays in the reference's in-memory layout (what bwt_restore_bwt/sa produce)."""
        bwt = np.ascontiguousarray(bwt, dtype=np.uint32)
        sa = np.ascontiguousarray(sa, dtype=np.uint64)
        v = _BwtView()
        v.primary, v.seq_len, v.bwt_size = int(primary), int(seq_len), int(bwt.shape[0])
        for i in range(5):
            v.L2[i] = int(L2[i])
        v.bwt, v.sa_intv, v.n_sa, v.sa = _ptr(bwt), int(sa_intv), int(sa.shape[0]), _ptr(sa)
        cfg = (config or IndexConfig())._c()
        return cls(load_library().cs_index_upload_ex(C.byref(v), device, dense_sa_intv, C.byref(cfg)), device)

    @classmethod
    def load(cls, prefix: str, device: int = 0, dense_sa_intv: int = 0, config: "IndexConfig | None" = None) -> "FMIndex":
        """From P.bwt / P.sa written by bwaidx (bwt_restore_bwt / bwt_restore_sa, bwt.c:421-462)."""
        cfg = (config or IndexConfig())._c()
        return cls(load_library().cs_index_load_ex(prefix.encode(), device, dense_sa_intv, C.byref(cfg)), device)

    @classmethod
    def build(cls, fwd: np.ndarray, device: int = 0, sa_intv: int = 32, config: "IndexConfig | None" = None) -> "FMIndex":
        """Construct the index of fwd + revcomp(fwd) on the GPU (what bwaidx computes on the CPU)."""
        fwd = np.ascontiguousarray(fwd, dtype=np.uint8)
        cfg = (config or IndexConfig())._c()
        return cls(load_library().cs_index_build_ex(_ptr(fwd), fwd.shape[0], device, sa_intv, C.byref(cfg)), device)

    def write(self, prefix: str, sa_intv: int = 32) -> None:
        """P.bwt / P.sa in the reference's on-disk format (bwt_dump_bwt / bwt_dump_sa, bwt.c:385-407)."""
        _check(load_library().cs_index_write(self.h, prefix.encode(), sa_intv))

    VERIFY_FIELDS = ("order_rows", "order_bad", "bwt_rows", "bwt_bad", "perm_bad", "occ_bad", "isa_samples", "isa_bad",
                     "filter_samples", "filter_bad", "table_samples", "table_bad", "text_bases", "text_bad", "rep_samples", "rep_bad")

    def verify(self, fwd: np.ndarray | None = None, stride: int = 1) -> dict:
        """cs_index_verify: the index checked against the definitions of its parts (needs the dense SA)."""
        out = np.zeros(16, dtype=np.uint64)
        if fwd is not None:
            fwd = np.ascontiguousarray(fwd, dtype=np.uint8)
        _check(load_library().cs_index_verify(self.h, _ptr(fwd) if fwd is not None else None, 0 if fwd is None else fwd.shape[0], stride, _ptr(out)))
        d = {k: int(v) for k, v in zip(self.VERIFY_FIELDS, out)}
        d["ok"] = all(d[k] == 0 for k in d if k.endswith("_bad")) and d["order_rows"] > 0
        return d

    def probe_gather(self, n_loads: int = 1 << 28, iters: int = 2, unroll: int = 4):
        """(GB/s, Gloads/s) of random 32-byte sector loads over this index's own arrays (cs_probe_index_gather)."""
        gb, gl = C.c_double(), C.c_double()
        _check(load_library().cs_probe_index_gather(self.h, n_loads, iters, unroll, C.byref(gb), C.byref(gl)))
        return gb.value, gl.value

    def download(self, sa_intv: int = 32):
        """Back to the reference layout: dict(primary, L2, seq_len, bwt, sa, sa_intv)."""
        L = load_library()
        v = _BwtView()
        _check(L.cs_index_download(self.h, C.byref(v), None, None, sa_intv))
        bwt = np.zeros(int(v.bwt_size), dtype=np.uint32)
        sa = np.zeros(int(v.n_sa), dtype=np.uint64)
        _check(L.cs_index_download(self.h, C.byref(v), _ptr(bwt), _ptr(sa), sa_intv))
        return dict(primary=self.primary, L2=self.L2.copy(), seq_len=self.seq_len, bwt=bwt, sa=sa, sa_intv=sa_intv)

    def close(self) -> None:
        if self.h:
            load_library().cs_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- unit-level queries (bwt_occ4 / bwt_extend / bwt_sa) ----------------------------------
    def occ4(self, k) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty((k.shape[0], 4), dtype=np.uint64)
        _check(load_library().cs_occ4(self.h, k.shape[0], _ptr(k), _ptr(out)))
        return out

    def extend(self, ik, is_back) -> np.ndarray:
        ik = np.ascontiguousarray(ik, dtype=np.uint64)
        is_back = np.ascontiguousarray(is_back, dtype=np.int32)
        out = np.empty((ik.shape[0], 4, 3), dtype=np.uint64)
        _check(load_library().cs_extend(self.h, ik.shape[0], _ptr(ik), _ptr(is_back), _ptr(out)))
        return out

    def sa(self, k) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.empty(k.shape[0], dtype=np.uint64)
        _check(load_library().cs_sa(self.h, k.shape[0], _ptr(k), _ptr(out)))
        return out


class SeedContext:
    """Slots of pinned + device buffers with one CUDA stream each (the kt_for worker pool's
    replacement, bwamem.c:1343 / comp_seed.cpp:2541-2548)."""

    def __init__(self, index: FMIndex, max_reads: int, max_bases: int, max_read_len: int = 256,
                 max_mems: int = 0, max_seeds: int = 0, n_slots: int = 2, config: "CtxConfig | None" = None):
        self.index = index
        self.max_reads, self.max_bases, self.max_read_len = max_reads, max_bases, max_read_len
        self.n_slots = n_slots
        cfg = (config or CtxConfig())._c()
        h = load_library().cs_ctx_create_ex(index.h, max_reads, max_bases, max_read_len, max_mems, max_seeds, n_slots, C.byref(cfg))
        if not h:
            raise CompSeedError(load_library().cs_last_error_code() or CS_E_CUDA, load_library().cs_last_error().decode())
        self.h = C.c_void_p(h)

    def close(self) -> None:
        if self.h:
            load_library().cs_ctx_free(self.h)
            self.h = None

    def need(self, slot: int):
        """After CS_E_OVERFLOW on `slot`: (max_mems, max_seeds) that batch needs (cs_ctx_need)."""
        m, sd = C.c_uint64(), C.c_uint64()
        _check(load_library().cs_ctx_need(self.h, slot, C.byref(m), C.byref(sd)))
        return int(m.value), int(sd.value)

    @property
    def launches(self) -> int:
        """Kernels launched on behalf of this context so far (cs_ctx_launches)."""
        return int(load_library().cs_ctx_launches(self.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _prep(bases, off):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        return bases, off

    def submit(self, slot: int, bases, off, opt: SeedOpt) -> None:
        bases, off = self._prep(bases, off)
        o = opt._c()
        _check(load_library().cs_seed_batch_submit(self.h, slot, off.shape[0] - 1, _ptr(bases), _ptr(off), C.byref(o)))

    def submit_packed(self, slot: int, packed, nmask, off, opt: SeedOpt) -> None:
        """Reads already 2-bit packed by the caller (pack_reads): cs_seed_batch_submit_packed."""
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        nmask = np.ascontiguousarray(nmask, dtype=np.uint32)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        o = opt._c()
        _check(load_library().cs_seed_batch_submit_packed(self.h, slot, off.shape[0] - 1, _ptr(packed), _ptr(nmask), _ptr(off), C.byref(o)))

    def stage(self, slot: int, bases, off) -> None:
        bases, off = self._prep(bases, off)
        _check(load_library().cs_seed_batch_stage(self.h, slot, off.shape[0] - 1, _ptr(bases), _ptr(off)))

    def run_staged(self, slot: int, opt: SeedOpt) -> None:
        o = opt._c()
        _check(load_library().cs_seed_batch_run_staged(self.h, slot, C.byref(o)))

    def wait_device(self, slot: int) -> SeedResult:
        r = _Result()
        _check(load_library().cs_seed_batch_wait_device(self.h, slot, C.byref(r)))
        return self._result(r, copy=False)

    def debug_stats(self, slot: int) -> np.ndarray:
        """Raw device counters of the last finished run on `slot` (cs_debug_stats)."""
        out = np.zeros(40, dtype=np.uint64)
        _check(load_library().cs_debug_stats(self.h, slot, _ptr(out)))
        return out

    def fetch(self, slot: int, copy: bool = True) -> SeedResult:
        r = _Result()
        _check(load_library().cs_seed_batch_fetch(self.h, slot, C.byref(r)))
        return self._result(r, copy=copy)

    def wait(self, slot: int, copy: bool = True) -> SeedResult:
        r = _Result()
        _check(load_library().cs_seed_batch_wait(self.h, slot, C.byref(r)))
        return self._result(r, copy=copy)

    def set_chaining(self, contig_lens=None, opt: "ChainOpt | None" = None, is_alt=None) -> None:
        """cs_ctx_set_chaining: batches submitted from now on are also chained on the device (mem_chain + mem_chain_flt).
        contig_lens: lengths of the reference sequences in order (None switches chaining off)."""
        if contig_lens is None:
            _check(load_library().cs_ctx_set_chaining(self.h, None, None))
            return
        lens = np.asarray(contig_lens, dtype=np.int64)
        offs = np.ascontiguousarray(np.concatenate([[0], np.cumsum(lens)[:-1]]), dtype=np.int64)
        alt = np.ascontiguousarray(is_alt, dtype=np.uint8) if is_alt is not None else None
        v = _BnsView(int(lens.sum()), lens.shape[0], _ptr(offs), _ptr(alt) if alt is not None else None)
        o = (opt or ChainOpt())._c()
        _check(load_library().cs_ctx_set_chaining(self.h, C.byref(v), C.byref(o)))

    def wait_chains(self, slot: int) -> ChainResult:
        """cs_seed_batch_wait_chains: only the filtered chains of the batch cross the link."""
        r = _ChainResult()
        _check(load_library().cs_seed_batch_wait_chains(self.h, slot, C.byref(r)))
        n, nc, ns = int(r.n_reads), int(r.n_chains), int(r.n_cseeds)
        chain_off = np.ctypeslib.as_array(r.chain_off, shape=(n + 1,)).copy()
        cseed_off = np.ctypeslib.as_array(r.cseed_off, shape=(n + 1,)).copy()
        ch = np.ctypeslib.as_array(r.chains, shape=(nc, 4)).copy() if nc else np.empty((0, 4), dtype=np.uint32)
        if ns:
            lo = np.ctypeslib.as_array(r.rbeg_lo, shape=(ns,)).astype(np.int64)
            hi = np.ctypeslib.as_array(r.rbeg_hi, shape=(ns,)).astype(np.int64)
            rb = ((lo | (hi << 32)) << 24) >> 24          # 40 bits, sign-extended (cs_crbeg)
            qb = np.ctypeslib.as_array(r.qbeg, shape=(ns,)).astype(np.int32)
            ln = np.ctypeslib.as_array(r.len, shape=(ns,)).astype(np.int32)
        else:
            rb, qb, ln = np.empty(0, np.int64), np.empty(0, np.int32), np.empty(0, np.int32)
        wk = ch[:, 1]
        return ChainResult(chain_off, cseed_off, ch[:, 0].astype(np.int32), (wk & 0x1fffffff).astype(np.int32),
                           (((wk >> 29) & 3) | ((wk >> 31) << 8)).astype(np.int32), ch[:, 2].astype(np.int32), ch[:, 3].copy(), rb, qb, ln,
                           wire_bytes=8 * (n + 1) + 16 * nc + 9 * ns)

    def wait_compact(self, slot: int, expand_threads: int = 0):
        """cs_seed_batch_wait_compact.  expand_threads == 0: the compact arrays as they arrived (views into the slot's
        pinned buffers: mem_off, cmems u32[n_mems, 5], seed_off, rbeg_lo, rbeg_hi); > 0: a SeedResult expanded by
        cs_compact_expand on that many host threads."""
        r = _CompactResult()
        _check(load_library().cs_seed_batch_wait_compact(self.h, slot, C.byref(r)))
        n, nm, ns = int(r.n_reads), int(r.n_mems), int(r.n_seeds)
        mem_off = np.ctypeslib.as_array(r.mem_off, shape=(n + 1,))
        seed_off = np.ctypeslib.as_array(r.seed_off, shape=(n + 1,))
        if expand_threads <= 0:
            cm = np.ctypeslib.as_array(r.cmems, shape=(nm, 5)) if nm else np.empty((0, 5), dtype=np.uint32)
            lo = np.ctypeslib.as_array(r.rbeg_lo, shape=(ns,)) if ns else np.empty(0, dtype=np.uint32)
            hi = np.ctypeslib.as_array(r.rbeg_hi, shape=(ns,)) if ns else np.empty(0, dtype=np.uint8)
            return mem_off, cm, seed_off, lo, hi
        mems = np.empty((nm, 4), dtype=np.uint64)
        rbeg = np.empty(ns, dtype=np.int64)
        _check(load_library().cs_compact_expand(C.byref(r), _ptr(mems), _ptr(rbeg), expand_threads))
        cnt = dict(ext_queries=int(r.counters.ext_queries), ext_calls=int(r.counters.ext_calls),
                   sal_queries=int(r.counters.sal_queries), sal_calls=int(r.counters.sal_calls))
        return SeedResult(mem_off.copy(), mems, seed_off.copy(), rbeg, cnt)

    @staticmethod
    def _result(r: _Result, copy: bool) -> SeedResult:
        n, nm, ns = int(r.n_reads), int(r.n_mems), int(r.n_seeds)
        cnt = dict(ext_queries=int(r.counters.ext_queries), ext_calls=int(r.counters.ext_calls),
                   sal_queries=int(r.counters.sal_queries), sal_calls=int(r.counters.sal_calls),
                   deferred_calls=int(r.n_deferred))
        ms = tuple(float(x) for x in r.kernel_ms)
        gr = tuple(int(x) for x in r.gather_requests)
        ms2 = tuple(float(x) for x in r.kernel_ms2)
        if not r.mem_off:  # device-resident result
            e = np.empty(0, dtype=np.uint32)
            res = SeedResult(e, np.empty((0, 4), dtype=np.uint64), e, np.empty(0, dtype=np.int64), cnt, ms, gr, ms2)
            res.n_mems_device, res.n_seeds_device, res.n_reads_device = nm, ns, n
            return res
        mem_off = np.ctypeslib.as_array(r.mem_off, shape=(n + 1,))
        seed_off = np.ctypeslib.as_array(r.seed_off, shape=(n + 1,))
        mems = np.ctypeslib.as_array(r.mems, shape=(nm, 4)) if nm else np.empty((0, 4), dtype=np.uint64)
        rbeg = np.ctypeslib.as_array(r.rbeg, shape=(ns,)) if ns else np.empty(0, dtype=np.int64)
        if copy:
            mem_off, seed_off, mems, rbeg = mem_off.copy(), seed_off.copy(), mems.copy(), rbeg.copy()
        return SeedResult(mem_off, mems, seed_off, rbeg, cnt, ms, gr, ms2)


def packed_words(off: np.ndarray) -> int:
    """Length of the packed / nmask arrays of a batch: (off[n] >> 5) + 2 n (== cs_packed_words)."""
    return (int(off[-1]) >> 5) + 2 * (off.shape[0] - 1)


def pack_reads(bases: np.ndarray, off: np.ndarray):
    """Host-side packing into the layout cs_seed_batch_submit_packed takes (include/compseed_b200.h): read r owns
    the words [(off[r] >> 5) + 2r, ... + (len_r >> 5) + 2); base j of a word at bits 2j of packed and bit j of nmask
    (set for codes > 3 and past the end of the read).  Pure numpy; no device involved."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint32).astype(np.int64)
    n = off.shape[0] - 1
    nw_tot = (int(off[-1]) >> 5) + 2 * n
    packed = np.zeros(nw_tot, dtype=np.uint64)
    nmask = np.full(nw_tot, 0xffffffff, dtype=np.uint32)
    if n == 0 or off[-1] == 0:
        return packed, nmask
    lens = np.diff(off)
    w0 = (off[:-1] >> 5) + 2 * np.arange(n, dtype=np.int64)         # first word of each read
    nw = (lens >> 5) + 2
    # words between a read's last word and the next read's first one stay "all N"; so do the pad words of a read
    read_of_base = np.repeat(np.arange(n, dtype=np.int64), lens)
    pos = np.arange(int(off[-1]), dtype=np.int64) - off[:-1][read_of_base]
    word = w0[read_of_base] + (pos >> 5)
    sh = (pos & 31).astype(np.uint64)
    amb = bases > 3
    code = np.where(amb, 0, bases).astype(np.uint64) << (np.uint64(2) * sh)
    np.bitwise_or.at(packed, word, code)
    clear = (~amb).astype(np.uint32) << sh.astype(np.uint32)
    keep = np.zeros(nw_tot, dtype=np.uint32)
    np.bitwise_or.at(keep, word, clear)
    nmask &= ~keep
    assert int((w0 + nw).max()) <= nw_tot
    return packed, nmask


class MultiSeeder:
    """cs_multi_t: one index replica and one context per device, reads of a set split into contiguous blocks (multiples
    of 512 reads, comp_seed.h:36) in input order, pipelined by one host thread per device; two sets may be in flight.
    Replaces kt_for over the reads of a -K batch (bwamem.c:1343, comp_seed.cpp:2541-2548) for the seeding part."""

    def __init__(self, indexes, batch_reads: int = 1 << 20, max_read_len: int = 256, n_slots: int = 3,
                 mems_per_read: int = 0, seeds_per_read: int = 0, config: "CtxConfig | None" = None):
        self.indexes = list(indexes)
        arr = (C.c_void_p * len(self.indexes))(*[i.h for i in self.indexes])
        cfg = (config or CtxConfig())._c()
        h = load_library().cs_multi_create(arr, len(self.indexes), batch_reads, max_read_len, n_slots, mems_per_read, seeds_per_read, C.byref(cfg))
        if not h:
            raise CompSeedError(load_library().cs_last_error_code() or CS_E_CUDA, load_library().cs_last_error().decode())
        self.h = C.c_void_p(h)
        self._keep = {}

    def close(self) -> None:
        if self.h:
            load_library().cs_multi_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(load_library().cs_multi_launches(self.h))

    def trace(self, set_id: int, dev: int = 0) -> np.ndarray:
        """Diagnostics timeline of the last finished run of a set on one device: [n_batches, 8] ms (cs_multi_trace)."""
        L = load_library()
        nb = int(L.cs_multi_trace(self.h, set_id, dev, None, 0))
        out = np.zeros((nb, 8), dtype=np.float32)
        if nb:
            L.cs_multi_trace(self.h, set_id, dev, _ptr(out), nb)
        return out

    def submit(self, set_id: int, bases, off64, opt: SeedOpt) -> None:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off64 = np.ascontiguousarray(off64, dtype=np.uint64)
        self._keep[set_id] = (bases, off64)
        o = opt._c()
        _check(load_library().cs_multi_submit(self.h, set_id, off64.shape[0] - 1, _ptr(bases), _ptr(off64), C.byref(o)))

    def submit_packed(self, set_id: int, packed, nmask, off64, opt: SeedOpt) -> None:
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        nmask = np.ascontiguousarray(nmask, dtype=np.uint32)
        off64 = np.ascontiguousarray(off64, dtype=np.uint64)
        self._keep[set_id] = (packed, nmask, off64)
        o = opt._c()
        _check(load_library().cs_multi_submit_packed(self.h, set_id, off64.shape[0] - 1, _ptr(packed), _ptr(nmask), _ptr(off64), C.byref(o)))

    def set_chaining(self, contig_lens=None, opt: "ChainOpt | None" = None, is_alt=None) -> None:
        """cs_multi_set_chaining: from the next set on the results are the filtered chains (None: mems + seed positions again)."""
        if contig_lens is None:
            _check(load_library().cs_multi_set_chaining(self.h, None, None))
            self._chaining = False
            return
        lens = np.asarray(contig_lens, dtype=np.int64)
        offs = np.ascontiguousarray(np.concatenate([[0], np.cumsum(lens)[:-1]]), dtype=np.int64)
        alt = np.ascontiguousarray(is_alt, dtype=np.uint8) if is_alt is not None else None
        v = _BnsView(int(lens.sum()), lens.shape[0], _ptr(offs), _ptr(alt) if alt is not None else None)
        o = (opt or ChainOpt())._c()
        _check(load_library().cs_multi_set_chaining(self.h, C.byref(v), C.byref(o)))
        self._chaining = True

    @staticmethod
    def _gather_chains(r: "_MultiResult") -> ChainResult:
        """Flat chains in input order from the blocks of a chained result (what cs_multi_read_chains walks), in numpy."""
        offs_c, offs_s, recs, lo, hi, qb, ln = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [], [], [], [], []
        cb = sb = 0
        for k in range(r.n_blocks):
            b = r.blocks[k]
            n = int(b.r1 - b.r0)
            if n == 0:
                continue
            B, nb = int(b.batch_reads), int(b.n_batches)
            cbase = np.ctypeslib.as_array(b.mem_base, shape=(nb + 1,)).astype(np.int64)
            sbase = np.ctypeslib.as_array(b.seed_base, shape=(nb + 1,)).astype(np.int64)
            co = np.ctypeslib.as_array(b.mem_off, shape=(nb, B + 1)).astype(np.int64)
            so = np.ctypeslib.as_array(b.seed_off, shape=(nb, B + 1)).astype(np.int64)
            for bi in range(nb):
                m = min(B, n - bi * B)
                offs_c.append(cb + cbase[bi] + co[bi, 1:m + 1]); offs_s.append(sb + sbase[bi] + so[bi, 1:m + 1])
            nc, ns = int(cbase[nb]), int(sbase[nb])
            if nc:
                recs.append(np.ctypeslib.as_array(b.chains, shape=(nc, 4)).copy())
            if ns:
                lo.append(np.ctypeslib.as_array(b.rbeg_lo, shape=(ns,)).astype(np.int64)); hi.append(np.ctypeslib.as_array(b.rbeg_hi, shape=(ns,)).astype(np.int64))
                qb.append(np.ctypeslib.as_array(b.qbeg, shape=(ns,)).astype(np.int32)); ln.append(np.ctypeslib.as_array(b.len, shape=(ns,)).astype(np.int32))
            cb += nc; sb += ns
        ch = np.concatenate(recs) if recs else np.empty((0, 4), np.uint32)
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.empty(0, dt)
        rb = ((cat(lo, np.int64) | (cat(hi, np.int64) << 32)) << 24) >> 24
        wk = ch[:, 1]
        return ChainResult(np.concatenate(offs_c).astype(np.uint32), np.concatenate(offs_s).astype(np.uint32), ch[:, 0].astype(np.int32),
                           (wk & 0x1fffffff).astype(np.int32), (((wk >> 29) & 3) | ((wk >> 31) << 8)).astype(np.int32), ch[:, 2].astype(np.int32),
                           ch[:, 3].copy(), rb, cat(qb, np.int32), cat(ln, np.int32), wire_bytes=8 * int(r.n_reads) + 16 * ch.shape[0] + 9 * rb.shape[0])

    def wait(self, set_id: int, gather: bool = True, n_threads: int = 4):
        """Waits for the set.  gather=True: a SeedResult with flat arrays in input order (cs_multi_gather; offsets as u64);
        gather=False: dict(n_reads, n_mems, n_seeds, seconds, blocks=[(device, r0, r1)]) -- the results stay where the DMA put them."""
        r = _MultiResult()
        _check(load_library().cs_multi_wait(self.h, set_id, C.byref(r)))
        self._keep.pop(set_id, None)
        info = dict(n_reads=int(r.n_reads), n_mems=int(r.n_mems), n_seeds=int(r.n_seeds), seconds=float(r.seconds),
                    blocks=[(int(r.blocks[k].device), int(r.blocks[k].r0), int(r.blocks[k].r1)) for k in range(r.n_blocks)],
                    wire_bytes=8 * int(r.n_reads) + 20 * int(r.n_mems) + 5 * int(r.n_seeds),
                    host_s=dict(submit=float(r.host_s[0]), finish_and_enqueue=float(r.host_s[1]), idle=float(r.host_s[2])),
                    gpu_ms=dict(before_kernels=float(r.gpu_ms[0]), seeding=float(r.gpu_ms[1]), collect_sa=float(r.gpu_ms[2]), results_to_host=float(r.gpu_ms[3])))
        if getattr(self, "_chaining", False):      # n_mems / n_seeds count chains / chain seeds then
            info["wire_bytes"] = 8 * int(r.n_reads) + 16 * int(r.n_mems) + 9 * int(r.n_seeds)
            if not gather:
                return info
            res = self._gather_chains(r)
            res.info = info
            return res
        if not gather:
            return info
        n = int(r.n_reads)
        mem_off = np.empty(n + 1, dtype=np.uint64); seed_off = np.empty(n + 1, dtype=np.uint64)
        mems = np.empty((int(r.n_mems), 4), dtype=np.uint64); rbeg = np.empty(int(r.n_seeds), dtype=np.int64)
        _check(load_library().cs_multi_gather(C.byref(r), _ptr(mem_off), _ptr(mems), _ptr(seed_off), _ptr(rbeg), n_threads))
        cnt = dict(ext_queries=int(r.counters.ext_queries), ext_calls=int(r.counters.ext_calls),
                   sal_queries=int(r.counters.sal_queries), sal_calls=int(r.counters.sal_calls))
        res = SeedResult(mem_off, mems, seed_off, rbeg, cnt)
        res.info = info
        return res


def replicate_index(index: "FMIndex", device: int) -> "FMIndex":
    """cs_index_replicate: a copy of the index on another device (device-to-device, over NVLink between peers)."""
    return FMIndex(load_library().cs_index_replicate(index.h, device), device)


def pack_reads_host64(bases: np.ndarray, off64: np.ndarray, n_threads: int = 1):
    """cs_pack_reads_host64: the set-global packed layout cs_multi_submit_packed takes (64-bit offsets)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    off64 = np.ascontiguousarray(off64, dtype=np.uint64)
    n = off64.shape[0] - 1
    nw = (int(off64[-1]) >> 5) + 2 * n
    packed = np.empty(nw, dtype=np.uint64)
    nmask = np.empty(nw, dtype=np.uint32)
    _check(load_library().cs_pack_reads_host64(n, _ptr(bases), _ptr(off64), _ptr(packed), _ptr(nmask), n_threads))
    return packed, nmask


def pack_reads_host(bases: np.ndarray, off: np.ndarray, n_threads: int = 1):
    """cs_pack_reads_host: the same layout as pack_reads, packed by the library's host-side C++ (no device involved)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    nw = packed_words(off)
    packed = np.empty(nw, dtype=np.uint64)
    nmask = np.empty(nw, dtype=np.uint32)
    _check(load_library().cs_pack_reads_host(off.shape[0] - 1, _ptr(bases), _ptr(off), _ptr(packed), _ptr(nmask), n_threads))
    return packed, nmask


def seed_reads(index: FMIndex, bases, off, opt: SeedOpt | None = None, batch_reads: int = 1 << 19,
               n_slots: int = 2, max_mems_per_read: int = 16, max_seeds_per_read: int = 32,
               config: "CtxConfig | None" = None) -> SeedResult:
    """Seed a whole read set: contiguous batches pipelined through the slots of one context
    (batch i+1 is submitted before batch i is waited on), results concatenated in input order.
    On CS_E_OVERFLOW the context is re-created with the capacities cs_ctx_need reports and the set is redone;
    CS_E_READ_OVERFLOW (one read beyond the per-read scratch) is raised: larger buffers cannot fix it."""
    opt = opt or SeedOpt()
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    n = off.shape[0] - 1
    if n == 0:
        z = np.zeros(1, dtype=np.uint32)
        return SeedResult(z, np.empty((0, 4), dtype=np.uint64), z.copy(), np.empty(0, dtype=np.int64))
    lens = np.diff(off.astype(np.int64))
    max_len = max(1, int(lens.max()))
    starts = list(range(0, n, batch_reads))
    max_b = max(int(off[min(n, s + batch_reads)]) - int(off[s]) for s in starts)
    cap_m, cap_s = min(batch_reads, n) * max_mems_per_read, min(batch_reads, n) * max_seeds_per_read
    while True:
        ctx = SeedContext(index, min(batch_reads, n), max(1, max_b), max_len, cap_m, cap_s, n_slots, config)
        try:
            parts: list[SeedResult] = []
            inflight: list[int] = []

            def _submit(bi: int) -> None:
                s = starts[bi]
                e = min(n, s + batch_reads)
                o = (off[s:e + 1].astype(np.int64) - int(off[s])).astype(np.uint32)
                ctx.submit(bi % n_slots, bases[int(off[s]):int(off[e])], o, opt)
                inflight.append(bi)

            nxt = 0
            while nxt < len(starts) and len(inflight) < n_slots:
                _submit(nxt)
                nxt += 1
            while inflight:
                bi = inflight[0]
                parts.append(ctx.wait(bi % n_slots))
                inflight.pop(0)
                if nxt < len(starts):
                    _submit(nxt)
                    nxt += 1
            break
        except CompSeedError as e:
            if e.code != CS_E_OVERFLOW or not inflight:
                raise
            need_m, need_s = ctx.need(inflight[0] % n_slots)
            if (need_m <= cap_m and need_s <= cap_s) or need_m >= (1 << 32) or need_s >= (1 << 32):
                raise
            cap_m, cap_s = max(cap_m, need_m), max(cap_s, need_s)
        finally:
            ctx.close()
    return concat_results(parts)


def concat_results(parts: list[SeedResult]) -> SeedResult:
    if len(parts) == 1:
        return parts[0]
    mem_off = [np.zeros(1, dtype=np.uint32)]
    seed_off = [np.zeros(1, dtype=np.uint32)]
    mb = sb = 0
    cnt: dict = {}
    ms = [0.0] * 8
    for p in parts:
        mem_off.append((p.mem_off[1:].astype(np.int64) + mb).astype(np.uint32))
        seed_off.append((p.seed_off[1:].astype(np.int64) + sb).astype(np.uint32))
        mb += int(p.mem_off[-1])
        sb += int(p.seed_off[-1])
        for k, v in p.counters.items():
            cnt[k] = cnt.get(k, 0) + v
        ms = [a + b for a, b in zip(ms, p.kernel_ms)]
    return SeedResult(np.concatenate(mem_off), np.concatenate([p.mems for p in parts]), np.concatenate(seed_off),
                      np.concatenate([p.rbeg for p in parts]), cnt, tuple(ms))


def probe_random_gather(device: int = 0, table_bytes: int = 4 << 30, granule: int = 32, n_loads: int = 1 << 28, iters: int = 3,
                        unroll: int = 1, l2_fetch_granularity: int = 0):
    """(GB/s, Gloads/s) of independent uniformly random granule-sized loads: the random-sector roofline."""
    gb, gl = C.c_double(), C.c_double()
    _check(load_library().cs_probe_random_gather_ex(device, table_bytes, granule, n_loads, iters, unroll, l2_fetch_granularity,
                                                    C.byref(gb), C.byref(gl)))
    return gb.value, gl.value


def host_register(a: np.ndarray) -> None:
    """Page-lock a numpy array so that submit() DMAs straight out of it (no staging copy)."""
    _check(load_library().cs_host_register(_ptr(a), a.nbytes))


def host_unregister(a: np.ndarray) -> None:
    _check(load_library().cs_host_unregister(_ptr(a)))


def flush_l2(device: int = 0) -> None:
    _check(load_library().cs_flush_l2(device))


# --- banded Smith-Waterman extension (cs_bsw_*: BandedPairWiseSW::scalarBandedSWAWrapper / getScores8 / getScores16) ---
# A batch of pairs is an int32 array [n, 14] in the layout of the reference's SeqPair (bandedSWA.h:91-99):
PAIR_IDR, PAIR_IDQ, PAIR_ID, PAIR_LEN1, PAIR_LEN2, PAIR_H0, PAIR_SEQID, PAIR_REGID, PAIR_SCORE, PAIR_TLE, PAIR_GTLE, PAIR_QLE, PAIR_GSCORE, PAIR_MAX_OFF = range(14)


def bwa_fill_scmat(a: int = 1, b: int = 4, ambig: int = -1) -> np.ndarray:
    """mem_opt_t.mat as bwa_fill_scmat builds it (bwalib/bwa.c:419-431): a on the diagonal, -b off it, `ambig` in the N row and column."""
    m = np.full((5, 5), -b, dtype=np.int8)
    np.fill_diagonal(m, a)
    m[4, :] = ambig
    m[:, 4] = ambig
    return m.reshape(25)


@dataclass
class BswOpt:
    """The constructor arguments of BandedPairWiseSW (bandedSWA.cpp:48-58); defaults of mem_opt_init (comp_seed.cpp:26-61)."""
    o_del: int = 6
    e_del: int = 1
    o_ins: int = 6
    e_ins: int = 1
    zdrop: int = 100
    end_bonus: int = 5
    mat: np.ndarray | None = None

    def _c(self) -> _BswOpt:
        m = self.mat if self.mat is not None else bwa_fill_scmat()
        return _BswOpt(self.o_del, self.e_del, self.o_ins, self.e_ins, self.zdrop, self.end_bonus, (C.c_int8 * 25)(*[int(x) for x in m]))


class BswExtender:
    """cs_bsw_t: batches of extension pairs on one device."""

    def __init__(self, device: int = 0, max_pairs: int = 1 << 20, max_ref_bytes: int = 1 << 26, max_qer_bytes: int = 1 << 26, max_qlen: int = 256):
        self.h = load_library().cs_bsw_create(device, max_pairs, max_ref_bytes, max_qer_bytes, max_qlen)
        if not self.h:
            raise CompSeedError(load_library().cs_last_error_code(), load_library().cs_last_error().decode())

    def close(self) -> None:
        if self.h:
            load_library().cs_bsw_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(load_library().cs_bsw_launches(self.h))

    def set_rows_in_smem(self, on: bool) -> None:
        _check(load_library().cs_bsw_set_rows_in_smem(self.h, 1 if on else 0))

    def set_ctas_per_sm(self, n: int) -> None:
        _check(load_library().cs_bsw_set_ctas_per_sm(self.h, n))

    @staticmethod
    def _check_pairs(pairs):
        if pairs.dtype != np.int32 or pairs.ndim != 2 or pairs.shape[1] != 14 or not pairs.flags.c_contiguous:
            raise ValueError("pairs: C-contiguous int32 [n, 14] (SeqPair)")

    def extend(self, pairs: np.ndarray, seq_buf_ref: np.ndarray, seq_buf_qer: np.ndarray, w: int = 100, opt: BswOpt | None = None) -> np.ndarray:
        """scalarBandedSWAWrapper(pairs, seqBufRef, seqBufQer, n, 1, w): fills score .. max_off of `pairs` in place and returns it."""
        self._check_pairs(pairs)
        o = (opt or BswOpt())._c()
        _check(load_library().cs_bsw_extend(self.h, _ptr(pairs), _ptr(seq_buf_ref), seq_buf_ref.nbytes, _ptr(seq_buf_qer), seq_buf_qer.nbytes,
                                            pairs.shape[0], w, C.byref(o)))
        return pairs

    def stage(self, pairs: np.ndarray, seq_buf_ref: np.ndarray, seq_buf_qer: np.ndarray) -> None:
        self._check_pairs(pairs)
        _check(load_library().cs_bsw_stage(self.h, _ptr(pairs), _ptr(seq_buf_ref), seq_buf_ref.nbytes, _ptr(seq_buf_qer), seq_buf_qer.nbytes, pairs.shape[0]))

    def run_staged(self, w: int = 100, opt: BswOpt | None = None):
        """Returns (kernel_ms, cells)."""
        o = (opt or BswOpt())._c()
        ms, cells = C.c_float(), C.c_uint64()
        _check(load_library().cs_bsw_run_staged(self.h, w, C.byref(o), C.byref(ms), C.byref(cells)))
        return float(ms.value), int(cells.value)

    def fetch(self, pairs: np.ndarray) -> np.ndarray:
        self._check_pairs(pairs)
        _check(load_library().cs_bsw_fetch(self.h, _ptr(pairs)))
        return pairs

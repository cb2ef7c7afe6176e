"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/compseed_b200.h declares, and fails loudly (no CPU fallback) without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from compseed_b200 import build as B
    return B.build()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "compseed_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    inline = set(re.findall(r"static inline [^(]*?\b(cs_[a-z0-9_]+)\s*\(", txt))   # accessors defined in the header itself
    return sorted(set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", txt)) - inline)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/compseed_b200.h but not exported"


def test_library_contains_sm100a_code(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device():
    import compseed_b200 as cs
    if cs.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(cs.CompSeedError):
        cs.FMIndex.upload(1, [0, 1, 2, 3, 4], 4, np.zeros(24, np.uint32), np.zeros(1, np.uint64), 32)
    with pytest.raises(cs.CompSeedError):
        cs.FMIndex.load("/nonexistent/prefix")
    with pytest.raises(cs.CompSeedError):
        cs.probe_random_gather(0, 1 << 20, 32, 1024, 1)


def test_product_path_does_not_touch_the_oracle():
    """The shipped package must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "compseed_b200")
    bad = re.compile(r"(from\s+oracle|import\s+oracle|oracle_py|liboracle|libcsref|cs_oracle|oracle/)")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), f"{f} references the oracle"


def test_option_mirror_defaults():
    import compseed_b200 as cs
    o = cs.SeedOpt()   # mem_opt_init, comp_seed.cpp:26-61
    assert (o.min_seed_len, o.split_width, o.max_mem_intv, o.max_occ, o.split_len) == (19, 10, 20, 500, 28)
    assert cs.SeedOpt(split_factor=1.0).split_len == 19


def test_pack_reads_layout():
    """Host-side packer of cs_seed_batch_submit_packed: layout of include/compseed_b200.h (no device needed)."""
    import numpy as np
    import compseed_b200 as cs
    from compseed_b200 import synth
    ref = synth.random_reference(5000, seed=3)
    bases, off, _ = synth.simulate_reads(ref, 40, [1, 31, 32, 33, 64, 100, 150, 250], 0.02, seed=4, n_rate=0.05)
    packed, nmask = cs.pack_reads(bases, off)
    assert packed.shape[0] == nmask.shape[0] == cs.packed_words(off)
    for r in range(off.shape[0] - 1):
        o, ln = int(off[r]), int(off[r + 1] - off[r])
        w0 = (o >> 5) + 2 * r
        for p in range(ln + 40):                       # 40 positions past the end: flagged, packed bits 0
            w, j = w0 + (p >> 5), p & 31
            if w >= w0 + (ln >> 5) + 2:
                break
            code = (int(packed[w]) >> (2 * j)) & 3
            flag = (int(nmask[w]) >> j) & 1
            if p < ln and bases[o + p] <= 3:
                assert flag == 0 and code == bases[o + p]
            else:
                assert flag == 1 and code == 0
    e, m = cs.pack_reads(np.zeros(0, np.uint8), np.zeros(1, np.uint32))
    assert e.shape[0] == 0 and m.shape[0] == 0
    # the library's own host-side packer (cs_pack_reads_host, plain C++ threads, no device) writes the same words
    for t in (1, 3):
        p2, m2 = cs.pack_reads_host(bases, off, t)
        assert np.array_equal(p2, packed) and np.array_equal(m2, nmask)
    big_b, big_o, _ = synth.simulate_reads(ref, 9000, [100, 150, 151], 0.02, seed=5, n_rate=0.01)
    pa, ma = cs.pack_reads(big_b, big_o)
    pb, mb = cs.pack_reads_host(big_b, big_o, 4)       # the threaded path
    assert np.array_equal(pa, pb) and np.array_equal(ma, mb)


def test_config_defaults_and_error_codes():
    """cs_index_config_default / cs_ctx_config_default mirror the dataclasses; the error codes of the header are the ones
    the Python side names."""
    import ctypes as C
    import compseed_b200 as cs
    from compseed_b200 import seeding as S
    L = cs.load_library()
    ic, cc = S._IndexConfig(), S._CtxConfig()
    L.cs_index_config_default(C.byref(ic)); L.cs_ctx_config_default(C.byref(cc))
    assert (ic.kmer_table_depth, ic.prune_k, ic.isa_intv) == (-1, -1, -1)
    d = cs.CtxConfig()
    assert (cc.use_fast, cc.use_r3_fast, cc.defer_cap, cc.lit_ctas_per_sm, cc.prefetch_results, cc.l2_persist_mb, cc.overlap_streams, cc.compact_results, cc.batch_order) == \
           (d.use_fast, d.use_r3_fast, d.defer_cap, d.lit_ctas_per_sm, d.prefetch_results, d.l2_persist_mb, d.overlap_streams, d.compact_results, d.batch_order)
    hdr = open(os.path.join(ROOT, "include", "compseed_b200.h")).read()
    codes = dict(re.findall(r"#define (CS_E_[A-Z_]+)\s+(-\d+)", hdr))
    assert int(codes["CS_E_OVERFLOW"]) == S.CS_E_OVERFLOW and int(codes["CS_E_READ_OVERFLOW"]) == S.CS_E_READ_OVERFLOW
    assert int(codes["CS_E_IO"]) == S.CS_E_IO and int(codes["CS_E_NODEVICE"]) == S.CS_E_NODEVICE


def test_library_reads_no_environment_switches():
    """Behaviour is configured through cs_index_config_t / cs_ctx_config_t, never through getenv inside the library."""
    for f in os.listdir(os.path.join(ROOT, "compseed_b200", "csrc")):
        assert "getenv" not in open(os.path.join(ROOT, "compseed_b200", "csrc", f)).read(), f


def test_compact_wire_format_spec():
    """cs_cmem_t / cs_crbeg as the header defines them, exercised through the host-side cs_compact_expand (no device):
    20 bytes per mem (three low words, start << 16 | end, five high bits of each coordinate), 40-bit sign-extended positions."""
    import ctypes as C
    import compseed_b200 as cs
    from compseed_b200 import seeding as S
    L = cs.load_library()
    rng = np.random.default_rng(1)
    n_m, n_s = 5000, 7000
    x = rng.integers(0, 1 << 37, (n_m, 3), dtype=np.uint64)
    start = rng.integers(0, 1 << 16, n_m, dtype=np.uint64)
    end = rng.integers(0, 1 << 16, n_m, dtype=np.uint64)
    rbeg = rng.integers(0, 1 << 37, n_s, dtype=np.int64)
    rbeg[:3] = [-1, 0, (1 << 37) - 1]
    cm = np.zeros((n_m, 5), dtype=np.uint32)
    cm[:, 0:3] = (x & np.uint64(0xffffffff)).astype(np.uint32)
    cm[:, 3] = ((start << np.uint64(16)) | end).astype(np.uint32)
    cm[:, 4] = ((x[:, 0] >> np.uint64(32)) | ((x[:, 1] >> np.uint64(32)) << np.uint64(5)) | ((x[:, 2] >> np.uint64(32)) << np.uint64(10))).astype(np.uint32)
    lo = (rbeg & 0xffffffff).astype(np.uint32)
    hi = ((rbeg >> 32) & 0xff).astype(np.uint8)
    r = S._CompactResult()
    r.n_reads, r.n_mems, r.n_seeds = 1, n_m, n_s
    r.cmems = cm.ctypes.data_as(C.POINTER(C.c_uint32)); r.rbeg_lo = lo.ctypes.data_as(C.POINTER(C.c_uint32)); r.rbeg_hi = hi.ctypes.data_as(C.POINTER(C.c_uint8))
    for threads in (1, 5):
        mems = np.zeros((n_m, 4), dtype=np.uint64); out = np.zeros(n_s, dtype=np.int64)
        assert L.cs_compact_expand(C.byref(r), mems.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), threads) == 0
        assert np.array_equal(mems[:, :3], x) and np.array_equal(mems[:, 3], (start << np.uint64(32)) | end)
        assert np.array_equal(out, rbeg)

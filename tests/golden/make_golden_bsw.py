"""Golden vectors of the extension stage, written by the UNMODIFIED reference in this container:
BandedPairWiseSW::scalarBandedSWAWrapper (== ksw_extend2 per pair) through oracle/_ref/libcsref.so, cross-checked against the
SIMD twins getScores8 / getScores16 on the pairs mem_chain2aln_across_reads_V2 would hand them (comp_seed.cpp:1556-1564) and
against ksw_extend2 itself (bwalib/ksw.c:380) before the fixture is written.

    python tests/golden/make_golden_bsw.py        (needs /root/reference; run `make -C oracle ref` first)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from compseed_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# option sets: (w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, a, b)
OPTS = [(100, 6, 1, 6, 1, 100, 5, 1, 4),      # defaults of mem_opt_init (comp_seed.cpp:26-61); pen_clip5 = pen_clip3 = 5
        (200, 6, 1, 6, 1, 100, 5, 1, 4),      # second band try: w << 1 (comp_seed.cpp:1718)
        (10, 4, 2, 7, 1, 20, 0, 2, 3),        # narrow band, asymmetric gap costs, small z-drop, no end bonus
        (100, 6, 1, 6, 1, 0, 5, 1, 4)]        # z-drop off


def edge_pairs():
    """Hand-made: empty target, one-base query / target, identical sequences, all-N, a long deletion / insertion in the middle,
    h0 of 1, a target much longer than query + w (the band leaves the query)."""
    rng = np.random.default_rng(5)
    q = rng.integers(0, 4, 100).astype(np.uint8)
    cases = [(q[:30], np.zeros(0, np.uint8), 50), (q[:1], q[:1], 19), (q[:1], (q[:1] + 1) & 3, 19), (q[:40], q[:1], 25), (q, q, 100),
             (np.full(20, 4, np.uint8), np.full(25, 4, np.uint8), 30), (q, np.concatenate([q[:50], rng.integers(0, 4, 30).astype(np.uint8), q[50:]]), 60),
             (q, np.concatenate([q[:40], q[70:]]), 60), (q[:60], q[:60], 1), (q[:20], rng.integers(0, 4, 400).astype(np.uint8), 40),
             (q[:20], np.concatenate([q[:20], rng.integers(0, 4, 380).astype(np.uint8)]), 40), (q[:128], q[:128], 19), (q[:127], q[:127], 0)]
    pairs = np.zeros((len(cases), 14), np.int32)
    refs, qers, ro, qo = [], [], 0, 0
    for i, (qq, tt, h0) in enumerate(cases):
        pairs[i, :6] = (ro, qo, i, tt.shape[0], qq.shape[0], h0)
        refs.append(tt); qers.append(qq); ro += tt.shape[0]; qo += qq.shape[0]
    return pairs, np.concatenate(refs), np.concatenate(qers)


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/libcsref.so missing: run `make -C oracle ref` where /root/reference exists")
    pairs, ref, qer = synth.extension_pairs(3000, seed=77, max_qlen=150)
    lp, lr, lq = synth.extension_pairs(200, seed=78, max_qlen=400, max_seed_len=250)      # longer than the 8-bit class allows
    ep, er, eq = edge_pairs()
    for extra_p, extra_r, extra_q in ((lp, lr, lq), (ep, er, eq)):
        extra_p = extra_p.copy(); extra_p[:, 0] += ref.shape[0]; extra_p[:, 1] += qer.shape[0]; extra_p[:, 2] += pairs.shape[0]
        pairs = np.concatenate([pairs, extra_p]); ref = np.concatenate([ref, extra_r]); qer = np.concatenate([qer, extra_q])
    pairs = np.ascontiguousarray(pairs, np.int32)
    out = dict(pairs=pairs[:, :8].copy(), seq_buf_ref=ref, seq_buf_qer=qer, opts=np.array(OPTS, np.int32))
    for k, (w, o_del, e_del, o_ins, e_ins, zdrop, eb, a, b) in enumerate(OPTS):
        kw = dict(w=w, o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, zdrop=zdrop, end_bonus=eb, a=a, b=b)
        sc, _ = O.ref_bsw(pairs, ref, qer, mode=0, **kw)
        # ksw_extend2 itself on a sample
        for i in list(range(0, pairs.shape[0], 97)) + list(range(pairs.shape[0] - ep.shape[0], pairs.shape[0])):
            p = pairs[i]
            if p[5] <= 0:   # ksw_extend2 asserts h0 > 0 (ksw.c:385); scalarBandedSWA has the assert commented out (bandedSWA.cpp:130)
                continue
            got = O.ref_ksw_extend2(qer[p[1]:p[1] + p[4]], ref[p[0]:p[0] + p[3]], int(p[5]), **kw)
            assert got == (sc[i, 8], sc[i, 11], sc[i, 9], sc[i, 10], sc[i, 12], sc[i, 13]), (i, got, sc[i])
        if k < 2:   # the SIMD twins hard-code the default N score and saturate on other penalties: checked on the default sets
            m8 = (pairs[:, 3] < 128) & (pairs[:, 4] < 128) & (pairs[:, 5] + np.minimum(pairs[:, 3], pairs[:, 4]) * a < 128)
            v8, _ = O.ref_bsw(np.ascontiguousarray(pairs[m8]), ref, qer, mode=1, **kw)
            v16, _ = O.ref_bsw(np.ascontiguousarray(pairs[~m8]), ref, qer, mode=2, **kw)
            assert np.array_equal(v8[:, 8:], sc[m8][:, 8:]), "getScores8 disagrees with scalarBandedSWAWrapper"
            assert np.array_equal(v16[:, 8:], sc[~m8][:, 8:]), "getScores16 disagrees with scalarBandedSWAWrapper"
        out["res%d" % k] = sc[:, 8:].copy()
    np.savez_compressed(os.path.join(OUT, "bsw3k.npz"), **out)
    print("bsw3k.npz:", pairs.shape[0], "pairs,", ref.shape[0] + qer.shape[0], "bases")


if __name__ == "__main__":
    main()

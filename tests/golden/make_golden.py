"""Generates the committed golden vectors by running the UNMODIFIED reference in this container:
oracle/_ref/bwaidx builds the index, oracle/_ref/libcsref.so (reference objects + harness) produces
mems and seeds through bwt_smem1/bwt_seed_strategy1/bwt_sa ("bwamem") and through
collect_mem_with_sst/tem_forward_sst ("compseed").  Both must agree before a fixture is written.

    python tests/golden/make_golden.py        (needs /root/reference; run `make -C oracle ref` first)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from compseed_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (reference maker, reads maker, option sets)
    "random20k": dict(ref=lambda: synth.random_reference(20_000, seed=11),
                      reads=lambda ref: synth.simulate_reads(ref, 400, [100, 150, 250], 0.01, seed=12, n_rate=0.002),
                      opts=[dict(), dict(min_seed_len=15, split_factor=1.0, split_width=5, max_mem_intv=10, max_occ=50),
                            dict(split_factor=2.5, max_mem_intv=0)]),
    "repeat30k": dict(ref=lambda: synth.repeat_rich_reference(30_000, seed=21, n_segdup=40, segdup_len=600, n_tandem=12),
                      reads=lambda ref: synth.simulate_reads(ref, 300, [100, 150, 250], 0.02, seed=22, n_rate=0.004),
                      opts=[dict(), dict(max_occ=20, split_width=30, max_mem_intv=40)]),
}


def edge_reads(ref):
    """Hand-made edge cases: exact full-length match, all-N, N at both ends,
    read shorter than min_seed_len, a single base, homopolymer, reverse-complement exact match."""
    reads = [ref[100:250].copy(), np.full(60, 4, np.uint8), ref[300:450].copy(), ref[500:510].copy(),
             ref[7:8].copy(), np.zeros(80, np.uint8), (3 - ref[1000:1200])[::-1].copy(), ref[2000:2019].copy(),
             ref[2100:2120].copy(), ref[2200:2228].copy(), ref[2300:2329].copy()]
    reads[2][0] = 4
    reads[2][-1] = 4
    reads[2][70] = 4
    off = np.zeros(len(reads) + 1, np.uint32)
    off[1:] = np.cumsum([len(r) for r in reads])
    return np.concatenate(reads).astype(np.uint8), off


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/libcsref.so missing: run `make -C oracle ref` where /root/reference exists")
    for name, case in CASES.items():
        ref = case["ref"]()
        with tempfile.TemporaryDirectory() as d:
            synth.write_fasta(os.path.join(d, "ref.fa"), ref)
            O.bwaidx(os.path.join(d, "ref.fa"), os.path.join(d, "ref"))
            ri = O.RefIndex.load(os.path.join(d, "ref"))
            bases, off, _ = case["reads"](ref)
            eb, eo = edge_reads(ref)
            bases = np.concatenate([bases, eb])
            off = np.concatenate([off, (eo[1:].astype(np.int64) + int(off[-1])).astype(np.uint32)])
            out = dict(ref=ref, primary=np.uint64(ri.primary), L2=ri.L2, seq_len=np.uint64(ri.seq_len), bwt=ri.bwt.copy(),
                       sa=ri.sa.copy(), sa_intv=np.int32(ri.sa_intv), bases=bases, off=off)
            rng = np.random.default_rng(5)
            k = np.concatenate([rng.integers(0, ri.seq_len + 1, 500, dtype=np.uint64),
                                np.array([0, ri.primary - 1, ri.primary, ri.primary + 1, ri.seq_len - 1, ri.seq_len, 2**64 - 1], dtype=np.uint64)])
            out["occ_k"], out["occ_cnt"] = k, ri.occ4(k)
            sk = np.concatenate([rng.integers(1, ri.seq_len + 1, 300, dtype=np.uint64), np.array([ri.primary, ri.seq_len, 1, 32, 33], dtype=np.uint64)])
            out["sa_k"], out["sa_v"] = sk, ri.sa_lookup(sk)
            for i, o in enumerate(case["opts"]):
                a = ri.seed(bases, off, "bwamem", **o)
                b = ri.seed(bases, off, "compseed", **o)
                # the two callers round split_len differently (bwamem.c:223 vs comp_seed.cpp:2279); equal for these sets
                assert a.same_as(b), f"{name}/{i}: reference's two seeding paths disagree"
                out[f"opt{i}"] = np.array([o.get("min_seed_len", 19), synth.split_len_bwamem(o.get("min_seed_len", 19), o.get("split_factor", 1.5)),
                                           o.get("split_width", 10), o.get("max_mem_intv", 20), o.get("max_occ", 500)], dtype=np.int32)
                out[f"mem_off{i}"], out[f"mems{i}"], out[f"seed_off{i}"], out[f"rbeg{i}"] = a.mem_off, a.mems, a.seed_off, a.rbeg
                print(name, i, "reads", off.shape[0] - 1, "mems", a.mems.shape[0], "seeds", a.rbeg.shape[0], b.counters)
                if i == 0:   # chains of the default option set: the reference's own mem_chain + mem_chain_flt (comp_seed.cpp:241-354)
                    ch = O.ref_chain(off, a, [ref.shape[0]], **o)
                    for k in ("pos", "rid", "w", "kept", "n", "s_rbeg", "s_qbeg", "s_len", "frac_rep"):
                        out[f"chain_{k}0"] = getattr(ch, k)
                    out["chain_off0"] = ch.chain_off
                    print(name, "chains", ch.pos.shape[0], "chain seeds", ch.s_rbeg.shape[0])
            # extend probes: intervals taken from real mems plus the four 1-base intervals
            iv = [a.mems[:200, :3]]
            for c in range(4):
                iv.append(np.array([[ri.L2[c] + 1, ri.L2[3 - c] + 1, ri.L2[c + 1] - ri.L2[c]]], dtype=np.uint64))
            iv = np.concatenate(iv).astype(np.uint64)
            isb = (np.arange(iv.shape[0]) & 1).astype(np.int32)
            out["ext_ik"], out["ext_back"], out["ext_ok"] = iv, isb, ri.extend(iv, isb)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
            print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()

"""Seeded synthetic references and reordered read sets (BASELINE.json configs).

Pure numpy; no dependency on the oracle or on the CUDA library.  Bases are nt4 codes as the
reference uses them after conversion (FM_index/bntseq.c:46-63): A,C,G,T -> 0..3, anything else 4.
"""
from __future__ import annotations

import numpy as np

NT4 = np.full(256, 4, dtype=np.uint8)
for _i, _ch in enumerate("ACGT"):
    NT4[ord(_ch)] = _i
    NT4[ord(_ch.lower())] = _i
NT4[ord("-")] = 5  # bntseq.c:49
CODE2ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)


def random_reference(l_pac: int, seed: int = 20261018) -> np.ndarray:
    """i.i.d. uniform ACGT reference of l_pac bases (configs 1-3)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=l_pac, dtype=np.uint8)


def repeat_rich_reference(l_pac: int, seed: int = 7, n_segdup: int = 100, segdup_len: int = 3000,
                          n_tandem: int = 200, divergence: float = 0.02) -> np.ndarray:
    """Random backbone + segmental duplications + tandem repeats (config 4): many mems with
    x[2] > max_occ and with split_width < x[2] < max_mem_intv."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, size=l_pac, dtype=np.uint8)
    # segmental duplications: a few source segments copied many times with 0..divergence substitutions
    n_src = max(1, n_segdup // 25)
    for s in range(n_src):
        seg_len = min(segdup_len, l_pac // 4)
        src = int(rng.integers(0, l_pac - seg_len))
        seg = ref[src:src + seg_len].copy()
        for _ in range(n_segdup // n_src):
            dst = int(rng.integers(0, l_pac - seg_len))
            cp = seg.copy()
            d = float(rng.uniform(0, divergence))
            mut = rng.random(seg_len) < d
            cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            if rng.random() < 0.5:
                cp = (3 - cp)[::-1]
            ref[dst:dst + seg_len] = cp
    # tandem repeats
    for _ in range(n_tandem):
        unit = int(rng.integers(2, 61))
        copies = int(rng.integers(50, 600))
        tot = min(unit * copies, l_pac // 8)
        dst = int(rng.integers(0, l_pac - tot))
        u = rng.integers(0, 4, size=unit, dtype=np.uint8)
        ref[dst:dst + tot] = np.resize(u, tot)
    return ref


def simulate_reads(ref: np.ndarray, n_reads: int, read_len=150, sub_rate: float = 0.01, seed: int = 1,
                   sort_by_pos: bool = True, n_rate: float = 0.0, window: tuple[int, int] | None = None):
    """Sample reads at uniform start positions, `sub_rate` substitutions per base, 50 % reverse
    complemented, position-sorted to mimic SPRING reordering (config 1).  `read_len` may be an int
    or a sequence of lengths to mix.  Returns (bases u8 concatenated, offsets u32[n+1], pos i64[n])."""
    rng = np.random.default_rng(seed)
    l_pac = ref.shape[0]
    lens_choice = np.atleast_1d(np.asarray(read_len, dtype=np.int64))
    lens = lens_choice[rng.integers(0, lens_choice.size, size=n_reads)]
    lo, hi = (0, l_pac) if window is None else window
    pos = rng.integers(lo, np.maximum(lo + 1, hi - lens), size=n_reads)
    if sort_by_pos:
        order = np.argsort(pos, kind="stable")
        pos, lens = pos[order], lens[order]
    off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    rid = np.repeat(np.arange(n_reads, dtype=np.int64), lens)
    within = np.arange(total, dtype=np.int64) - off[rid]
    rev = rng.random(n_reads) < 0.5
    src = np.where(rev[rid], pos[rid] + lens[rid] - 1 - within, pos[rid] + within)
    bases = ref[src]
    bases = np.where(rev[rid], 3 - bases, bases).astype(np.uint8)
    mut = rng.random(total) < sub_rate
    bases[mut] = (bases[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    if n_rate > 0:
        bases[rng.random(total) < n_rate] = 4
    return bases, off.astype(np.uint32), pos


def boundary_reads(ref: np.ndarray, seed: int = 5):
    """Crafted reads for the places where index arithmetic changes character (SURVEY.md section 8c and the
    32-base word logic of the device kernels): matches at the first and last text position of both strands,
    matches that bridge the forward / reverse-complement boundary, read lengths around multiples of 32 and
    around 255/256, a substitution or an N at each word edge, an N next to a substitution."""
    rng = np.random.default_rng(seed)
    l_pac = ref.shape[0]
    text = np.concatenate([ref, (3 - ref)[::-1]])          # what the FM-index indexes: fwd + revcomp(fwd)
    reads = []

    def add(seg, subs=(), ns=()):
        r = np.array(seg, dtype=np.uint8, copy=True)
        for p in subs:
            if 0 <= p < r.shape[0]:
                r[p] = (r[p] + 1 + int(rng.integers(0, 3))) & 3
        for p in ns:
            if 0 <= p < r.shape[0]:
                r[p] = 4
        reads.append(r)

    edges = (0, 1, 30, 31, 32, 33, 62, 63, 64, 65, 95, 96, 127, 128, 148, 149)
    for start in (0, 1, l_pac - 150, l_pac - 151, l_pac - 75, l_pac, l_pac + 1, 2 * l_pac - 150, 2 * l_pac - 151):
        seg = text[start:start + 150]                      # text start/end of either strand, or across the strand boundary
        add(seg)
        for e in edges:
            add(seg, subs=(e,))
        add(seg, subs=(40, 41)); add(seg, subs=(20, 100)); add(seg, ns=(31,)); add(seg, ns=(32,), subs=(33,)); add(seg, subs=(63,), ns=(64,))
        add(seg, ns=(0,)); add(seg, ns=(149,)); add(seg, subs=(0, 149))
    mid = l_pac // 3
    for L in (19, 20, 21, 27, 28, 29, 31, 32, 33, 38, 39, 40, 63, 64, 65, 95, 96, 97, 127, 128, 129, 159, 160, 161, 191, 192, 193, 223, 224, 225, 250, 254, 255, 256):
        add(text[mid:mid + L])
        add(text[mid + 1000:mid + 1000 + L], subs=(L // 2,))
        add(text[mid + 2000:mid + 2000 + L], subs=(L // 3, 2 * L // 3))
        add(text[2 * l_pac - mid - L:2 * l_pac - mid], ns=(L // 2,))
    off = np.zeros(len(reads) + 1, np.uint32)
    off[1:] = np.cumsum([r.shape[0] for r in reads])
    return np.concatenate(reads).astype(np.uint8), off


def shuffle_reads(bases: np.ndarray, off: np.ndarray, seed: int = 3):
    """Random permutation of a read set (the 'shuffled' arm of config 5)."""
    rng = np.random.default_rng(seed)
    n = off.shape[0] - 1
    perm = rng.permutation(n)
    lens = np.diff(off.astype(np.int64))[perm]
    new_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    rid = np.repeat(np.arange(n, dtype=np.int64), lens)
    within = np.arange(int(new_off[-1]), dtype=np.int64) - new_off[rid]
    return bases[off.astype(np.int64)[perm][rid] + within], new_off.astype(np.uint32), perm


def write_fasta(path: str, ref: np.ndarray, name: str = "chr1", width: int = 80) -> None:
    asc = CODE2ASCII[ref]
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, asc.shape[0], width * 4096):
            blk = asc[i:i + width * 4096]
            n_full = blk.shape[0] // width
            if n_full:
                rows = blk[:n_full * width].reshape(n_full, width)
                out = np.concatenate([rows, np.full((n_full, 1), 10, dtype=np.uint8)], axis=1)
                f.write(out.tobytes())
            if blk.shape[0] % width:
                f.write(blk[n_full * width:].tobytes() + b"\n")


def write_reads_txt(path: str, bases: np.ndarray, off: np.ndarray) -> None:
    """One read per line, every line '\\n'-terminated (main.cpp:36-58 line reader)."""
    asc = CODE2ASCII[np.minimum(bases, 4)]
    with open(path, "wb") as f:
        for r in range(off.shape[0] - 1):
            f.write(asc[off[r]:off[r + 1]].tobytes() + b"\n")


def split_len_bwamem(min_seed_len: int, split_factor: float) -> int:
    """bwamem.c:223: int * float in float, + .499 promoted to double."""
    return int(float(np.float32(min_seed_len) * np.float32(split_factor)) + .499)


def split_len_compseed(min_seed_len: int, split_factor: float) -> int:
    """comp_seed.cpp:2279: 1.0 * int * float in double."""
    return int(1.0 * min_seed_len * float(np.float32(split_factor)) + .499)


# ---------------------------------------------------------------------------------------------
# Large workloads (configs 2/3): generated on the GPU with torch (plumbing only) because numpy
# would need minutes and tens of GB for 3.1 Gbp / 10 M reads.  Seeded and deterministic for a given
# torch build.
# ---------------------------------------------------------------------------------------------
def random_reference_torch(l_pac: int, seed: int, device: str):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(l_pac, dtype=torch.uint8, device=device)
    step = 1 << 28
    for s in range(0, l_pac, step):
        e = min(l_pac, s + step)
        out[s:e] = torch.randint(0, 4, (e - s,), dtype=torch.uint8, device=device, generator=g)
    return out


def simulate_reads_torch(ref, n_reads: int, read_len: int, sub_rate: float, seed: int, window: tuple[int, int] | None = None,
                         chunk: int = 1 << 20):
    """Fixed-length reads, uniform starts inside `window`, position-sorted, 50 % reverse-complemented,
    `sub_rate` substitutions.  `ref` is a uint8 CUDA tensor.  Returns numpy (bases u8, offsets u32, pos i64)."""
    import torch
    dev = ref.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    l_pac = ref.shape[0]
    lo, hi = (0, l_pac) if window is None else window
    pos = torch.randint(lo, max(lo + 1, hi - read_len), (n_reads,), device=dev, generator=g, dtype=torch.int64)
    pos, _ = torch.sort(pos)
    bases = np.empty(n_reads * read_len, dtype=np.uint8)
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)
    for s in range(0, n_reads, chunk):
        e = min(n_reads, s + chunk)
        p = pos[s:e]
        rev = torch.rand(e - s, device=dev, generator=g) < 0.5
        idx = torch.where(rev[:, None], p[:, None] + (read_len - 1) - ar[None, :], p[:, None] + ar[None, :])
        b = ref[idx]
        b = torch.where(rev[:, None], 3 - b, b)
        mut = torch.rand((e - s, read_len), device=dev, generator=g) < sub_rate
        add = torch.randint(1, 4, (e - s, read_len), device=dev, generator=g, dtype=torch.uint8)
        b = torch.where(mut, (b + add) & 3, b).to(torch.uint8)
        bases[s * read_len:e * read_len] = b.reshape(-1).cpu().numpy()
    off = (np.arange(n_reads + 1, dtype=np.int64) * read_len).astype(np.uint32)
    return bases, off, pos.cpu().numpy()


def extension_pairs(n: int, seed: int, max_qlen: int = 150, sub_rate: float = 0.02, indel_rate: float = 0.004, n_rate: float = 0.002,
                    unrelated_frac: float = 0.05, a: int = 1, w: int = 100, max_seed_len: int = 150, extra_gap: bool = True):
    """Synthetic input of the extension stage in the shape mem_chain2aln_across_reads_V2 builds it (comp_seed.cpp:1480-1660): for each
    pair a query of 1..max_qlen bases (the part of a read left or right of a seed), a target that is a mutated copy of it (substitutions,
    short indels, a few Ns) followed by unrelated flank up to the length the caller fetches (query + cal_max_gap, bwamem.c:66-73, capped
    at 2w), and h0 = seed length x a.  A fraction of the pairs has an unrelated target (extension dies at once / z-drop), some have an
    empty or one-base target, some a long indel.  Returns (pairs int32 [n, 14] (SeqPair layout), seq_buf_ref u8, seq_buf_qer u8)."""
    rng = np.random.default_rng(seed)
    pairs = np.zeros((n, 14), dtype=np.int32)
    refs, qers = [], []
    ro = qo = 0
    for i in range(n):
        ql = int(rng.integers(1, max_qlen + 1))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        kind = rng.random()
        if kind < unrelated_frac:
            t = rng.integers(0, 4, int(rng.integers(0, ql + 40))).astype(np.uint8)
        else:
            t = []
            j = 0
            long_indel = rng.random() < 0.03
            at = int(rng.integers(0, ql)) if long_indel else -1
            while j < ql:
                if j == at:
                    k = int(rng.integers(5, 40))
                    if rng.random() < 0.5:
                        t.extend(rng.integers(0, 4, k).tolist())      # deletion from the read = extra target bases
                    else:
                        j += k                                         # insertion in the read = target skips query bases
                        continue
                r = rng.random()
                if r < indel_rate:
                    t.append(int(rng.integers(0, 4)))
                    continue
                if r < 2 * indel_rate:
                    j += 1
                    continue
                b = int(q[j])
                if rng.random() < sub_rate:
                    b = (b + int(rng.integers(1, 4))) & 3
                t.append(b)
                j += 1
            # flank: what bns_fetch_seq returns beyond the end of the read (max gap for the rest of the query, capped at 2w)
            gap = min(2 * w, max(1, int((ql * a - 6) / 1.0 + 1))) if extra_gap else 0
            t.extend(rng.integers(0, 4, int(rng.integers(0, gap + 1))).tolist())
            t = np.asarray(t, dtype=np.uint8)
            if rng.random() < 0.01:
                t = t[:int(rng.integers(0, 2))]
        if n_rate > 0:
            q = q.copy(); t = t.copy()
            q[rng.random(q.shape[0]) < n_rate] = 4
            t[rng.random(t.shape[0]) < n_rate] = 4
        pairs[i, 0], pairs[i, 1], pairs[i, 2] = ro, qo, i
        pairs[i, 3], pairs[i, 4] = t.shape[0], ql
        pairs[i, 5] = int(rng.integers(19, max_seed_len + 1)) * a
        pairs[i, 6], pairs[i, 7] = i // 3, i % 3
        refs.append(t); qers.append(q)
        ro += t.shape[0]; qo += ql
    ref = np.concatenate(refs) if refs else np.zeros(0, np.uint8)
    qer = np.concatenate(qers) if qers else np.zeros(0, np.uint8)
    return pairs, np.ascontiguousarray(ref, dtype=np.uint8), np.ascontiguousarray(qer, dtype=np.uint8)


def extension_pairs_fast(n: int, seed: int, max_qlen: int = 150, sub_rate: float = 0.02, indel_frac: float = 0.15, n_rate: float = 0.001,
                         unrelated_frac: float = 0.03, a: int = 1, w: int = 100, max_seed_len: int = 150):
    """Vectorised variant of extension_pairs for large batches (benchmarks, full-size tests): queries are slices of a random genome
    with substitutions; the target of a pair is the same stretch of the genome plus the flank the caller would fetch (up to
    min(2w, query length) more bases), for `indel_frac` of the pairs with one indel of 1..12 bases at a random position, for
    `unrelated_frac` of them taken from somewhere else.  Same return value as extension_pairs."""
    rng = np.random.default_rng(seed)
    glen = 1 << 22
    g = rng.integers(0, 4, glen + 4096, dtype=np.uint8)
    ql = rng.integers(1, max_qlen + 1, n).astype(np.int64)
    flank = (rng.random(n) * (np.minimum(2 * w, ql) + 1)).astype(np.int64)
    d = np.where(rng.random(n) < indel_frac, rng.integers(-12, 13, n), 0).astype(np.int64)       # > 0: the target skips d genome bases
    tl = np.maximum(ql + flank - np.maximum(d, 0) * 0, 0)
    k1 = (rng.random(n) * ql).astype(np.int64)                                                 # where the indel sits
    s = rng.integers(16, glen - 2 * max_qlen - 512, n).astype(np.int64)
    ts = np.where(rng.random(n) < unrelated_frac, rng.integers(16, glen - 2 * max_qlen - 512, n), s)
    qoff = np.zeros(n + 1, np.int64); qoff[1:] = np.cumsum(ql)
    roff = np.zeros(n + 1, np.int64); roff[1:] = np.cumsum(tl)
    qi = np.arange(qoff[-1], dtype=np.int64) - np.repeat(qoff[:-1], ql)
    qer = g[np.repeat(s, ql) + qi].copy()
    sub = rng.random(qer.shape[0]) < sub_rate
    qer[sub] = (qer[sub] + rng.integers(1, 4, int(sub.sum())).astype(np.uint8)) & 3
    ti = np.arange(roff[-1], dtype=np.int64) - np.repeat(roff[:-1], tl)
    shift = np.where(ti >= np.repeat(k1, tl), np.repeat(d, tl), 0)
    ref = g[np.clip(np.repeat(ts, tl) + ti + shift, 0, glen + 4095)].copy()
    if n_rate > 0:
        qer[rng.random(qer.shape[0]) < n_rate] = 4
        ref[rng.random(ref.shape[0]) < n_rate] = 4
    pairs = np.zeros((n, 14), dtype=np.int32)
    pairs[:, 0] = roff[:-1]; pairs[:, 1] = qoff[:-1]; pairs[:, 2] = np.arange(n)
    pairs[:, 3] = tl; pairs[:, 4] = ql
    pairs[:, 5] = rng.integers(19, max_seed_len + 1, n) * a
    pairs[:, 6] = np.arange(n) // 3; pairs[:, 7] = np.arange(n) % 3
    return pairs, np.ascontiguousarray(ref), np.ascontiguousarray(qer)

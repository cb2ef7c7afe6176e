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

"""Multi-GPU read sharding, host-side helpers around the split of cs_multi_*: the index is replicated on every GPU, reads are
split in CONTIGUOUS blocks in input order (neighbouring reordered reads stay on one GPU), each block
a whole multiple of the reference's reuse block (BATCH_SIZE 512, comp_seed.h:36), and results are
concatenated on the host by block index.  There is no collective on the data path."""
from __future__ import annotations

import ctypes as C

import numpy as np

REUSE_BLOCK = 512


def shard_bounds(n_reads: int, world: int) -> list[tuple[int, int]]:
    """[start, end) of each rank's contiguous block; all but the last are multiples of 512.  This IS the split the
    library's multi-device pipeline applies (cs_multi_block_bounds, csrc/cs_multi.cu): host arithmetic inside the C-ABI."""
    from .seeding import load_library
    L = load_library()
    L.cs_multi_block_bounds.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.cs_multi_block_bounds.restype = None
    out = []
    for r in range(world):
        a, b = C.c_uint64(), C.c_uint64()
        L.cs_multi_block_bounds(n_reads, world, r, C.byref(a), C.byref(b))
        out.append((int(a.value), int(b.value)))
    return out


def take_shard(bases: np.ndarray, off: np.ndarray, bounds: tuple[int, int]):
    s, e = bounds
    o = off[s:e + 1].astype(np.int64)
    return bases[int(o[0]):int(o[-1])], (o - o[0]).astype(np.uint32)


def gather_in_input_order(parts):
    """Concatenate per-rank (mem_off, mems, seed_off, rbeg) tuples, ordered by rank == input order."""
    mem_off, seed_off = [np.zeros(1, dtype=np.uint32)], [np.zeros(1, dtype=np.uint32)]
    mb = sb = 0
    for p in parts:
        mem_off.append((p[0][1:].astype(np.int64) + mb).astype(np.uint32))
        seed_off.append((p[2][1:].astype(np.int64) + sb).astype(np.uint32))
        mb += int(p[0][-1])
        sb += int(p[2][-1])
    return (np.concatenate(mem_off), np.concatenate([p[1] for p in parts]), np.concatenate(seed_off),
            np.concatenate([p[3] for p in parts]))
